// host_pipeline.cu -- b200_adam_host_step_f32: fused differentiable Adam forward+backward over HOST arrays.
//
// The arrays are cut into slices; slice k is copied host->device on the H2D stream, processed by the two fused
// kernels on the compute stream and copied back on the D2H stream, with NSLOT device staging slots in rotation,
// so the PCIe link runs in both directions while the kernels (which need ~1% of the time) hide completely.
// This is the "application holds host buffers" path that bench.py times as `e2e`.
#include <cuda_runtime.h>

#include <mutex>

#include "b200_adam.h"

namespace {

constexpr int NSLOT = 3;
constexpr int NARR = 12;  // g mu nu du dme dne | mu' nu' u dg dmu dnu

struct Pipeline {
  int device = -1;
  int64_t slice = 0;
  float *buf[NSLOT][NARR] = {};
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t e_h2d[NSLOT] = {}, e_comp[NSLOT] = {}, e_d2h[NSLOT] = {};
  bool ready = false;
  std::mutex mu;  // one caller at a time per device; callers on different devices do not serialise

  // Frees whatever exists (a creation that failed midway leaves a partial set: every handle is checked on its own).
  void destroy() {
    for (int s = 0; s < NSLOT; ++s) {
      for (int a = 0; a < NARR; ++a)
        if (buf[s][a]) cudaFree(buf[s][a]), buf[s][a] = nullptr;
      if (e_h2d[s]) cudaEventDestroy(e_h2d[s]), e_h2d[s] = nullptr;
      if (e_comp[s]) cudaEventDestroy(e_comp[s]), e_comp[s] = nullptr;
      if (e_d2h[s]) cudaEventDestroy(e_d2h[s]), e_d2h[s] = nullptr;
    }
    if (s_h2d) cudaStreamDestroy(s_h2d), s_h2d = nullptr;
    if (s_comp) cudaStreamDestroy(s_comp), s_comp = nullptr;
    if (s_d2h) cudaStreamDestroy(s_d2h), s_d2h = nullptr;
    ready = false;
  }

  int create(int dev, int64_t want_slice) {
    device = dev, slice = want_slice;
    cudaError_t e;
    for (int s = 0; s < NSLOT; ++s) {
      for (int a = 0; a < NARR; ++a)
        if ((e = cudaMalloc(&buf[s][a], (size_t)slice * sizeof(float))) != cudaSuccess) return (int)e;
      if ((e = cudaEventCreateWithFlags(&e_h2d[s], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
      if ((e = cudaEventCreateWithFlags(&e_comp[s], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
      if ((e = cudaEventCreateWithFlags(&e_d2h[s], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    }
    if ((e = cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    ready = true;
    return 0;
  }

  int ensure(int dev, int64_t want_slice) {
    if (ready && dev == device && want_slice == slice) return 0;
    destroy();
    const int rc = create(dev, want_slice);
    if (rc != 0) destroy();  // no partial pipeline survives a failed creation
    return rc;
  }

  // After an error in the middle of a call copies may still be in flight into / out of the caller's host buffers:
  // wait for them before the error is reported, so the caller may free or reuse its arrays.
  void drain() {
    if (s_h2d) cudaStreamSynchronize(s_h2d);
    if (s_comp) cudaStreamSynchronize(s_comp);
    if (s_d2h) cudaStreamSynchronize(s_d2h);
  }
};

constexpr int MAX_DEVICES = 64;
Pipeline g_pipes[MAX_DEVICES];  // one per device

}  // namespace

#define CK(x)                                \
  do {                                       \
    cudaError_t _e = (x);                    \
    if (_e != cudaSuccess) return (int)_e;   \
  } while (0)

static int host_step_locked(Pipeline &p, const float *const h_in[6], float *const h_out[6], int64_t n, double b1,
                            double b2, double eps, double eps_root, uint64_t count, int64_t slice_elems) {
  int64_t k = 0;
  for (int64_t off = 0; off < n; off += slice_elems, ++k) {
    const int s = (int)(k % NSLOT);
    const int64_t len = (n - off < slice_elems) ? (n - off) : slice_elems;
    const size_t bytes = (size_t)len * sizeof(float);
    float **b = p.buf[s];
    // the slot is free once the D2H of slice k-NSLOT has drained
    if (k >= NSLOT) CK(cudaStreamWaitEvent(p.s_h2d, p.e_d2h[s], 0));
    for (int a = 0; a < 6; ++a) CK(cudaMemcpyAsync(b[a], h_in[a] + off, bytes, cudaMemcpyHostToDevice, p.s_h2d));
    CK(cudaEventRecord(p.e_h2d[s], p.s_h2d));
    CK(cudaStreamWaitEvent(p.s_comp, p.e_h2d[s], 0));
    {
      const void *pg = b[0], *pmu = b[1], *pnu = b[2];
      void *pmo = b[6], *pno = b[7], *puo = b[8];
      int rc = b200_adam_fused_forward(B200_ADAM_F32, 1, &pg, &pmu, &pnu, &pmo, &pno, &puo, &len, &count, b1, b2, eps,
                                       eps_root, p.s_comp);
      if (rc) return rc;
      const void *pdu = b[3], *pu = b[8], *pm = b[6], *pme = b[4], *pne = b[5];
      void *pdg = b[9], *pdm = b[10], *pdn = b[11];
      rc = b200_adam_fused_backward(B200_ADAM_F32, 1, &pdu, &pu, &pm, &pg, &pme, &pne, &pdg, &pdm, &pdn, &len, &count,
                                    b1, b2, eps_root, p.s_comp);
      if (rc) return rc;
    }
    CK(cudaEventRecord(p.e_comp[s], p.s_comp));
    CK(cudaStreamWaitEvent(p.s_d2h, p.e_comp[s], 0));
    for (int a = 0; a < 6; ++a) CK(cudaMemcpyAsync(h_out[a] + off, b[6 + a], bytes, cudaMemcpyDeviceToHost, p.s_d2h));
    CK(cudaEventRecord(p.e_d2h[s], p.s_d2h));
  }
  CK(cudaStreamSynchronize(p.s_d2h));
  return 0;
}

extern "C" int b200_adam_host_step_f32(const float *g, const float *mu, const float *nu, const float *du,
                                       const float *dmu_ext, const float *dnu_ext, float *mu_out, float *nu_out,
                                       float *u_out, float *dg, float *dmu, float *dnu, int64_t n, double b1,
                                       double b2, double eps, double eps_root, uint64_t count,
                                       int64_t slice_elems) {
  if (n < 0) return B200_ADAM_E_ARG;
  if (n == 0) return 0;
  if (!g || !mu || !nu || !du || !dmu_ext || !dnu_ext || !mu_out || !nu_out || !u_out || !dg || !dmu || !dnu)
    return B200_ADAM_E_ARG;
  if (slice_elems <= 0) slice_elems = 1 << 22;  // 16 MiB per array per slice
  if (slice_elems > n) slice_elems = (n + 7) & ~int64_t(7);
  slice_elems = (slice_elems + 7) & ~int64_t(7);

  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) return B200_ADAM_E_NO_DEVICE;
  Pipeline &p = g_pipes[dev];
  std::lock_guard<std::mutex> lock(p.mu);
  if (int rc = p.ensure(dev, slice_elems)) return rc;
  const float *h_in[6] = {g, mu, nu, du, dmu_ext, dnu_ext};
  float *h_out[6] = {mu_out, nu_out, u_out, dg, dmu, dnu};
  const int rc = host_step_locked(p, h_in, h_out, n, b1, b2, eps, eps_root, count, slice_elems);
  if (rc != 0) p.drain();  // nothing of this call may still touch the caller's buffers when the error is reported
  return rc;
}

// Frees the staging buffers/streams of b200_adam_host_step_f32 on every device (optional; they are reused across calls).
extern "C" int b200_adam_host_release(void) {
  int cur = 0;
  const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
  for (int d = 0; d < MAX_DEVICES; ++d) {
    Pipeline &p = g_pipes[d];
    std::lock_guard<std::mutex> lock(p.mu);
    if (p.device < 0) continue;
    if (cudaSetDevice(p.device) == cudaSuccess) p.destroy();
    p.device = -1;
  }
  if (have_cur) cudaSetDevice(cur);
  return 0;
}
