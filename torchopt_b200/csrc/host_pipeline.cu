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

  void destroy() {
    if (!ready) return;
    for (int s = 0; s < NSLOT; ++s) {
      for (int a = 0; a < NARR; ++a) cudaFree(buf[s][a]), buf[s][a] = nullptr;
      cudaEventDestroy(e_h2d[s]), cudaEventDestroy(e_comp[s]), cudaEventDestroy(e_d2h[s]);
    }
    cudaStreamDestroy(s_h2d), cudaStreamDestroy(s_comp), cudaStreamDestroy(s_d2h);
    ready = false;
  }

  int ensure(int dev, int64_t want_slice) {
    if (ready && dev == device && want_slice == slice) return 0;
    destroy();
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
};

Pipeline g_pipe;
std::mutex g_pipe_mu;

}  // namespace

#define CK(x)                          \
  do {                                 \
    cudaError_t _e = (x);              \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

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

  std::lock_guard<std::mutex> lock(g_pipe_mu);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (int rc = g_pipe.ensure(dev, slice_elems)) return rc;
  Pipeline &p = g_pipe;

  const float *h_in[6] = {g, mu, nu, du, dmu_ext, dnu_ext};
  float *h_out[6] = {mu_out, nu_out, u_out, dg, dmu, dnu};
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

// Frees the staging buffers/streams of b200_adam_host_step_f32 (optional; they are reused across calls).
extern "C" int b200_adam_host_release(void) {
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  g_pipe.destroy();
  return 0;
}
