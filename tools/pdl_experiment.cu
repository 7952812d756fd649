// pdl_experiment.cu -- would programmatic dependent launch (PDL) shorten the gap between the back-to-back stream
// kernels of a step?  (strong scaling at N = 8, where a launch lasts only 0.47 / 0.69 ms.)
//
//   python -c "import __graft_entry__ as g; g.build()"   (-> build/pdl_experiment)
//   build/pdl_experiment [log2_n=27] [pairs=40]
//
// Two add-only kernels with the read:write mixes of the fused forward (3:3) and backward (6:3) run alternately on one
// stream, `pairs` times, exactly like bench.py's timed loop.  Variant A launches them normally; variant B launches every
// kernel with cudaLaunchAttributeProgrammaticStreamSerialization: each CTA executes `griddepcontrol.launch_dependents`
// right after it has issued its loads (so the next grid's CTAs become resident while the current grid drains) and
// `griddepcontrol.wait` before its first global access (the next kernel reads what this one writes, and overwrites what
// this one reads, so nothing may be touched earlier).  One JSON line per variant: ms per forward+backward pair.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "adam_kernels.cuh"

using b200adam::VecIO;

struct Ptrs {
  const float *in[6];
  float *out[3];
};

template <int R, int W, int BLOCK, bool PDL>
__global__ void __launch_bounds__(BLOCK) mix_kernel(const __grid_constant__ Ptrs p, long long n) {
  constexpr int P = 8;
  const long long off = ((long long)blockIdx.x * BLOCK + threadIdx.x) * P;
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (off + P > n) return;
  uint32_t w[R][8];
#pragma unroll
  for (int a = 0; a < R; ++a) VecIO<32, false>::ld(w[a], p.in[a] + off);
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = 1.0f;
#pragma unroll
    for (int a = 0; a < R; ++a) s = __fadd_rn(s, __uint_as_float(w[a][k]));
    acc[k] = __float_as_uint(s);
  }
#pragma unroll
  for (int o = 0; o < W; ++o) VecIO<32, false>::st(p.out[o] + off, acc);
}

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

template <int R, int W, bool PDL>
static void launch(const Ptrs &p, long long n, cudaStream_t s) {
  constexpr int BLOCK = 512;
  const unsigned grid = (unsigned)((n + BLOCK * 8 - 1) / (BLOCK * 8));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(BLOCK), cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = PDL ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, mix_kernel<R, W, BLOCK, PDL>, p, n));
}

template <bool PDL>
static double run(const Ptrs &fwd, const Ptrs &bwd, long long n, int pairs, cudaStream_t s) {
  for (int i = 0; i < 3; ++i) launch<3, 3, PDL>(fwd, n, s), launch<6, 3, PDL>(bwd, n, s);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < pairs; ++i) launch<3, 3, PDL>(fwd, n, s), launch<6, 3, PDL>(bwd, n, s);
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / pairs;
}

int main(int argc, char **argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 27;
  const int pairs = argc > 2 ? atoi(argv[2]) : 40;
  const long long n = 1LL << lg;
  // forward: g, mu, nu -> mu', nu', u ; backward: du, u, mu', g, dme, dne -> dg, dmu, dnu (u, mu' chain the two)
  std::vector<float *> a(12);
  for (auto &d : a) {
    CK(cudaMalloc(&d, (size_t)n * 4));
    CK(cudaMemset(d, 0x3c, (size_t)n * 4));
  }
  Ptrs fwd{}, bwd{};
  fwd.in[0] = a[0], fwd.in[1] = a[1], fwd.in[2] = a[2], fwd.out[0] = a[3], fwd.out[1] = a[4], fwd.out[2] = a[5];
  bwd.in[0] = a[6], bwd.in[1] = a[5], bwd.in[2] = a[3], bwd.in[3] = a[0], bwd.in[4] = a[7], bwd.in[5] = a[8];
  bwd.out[0] = a[9], bwd.out[1] = a[10], bwd.out[2] = a[11];
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (int rep = 0; rep < 2; ++rep) {
    const double plain = run<false>(fwd, bwd, n, pairs, s);
    const double pdl = run<true>(fwd, bwd, n, pairs, s);
    printf("{\"n\": %lld, \"plain_ms_per_pair\": %.4f, \"pdl_ms_per_pair\": %.4f, \"pdl_gain_pct\": %.2f, "
           "\"plain_GBps\": %.0f, \"pdl_GBps\": %.0f}\n",
           n, plain, pdl, 100.0 * (plain - pdl) / plain, 60.0 * n / (plain * 1e-3) / 1e9, 60.0 * n / (pdl * 1e-3) / 1e9);
  }
  for (auto d : a) cudaFree(d);
  return 0;
}
