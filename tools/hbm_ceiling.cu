// hbm_ceiling.cu -- what does B200 HBM deliver to a TRIVIAL stream kernel with the same read:write mix as the
// Adam kernels?  ("speed of light" evidence.)
//
//   python -c "import __graft_entry__ as g; g.build()"    (-> build/hbm_ceiling)
//   build/hbm_ceiling [log2_n=28] [iters=20]
//
// Each kernel reads R fp32 arrays and writes W fp32 arrays of n elements with the engine's access pattern
// (256-bit ld.global.nc.L1::no_allocate / st.global.L1::no_allocate, one CTA of BLOCK threads per chunk of BLOCK*8*U
// elements, U front-batched vector loads per array) and does one add per element -- no division, no sqrt, no fp64.
// If the Adam kernels reach the same GB/s as the matching mix here, their arithmetic is fully hidden and the
// remaining gap to the nominal 8 TB/s is the memory system's (DRAM read/write turnaround, refresh), not the kernel's.
// Mixes: 1:0 (pure read), 0:1 (pure write), 1:1 (copy), 3:3 (fused forward), 4:3 (last-step backward / whole-step
// forward), 6:3 (fused backward).  One JSON line per mix x launch shape.
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

template <int R, int W, int U, int BLOCK>
__global__ void __launch_bounds__(BLOCK) mix_kernel(const __grid_constant__ Ptrs p, long long n) {
  constexpr int P = 8;
  const long long base = ((long long)blockIdx.x * BLOCK * U + threadIdx.x) * P;
  uint32_t w[R > 0 ? R : 1][U][8];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long off = base + (long long)u * BLOCK * P;
      if (off + P <= n) VecIO<32, true>::ld(w[a][u], p.in[a] + off);
    }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long off = base + (long long)u * BLOCK * P;
    if (off + P > n) continue;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = 1.0f;
#pragma unroll
      for (int a = 0; a < R; ++a) s = __fadd_rn(s, __uint_as_float(w[a][u][k]));
      acc[k] = __float_as_uint(s);
    }
    if (W == 0) {
      // keep the loads alive without writing: the condition is never true for the generated data
      bool hit = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) hit |= (acc[k] == 0x7fc12345u);
      if (hit) VecIO<32, true>::st(p.out[0] + off, acc);
    }
#pragma unroll
    for (int o = 0; o < W; ++o) {
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = acc[k] + (uint32_t)o;
      VecIO<32, true>::st(p.out[o] + off, v);
    }
  }
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

template <int R, int W, int U, int BLOCK>
static void run(const Ptrs &p, long long n, int iters) {
  const long long chunk = (long long)BLOCK * 8 * U;
  const unsigned grid = (unsigned)((n + chunk - 1) / chunk);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) mix_kernel<R, W, U, BLOCK><<<grid, BLOCK>>>(p, n);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) mix_kernel<R, W, U, BLOCK><<<grid, BLOCK>>>(p, n);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  const double gbs = (double)(R + W) * 4.0 * (double)n / (ms * 1e-3) / 1e9;
  printf("{\"reads\": %d, \"writes\": %d, \"unroll\": %d, \"block\": %d, \"n\": %lld, \"ms\": %.4f, \"GBps\": %.1f}\n", R, W,
         U, BLOCK, n, ms, gbs);
  fflush(stdout);
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
}

template <int R, int W>
static void run_shapes(const Ptrs &p, long long n, int iters) {
  run<R, W, 1, 512>(p, n, iters);
  run<R, W, 2, 512>(p, n, iters);
  run<R, W, 1, 256>(p, n, iters);
  run<R, W, 2, 256>(p, n, iters);
}

int main(int argc, char **argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 28;
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  const long long n = 1LL << lg;
  Ptrs p{};
  std::vector<void *> all;
  for (int a = 0; a < 9; ++a) {
    float *d = nullptr;
    CK(cudaMalloc(&d, (size_t)n * 4));
    CK(cudaMemset(d, a < 6 ? 0x3c : 0, (size_t)n * 4));  // finite small floats
    all.push_back(d);
    if (a < 6) p.in[a] = d; else p.out[a - 6] = d;
  }
  CK(cudaDeviceSynchronize());
  run_shapes<1, 0>(p, n, iters);
  run_shapes<0, 1>(p, n, iters);
  run_shapes<1, 1>(p, n, iters);
  run_shapes<2, 1>(p, n, iters);
  run_shapes<3, 3>(p, n, iters);
  run_shapes<4, 3>(p, n, iters);
  run_shapes<6, 3>(p, n, iters);
  // cudaMemcpy device-to-device for reference (what torch.copy_ / MEASURED_PEAKS.json times)
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(p.out[0], p.in[0], (size_t)n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(cudaMemcpyAsync(p.out[0], p.in[0], (size_t)n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("{\"reads\": 1, \"writes\": 1, \"kind\": \"cudaMemcpyD2D\", \"n\": %lld, \"ms\": %.4f, \"GBps\": %.1f}\n", n, ms,
           8.0 * (double)n / (ms * 1e-3) / 1e9);
  }
  for (void *d : all) cudaFree(d);
  return 0;
}
