// tma_experiment.cu -- does staging the streams through shared memory with the TMA bulk-copy engine beat direct
// 256-bit global loads/stores for the fused Adam forward?  (Design question recorded in DESIGN.md section 4.)
//
// Variant under test: persistent CTAs, S-stage ring in shared memory.  One elected thread issues
// cp.async.bulk (global -> shared, mbarrier complete_tx) for the g / mu / nu tiles of a stage; all threads wait on the
// stage's mbarrier, compute the Adam forward IN PLACE in shared memory (same arithmetic helpers as the product kernels,
// so the result must be bit-identical), fence.proxy.async, and the elected thread issues cp.async.bulk (shared -> global,
// bulk_group) for mu', nu', u.  A stage is refilled once its stores have finished reading shared memory
// (cp.async.bulk.wait_group.read).  SASS: UBLKCP (both directions), SYNCS.* for the mbarriers.
//
// The binary times this variant against the product kernel (b200_adam_fused_forward from libb200adam.so) on the same
// arrays and checks that both produce identical bits.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -Iinclude -Itorchopt_b200/csrc \
//        tools/tma_experiment.cu -Ltorchopt_b200 -lb200adam -Xlinker -rpath -Xlinker '$ORIGIN/../torchopt_b200' -o build/tma_experiment
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "adam_kernels.cuh"
#include "b200_adam.h"

using b200adam::Math;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      std::exit(2);                                                                        \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Scal {
  float B1, OMB1, B2, OMB2, ER, EPS, C1, C2;
};

// TILE floats per array per stage; STAGES ring slots; each slot holds g | mu | nu (3 * TILE floats).
// LAG = number of store groups allowed to be still reading shared memory when a slot is refilled: the prefetch depth
// is STAGES - 1 - LAG tiles.
template <int TILE, int STAGES, int THREADS, int LAG>
__global__ void __launch_bounds__(THREADS) adam_fwd_tma(const float *__restrict__ g, const float *__restrict__ mu,
                                                         const float *__restrict__ nu, float *__restrict__ mu_o,
                                                         float *__restrict__ nu_o, float *__restrict__ u_o,
                                                         long long n_tiles, Scal sc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *buf = reinterpret_cast<float *>(smem_raw);                                   // STAGES * 3 * TILE floats
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * 3 * TILE * sizeof(float));
  const int tid = threadIdx.x;
  constexpr uint32_t TILE_BYTES = TILE * sizeof(float);
  using M = Math<float>;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // tiles handled by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

  auto issue_load = [&](long long k) {  // k-th tile of this CTA into slot k % STAGES
    const int s = (int)(k % STAGES);
    const long long off = (first + k * stride) * TILE;
    float *slot = buf + (size_t)s * 3 * TILE;
    mbar_expect_tx(&full[s], 3 * TILE_BYTES);
    bulk_g2s(slot, g + off, TILE_BYTES, &full[s]);
    bulk_g2s(slot + TILE, mu + off, TILE_BYTES, &full[s]);
    bulk_g2s(slot + 2 * TILE, nu + off, TILE_BYTES, &full[s]);
  };

  constexpr int DEPTH = STAGES - 1 - LAG;   // tiles in flight ahead of the one being computed
  if (tid == 0) {
    for (long long k = 0; k < DEPTH && k < my_tiles; ++k) issue_load(k);
  }

  for (long long k = 0; k < my_tiles; ++k) {
    const int s = (int)(k % STAGES);
    const uint32_t parity = (uint32_t)((k / STAGES) & 1);
    // refill the slot that tile k+STAGES-1 will use: it was last used by tile k-1, whose stores were committed at the
    // end of the previous iteration; allow only that newest group... wait until it has finished READING shared memory
    if (tid == 0 && k + DEPTH < my_tiles) {
      // tile k+DEPTH goes into the slot last used by tile k+DEPTH-STAGES = k-1-LAG: all but the newest LAG store
      // groups must have finished reading shared memory
      bulk_wait_read<LAG>();
      issue_load(k + DEPTH);
    }
    mbar_wait(&full[s], parity);
    float *slot = buf + (size_t)s * 3 * TILE;
    float4 *g4 = reinterpret_cast<float4 *>(slot), *m4 = reinterpret_cast<float4 *>(slot + TILE),
           *v4 = reinterpret_cast<float4 *>(slot + 2 * TILE);
#pragma unroll
    for (int i = tid; i < TILE / 4; i += THREADS) {
      float4 gg = g4[i], mm = m4[i], vv = v4[i];
      float *pg = reinterpret_cast<float *>(&gg), *pm = reinterpret_cast<float *>(&mm), *pv = reinterpret_cast<float *>(&vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x = pg[e];
        const float m_out = M::add(M::mul(sc.B1, pm[e]), M::mul(sc.OMB1, x));
        const float v_out = M::add(M::mul(sc.B2, pv[e]), M::mul(M::mul(sc.OMB2, x), x));
        const float u = M::div(M::mul(m_out, sc.C1), M::add(M::sqrt(M::add(M::mul(v_out, sc.C2), sc.ER)), sc.EPS));
        pm[e] = m_out, pv[e] = v_out, pg[e] = u;
      }
      g4[i] = gg, m4[i] = mm, v4[i] = vv;
    }
    fence_async_smem();   // make the generic-proxy writes visible to the async proxy (TMA store)
    __syncthreads();
    if (tid == 0) {
      const long long off = (first + k * stride) * TILE;
      bulk_s2g(mu_o + off, slot + TILE, TILE_BYTES);
      bulk_s2g(nu_o + off, slot + 2 * TILE, TILE_BYTES);
      bulk_s2g(u_o + off, slot, TILE_BYTES);
      bulk_commit();
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int TILE, int STAGES, int THREADS, int LAG>
static float run_tma(const float *g, const float *mu, const float *nu, float *mu_o, float *nu_o, float *u_o, long long n,
                     Scal sc, int ctas_per_sm, int sms, int iters) {
  const size_t smem = (size_t)STAGES * 3 * TILE * sizeof(float) + STAGES * sizeof(uint64_t);
  auto kern = adam_fwd_tma<TILE, STAGES, THREADS, LAG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = n / TILE;
  const int grid = (int)std::min<long long>(n_tiles, (long long)sms * ctas_per_sm);
  for (int i = 0; i < 3; ++i) kern<<<grid, THREADS, smem>>>(g, mu, nu, mu_o, nu_o, u_o, n_tiles, sc);
  CK(cudaGetLastError());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) kern<<<grid, THREADS, smem>>>(g, mu, nu, mu_o, nu_o, u_o, n_tiles, sc);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / iters;
}

int main(int argc, char **argv) {
  const int lg = argc > 1 ? std::atoi(argv[1]) : 28;
  const long long n = 1ll << lg;
  const int iters = 20;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float *g, *mu, *nu, *o[6];
  CK(cudaMalloc(&g, n * 4));
  CK(cudaMalloc(&mu, n * 4));
  CK(cudaMalloc(&nu, n * 4));
  for (auto &p : o) CK(cudaMalloc(&p, n * 4));
  {  // deterministic host init of a 16 Mi-element pattern, replicated
    const long long m = 1ll << 24;
    std::vector<float> hg(m), hm(m), hv(m);
    uint32_t x = 12345u;
    auto rnd = [&]() { x = x * 1664525u + 1013904223u; return (float)((x >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (long long i = 0; i < m; ++i) hg[i] = rnd() * 2e-2f, hm[i] = rnd() * 2e-2f, hv[i] = std::fabs(rnd()) * 2e-4f;
    for (long long i = 0; i < m; i += 1024) hg[i] = 0.f, hm[i] = 0.f;
    for (long long off = 0; off < n; off += m) {
      const long long len = std::min(m, n - off);
      CK(cudaMemcpy(g + off, hg.data(), len * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(mu + off, hm.data(), len * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(nu + off, hv.data(), len * 4, cudaMemcpyHostToDevice));
    }
  }
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8, er = 1e-8;
  const uint64_t count = 3;
  Scal sc;
  sc.B1 = (float)b1, sc.OMB1 = 1.f - sc.B1, sc.B2 = (float)b2, sc.OMB2 = 1.f - sc.B2, sc.ER = (float)er, sc.EPS = (float)eps;
  sc.C1 = (float)(1 / (1 - std::pow(b1, (double)count))), sc.C2 = (float)(1 / (1 - std::pow(b2, (double)count)));

  // product kernel
  const void *pg = g, *pm = mu, *pv = nu;
  void *q0 = o[0], *q1 = o[1], *q2 = o[2];
  const int64_t nn = n;
  for (int i = 0; i < 3; ++i)
    if (b200_adam_fused_forward(B200_ADAM_F32, 1, &pg, &pm, &pv, &q0, &q1, &q2, &nn, &count, b1, b2, eps, er, nullptr)) return 3;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i)
    b200_adam_fused_forward(B200_ADAM_F32, 1, &pg, &pm, &pv, &q0, &q1, &q2, &nn, &count, b1, b2, eps, er, nullptr);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms_prod;
  CK(cudaEventElapsedTime(&ms_prod, e0, e1));
  ms_prod /= iters;
  std::printf("{\"variant\": \"product: direct LDG.256/STG.256, one CTA per chunk\", \"n_log2\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", lg,
              ms_prod, 24.0 * n / (ms_prod * 1e-3) / 1e9);

  auto report = [&](const char *name, float ms) {
    // bit-compare against the product kernel's outputs
    std::vector<uint32_t> a(1 << 20), b(1 << 20);
    long long bad = 0;
    for (int k = 0; k < 3; ++k)
      for (long long off : {0ll, n / 2, n - (1ll << 20)}) {
        CK(cudaMemcpy(a.data(), o[k] + off, a.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), o[3 + k] + off, b.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b[i];
      }
    std::printf("{\"variant\": \"%s\", \"n_log2\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"mismatching_words_in_sample\": %lld}\n", name, lg,
                ms, 24.0 * n / (ms * 1e-3) / 1e9, bad);
  };
#define RUN(TILE, STAGES, THREADS, CPS, LAG)                                                                      \
  {                                                                                                               \
    for (int k = 3; k < 6; ++k) CK(cudaMemset(o[k], 0, n * 4));                                                   \
    float ms = run_tma<TILE, STAGES, THREADS, LAG>(g, mu, nu, o[3], o[4], o[5], n, sc, CPS, sms, iters);          \
    report("tma: tile " #TILE " floats/array, " #STAGES " stages, " #THREADS " threads, " #CPS " CTA/SM, store lag " #LAG, ms); \
  }
  RUN(2048, 4, 128, 2, 0)
  RUN(4096, 3, 512, 1, 0)
  RUN(4096, 4, 512, 1, 0)
  RUN(4096, 4, 512, 1, 1)
  RUN(4096, 4, 1024, 1, 1)
  RUN(2048, 4, 512, 2, 0)
  RUN(2048, 4, 512, 2, 1)
  RUN(2048, 4, 1024, 2, 1)
  RUN(2048, 8, 512, 1, 1)
  RUN(2048, 8, 1024, 1, 2)
  RUN(1024, 8, 512, 2, 1)
  RUN(1024, 4, 256, 4, 1)
  return 0;
}
