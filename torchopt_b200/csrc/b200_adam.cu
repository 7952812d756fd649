// b200_adam.cu -- host side of libb200adam.so: the C ABI declared in include/b200_adam.h.
//
// Each entry point derives the host scalars exactly like the reference launchers do
// (/root/reference/src/adam_op/adam_op_impl_cuda.cu:73-75,253-255,452-454 == adam_op_impl_cpu.cpp:72-74,
// 190-192,327-329: std::pow in double, one rounding to the tensor dtype), fills a leaf table that travels in
// the kernel parameters and launches ONE streaming kernel (adam_kernels.cuh) on the caller's stream.
#include "b200_adam.h"

#include <atomic>
#include <cmath>
#include <type_traits>
#include <vector>

#include "adam_kernels.cuh"

namespace b200adam {

// ---- launch configuration (chosen from the on-device sweep recorded in profiles/) ------------------
// vector width P (elements), unroll U (vectors in flight per array per thread), CTA size.
#ifndef B200_ADAM_BLOCK
#define B200_ADAM_BLOCK 512
#endif
template <typename TG, typename TS>
struct Cfg;  // fwd / bwd / unfused vector configs per dtype family
// fp32 defaults can be overridden at compile time by the tuning sweep (tools/tune_sweep.py)
#ifndef B200_P_FWD
#define B200_P_FWD 8
#endif
#ifndef B200_U_FWD
#define B200_U_FWD 1
#endif
#ifndef B200_P_BWD
#define B200_P_BWD 8
#endif
#ifndef B200_U_BWD
#define B200_U_BWD 2
#endif
#ifndef B200_STEP_U_BWD
#define B200_STEP_U_BWD 1
#endif
#ifndef B200_STEP_BLOCK_BWD
#define B200_STEP_BLOCK_BWD 128
#endif
template <>
struct Cfg<float, float> {
  static constexpr int P_FWD = B200_P_FWD, U_FWD = B200_U_FWD, P_BWD = B200_P_BWD, U_BWD = B200_U_BWD, P_UNF = 8,
                       U_UNF = 2;
};
template <>
struct Cfg<double, double> {
  static constexpr int P_FWD = 4, U_FWD = 1, P_BWD = 4, U_BWD = 1, P_UNF = 4, U_UNF = 2;
};
template <>
struct Cfg<bf16_t, float> {
  static constexpr int P_FWD = 8, U_FWD = 2, P_BWD = 8, U_BWD = 1, P_UNF = 8, U_UNF = 2;
};

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_ctas_per_sm{0};
static std::atomic<int> g_tiles_per_chunk{0};
static std::atomic<int> g_pdl{1};  // programmatic dependent launch on every launch (b200_adam_set_pdl)

struct DeviceInfo {
  int sms = 0;
  int cc_major = 0;
};

static int device_info(DeviceInfo *out) {
  static DeviceInfo cache[64];
  static std::atomic<int> ready[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return B200_ADAM_E_NO_DEVICE;
  if (!ready[dev].load(std::memory_order_acquire)) {
    DeviceInfo d;
    e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return (int)e;
    cache[dev] = d;
    ready[dev].store(1, std::memory_order_release);
  }
  *out = cache[dev];
  // the library carries sm_100a SASS only: fail with a clear code instead of "no kernel image" on anything else
  if (out->cc_major != 10) return B200_ADAM_E_NO_DEVICE;
  return 0;
}

// One host-side leaf handed to the generic launcher.
template <class Op>
struct HostLeaf {
  void *ptr[Op::NPTR];
  long long n;
  typename Op::CT ls[Op::NLS > 0 ? Op::NLS : 1];
};

template <class Op, int P, int U, int BLOCK>
static int pick_chunk_elems(long long total_elems, int sms) {
  constexpr int TILE = BLOCK * P * U;
  int tiles = g_tiles_per_chunk.load(std::memory_order_relaxed);
  if (tiles <= 0) tiles = 1;  // measured best (profiles/r01_tune_sweep.md): one block-tile per chunk
  // small problems: one tile per chunk so the work spreads over as many SMs as possible
  if (total_elems < (long long)TILE * tiles * sms * 4) tiles = 1;
  return TILE * tiles;
}

template <class Op, int P, int U, int BLOCK, int MAXL>
static int launch_table(Table<Op, MAXL> &tab, const DeviceInfo &di, cudaStream_t stream) {
  static_assert(sizeof(Table<Op, MAXL>) <= 32764, "the leaf table must fit the 32 KB kernel-parameter space of sm_100");
  if (tab.total_chunks <= 0) return 0;
  // default: one CTA per chunk (non-persistent).  ctas_per_sm > 0 selects a persistent grid of that many CTAs/SM.
  const int per_sm = g_ctas_per_sm.load(std::memory_order_relaxed);
  long long grid = tab.total_chunks;
  if (per_sm > 0 && (long long)di.sms * per_sm < grid) grid = (long long)di.sms * per_sm;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(BLOCK), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel<Op, P, U, BLOCK, MAXL>, tab);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

// Feeds leaves (produced by `next(i, leaf)`, which returns false for a leaf to skip) through tables of
// capacity MAXL, launching whenever a table is full.
template <class Op, int P, int U, int BLOCK, int MAXL, class NextFn>
static int run_batches(long long num_leaves, long long total_elems, int flags, const typename Op::CT *gs,
                       NextFn &&next, cudaStream_t stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  Table<Op, MAXL> tab;
  const int chunk = pick_chunk_elems<Op, P, U, BLOCK>(total_elems, di.sms);
  auto reset = [&]() {
    tab.num_leaves = 0;
    tab.chunk_elems = chunk;
    tab.total_chunks = 0;
    tab.flags = flags;
    for (int k = 0; k < Op::NGS; ++k) tab.gs[k] = gs[k];
    tab.chunk_start[0] = 0;
  };
  reset();
  for (long long i = 0; i < num_leaves; ++i) {
    HostLeaf<Op> hl{};  // optional pointers (capturable count / tables) default to NULL
    if (!next(i, hl)) continue;
    long long off = 0;
    // a leaf with more than INT_MAX/2 chunks is split into several table entries (keeps chunk ids in int)
    while (off < hl.n) {
      const long long max_elems = (long long)chunk * (1ll << 28);
      const long long part = (hl.n - off < max_elems) ? (hl.n - off) : max_elems;
      const long long nchunks = (part + chunk - 1) / chunk;
      if (tab.num_leaves == MAXL || (long long)tab.total_chunks + nchunks > (1ll << 30)) {
        if (int rc = launch_table<Op, P, U, BLOCK, MAXL>(tab, di, stream)) return rc;
        reset();
      }
      const int l = tab.num_leaves++;
      tab.n[l] = part;
      for (int a = 0; a < Op::NPTR; ++a) tab.ptr[a][l] = hl.ptr[a];
      for (int a = 0; a < Op::NLS; ++a) tab.ls[a][l] = hl.ls[a];
      tab.total_chunks += (int)nchunks;
      tab.chunk_start[l + 1] = tab.total_chunks;
      off += part;
      if (off < hl.n) {  // advance the pointers of the split leaf; element sizes are Op specific
        Op::advance(hl.ptr, part);
      }
    }
  }
  return launch_table<Op, P, U, BLOCK, MAXL>(tab, di, stream);
}

// capacity of the largest leaf table: 256 unless the operator declares a smaller MAXL_BIG (32 KB parameter limit)
template <class Op, class = void>
struct MaxLeaves {
  static constexpr int value = 256;
};
template <class Op>
struct MaxLeaves<Op, std::void_t<decltype(Op::MAXL_BIG)>> {
  static constexpr int value = Op::MAXL_BIG;
};
template <class Op>
constexpr int max_leaves() {
  return MaxLeaves<Op>::value;
}

template <class Op, int P, int U, int BLOCK = B200_ADAM_BLOCK, class NextFn>
static int run_op(long long num_leaves, long long total_elems, int flags, const typename Op::CT *gs, NextFn &&next,
                  cudaStream_t stream) {
  if (num_leaves <= 1) return run_batches<Op, P, U, BLOCK, 1>(num_leaves, total_elems, flags, gs, next, stream);
  if (num_leaves <= 64) return run_batches<Op, P, U, BLOCK, 64>(num_leaves, total_elems, flags, gs, next, stream);
  return run_batches<Op, P, U, BLOCK, max_leaves<Op>()>(num_leaves, total_elems, flags, gs, next, stream);
}

// ---- size-dependent launch shape (profiles/r02_tune_mid.md) ------------------------------------------
// The shapes tuned at 2^28 elements (r01_tune_sweep.md) are not the best for mid-size launches, where a launch is only
// 3-30 waves of CTAs and the last, partially filled wave is a visible share of it:
//   * forward: below 2^23 elements 256-thread CTAs (finer last wave) beat 512-thread CTAs by 12 % at 2^22 and tie above;
//   * backward: the U=2 shape (two vectors in flight per array per thread, 8192 elements per CTA, one CTA per SM) loses
//     20 % at 2^22 and 3 % at 2^24 against U=1 and only pays off (+0-1 %) from 2^26 on.
constexpr long long kFwdSmallBlockBelow = 1ll << 23;
constexpr long long kBwdSingleVectorBelow = 1ll << 26;

template <class Op, int P, int U, class NextFn>
static int run_fwd_sized(long long num_leaves, long long total, int flags, const typename Op::CT *gs, NextFn &&next,
                         cudaStream_t stream) {
  if (B200_ADAM_BLOCK > 256 && total < kFwdSmallBlockBelow)
    return run_op<Op, P, U, 256>(num_leaves, total, flags, gs, next, stream);
  return run_op<Op, P, U>(num_leaves, total, flags, gs, next, stream);
}

template <class Op, int P, int U, class NextFn>
static int run_bwd_sized(long long num_leaves, long long total, int flags, const typename Op::CT *gs, NextFn &&next,
                         cudaStream_t stream) {
  if constexpr (U > 1) {
    if (total < kBwdSingleVectorBelow) return run_op<Op, P, 1>(num_leaves, total, flags, gs, next, stream);
  }
  return run_op<Op, P, U>(num_leaves, total, flags, gs, next, stream);
}

// ---- host wrappers that give every Op an `advance` (pointer bump for >2^28-chunk leaves) ------------
template <typename TG, typename TS, typename CT, bool INPLACE>
struct FwdH : FusedFwd<TG, TS, CT, INPLACE> {
  static void advance(void **p, long long k) {
    const size_t sz[6] = {sizeof(TG), sizeof(TS), sizeof(TS), sizeof(TS), sizeof(TS), sizeof(TG)};
    for (int a = 0; a < 6; ++a)
      if (p[a]) p[a] = (char *)p[a] + (size_t)k * sz[a];
  }
};
template <typename TG, typename TS, typename CT, int MODE>
struct BwdH : FusedBwd<TG, TS, CT, MODE> {
  static void advance(void **p, long long k) {
    const size_t sz[9] = {sizeof(TG), sizeof(TG), sizeof(TS), sizeof(TG), sizeof(TS),
                          sizeof(TS), sizeof(TG), sizeof(TS), sizeof(TS)};
    for (int a = 0; a < 9; ++a)
      if (p[a]) p[a] = (char *)p[a] + (size_t)k * sz[a];
  }
};
template <typename T, int KIND>
struct UnfH : UnfusedOp<T, KIND> {
  static void advance(void **p, long long k) {
    for (int a = 0; a < UnfusedOp<T, KIND>::NPTR; ++a)
      if (p[a]) p[a] = (char *)p[a] + (size_t)k * sizeof(T);
  }
};

// ---- host scalars (the reference's derivation; see file header) -----------------------------------
static inline double inv_one_minus_pow(double b, uint64_t count) { return 1 / (1 - std::pow(b, (double)count)); }
static inline double one_minus_pow(double b, uint64_t count) { return 1 - std::pow(b, (double)count); }
static inline double inv_one_minus_pow_eps(double b, double eps_root, uint64_t count) {
  return 1 / (1 - std::pow(b, (double)count) + eps_root);
}

// ---- fused forward ---------------------------------------------------------------------------------
template <typename TG, typename TS, typename CT, bool INPLACE>
static int fused_forward_impl(int64_t L, const void *const *g, const void *const *mu, const void *const *nu,
                              void *const *mu_out, void *const *nu_out, void *const *u_out, const int64_t *n,
                              const uint64_t *count, double b1, double b2, double eps, double eps_root,
                              cudaStream_t stream) {
  using Op = FwdH<TG, TS, CT, INPLACE>;
  const CT B1 = (CT)b1, B2 = (CT)b2;
  const CT gs[6] = {B1, CT(1) - B1, B2, CT(1) - B2, (CT)eps_root, (CT)eps};
  long long total = 0;
  for (int64_t i = 0; i < L; ++i) {
    if (n[i] < 0) return B200_ADAM_E_ARG;
    if (n[i] > 0 && (!g[i] || !mu[i] || !nu[i] || !mu_out[i] || !nu_out[i] || !u_out[i])) return B200_ADAM_E_ARG;
    total += n[i];
  }
  uint64_t last_count = ~0ull;
  CT c1 = 0, c2 = 0;
  auto next = [&](long long i, HostLeaf<Op> &hl) {
    if (n[i] == 0) return false;
    if (count[i] != last_count) {  // pow() once per distinct count
      last_count = count[i];
      c1 = (CT)inv_one_minus_pow(b1, last_count);
      c2 = (CT)inv_one_minus_pow(b2, last_count);
    }
    hl.ptr[0] = const_cast<void *>(g[i]), hl.ptr[1] = const_cast<void *>(mu[i]), hl.ptr[2] = const_cast<void *>(nu[i]);
    hl.ptr[3] = mu_out[i], hl.ptr[4] = nu_out[i], hl.ptr[5] = u_out[i];
    hl.n = n[i];
    hl.ls[0] = c1, hl.ls[1] = c2;
    return true;
  };
  using C = Cfg<TG, TS>;
  return run_fwd_sized<Op, C::P_FWD, C::U_FWD>(L, total, 0, gs, next, stream);
}

template <bool INPLACE>
static int fused_forward_dispatch(int dtype, int64_t L, const void *const *g, const void *const *mu,
                                  const void *const *nu, void *const *mu_out, void *const *nu_out,
                                  void *const *u_out, const int64_t *n, const uint64_t *count, double b1, double b2,
                                  double eps, double eps_root, void *stream) {
  if (L < 0) return B200_ADAM_E_ARG;
  if (L == 0) return 0;
  if (!g || !mu || !nu || !mu_out || !nu_out || !u_out || !n || !count) return B200_ADAM_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case B200_ADAM_F32:
      return fused_forward_impl<float, float, float, INPLACE>(L, g, mu, nu, mu_out, nu_out, u_out, n, count, b1, b2,
                                                              eps, eps_root, s);
#ifndef B200_ADAM_TUNE_F32_ONLY
    case B200_ADAM_F64:
      return fused_forward_impl<double, double, double, INPLACE>(L, g, mu, nu, mu_out, nu_out, u_out, n, count, b1,
                                                                 b2, eps, eps_root, s);
    case B200_ADAM_BF16_F32:
      return fused_forward_impl<bf16_t, float, float, INPLACE>(L, g, mu, nu, mu_out, nu_out, u_out, n, count, b1, b2,
                                                               eps, eps_root, s);
#endif
    default:
      return B200_ADAM_E_DTYPE;
  }
}

// ---- fused backward --------------------------------------------------------------------------------
template <typename TG, typename TS, typename CT, int MODE>
static int fused_backward_mode(int64_t L, const void *const *du, const void *const *u, const void *const *new_mu,
                               const void *const *g, const void *const *dmu_ext, const void *const *dnu_ext,
                               void *const *dg, void *const *dmu, void *const *dnu, const int64_t *n,
                               const uint64_t *count, double b1, double b2, double eps_root, long long total,
                               cudaStream_t stream) {
  using Op = BwdH<TG, TS, CT, MODE>;
  const CT B1 = (CT)b1, B2 = (CT)b2;
  const CT gs[4] = {B1, CT(1) - B1, B2, CT(2) * (CT(1) - B2)};
  uint64_t last_count = ~0ull;
  CT o1 = 0, c2e = 0;
  auto next = [&](long long i, HostLeaf<Op> &hl) {
    if (n[i] == 0) return false;
    if (!dg[i] && !dmu[i] && !dnu[i]) return false;  // nothing to write
    if (count[i] != last_count) {
      last_count = count[i];
      o1 = (CT)one_minus_pow(b1, last_count);
      c2e = (CT)inv_one_minus_pow_eps(b2, eps_root, last_count);
    }
    const void *p[9] = {du[i], u[i], new_mu[i], g[i], dmu_ext[i], dnu_ext[i], dg[i], dmu[i], dnu[i]};
    for (int a = 0; a < 9; ++a) hl.ptr[a] = const_cast<void *>(p[a]);
    hl.n = n[i];
    hl.ls[0] = o1, hl.ls[1] = c2e;
    return true;
  };
  using C = Cfg<TG, TS>;
  if constexpr (MODE == BWD_GENERIC) {
    // run-time presence flags cost registers: half-width vectors, no unroll, 128-thread CTAs (spill-free)
    return run_op<Op, C::P_BWD / 2, 1, 128>(L, total, 0, gs, next, stream);
  } else {
    return run_bwd_sized<Op, C::P_BWD, C::U_BWD>(L, total, 0, gs, next, stream);
  }
}

template <typename TG, typename TS, typename CT>
static int fused_backward_impl(int64_t L, const void *const *du, const void *const *u, const void *const *new_mu,
                               const void *const *g, const void *const *dmu_ext, const void *const *dnu_ext,
                               void *const *dg, void *const *dmu, void *const *dnu, const int64_t *n,
                               const uint64_t *count, double b1, double b2, double eps_root, cudaStream_t stream) {
  bool all_full = true, all_last = true;
  long long total = 0;
  for (int64_t i = 0; i < L; ++i) {
    if (n[i] < 0) return B200_ADAM_E_ARG;
    if (n[i] == 0) continue;
    total += n[i];
    const bool hdu = du[i] != nullptr, hme = dmu_ext[i] != nullptr, hne = dnu_ext[i] != nullptr;
    if (hdu && (!u[i] || !new_mu[i])) return B200_ADAM_E_ARG;
    if (dg[i] && (hdu || hne) && !g[i]) return B200_ADAM_E_ARG;
    const bool outs = dg[i] && dmu[i] && dnu[i];
    all_full = all_full && hdu && hme && hne && outs;
    all_last = all_last && hdu && !hme && !hne && outs;
  }
  if (total == 0) return 0;
  if (all_full)
    return fused_backward_mode<TG, TS, CT, BWD_FULL>(L, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n, count,
                                                     b1, b2, eps_root, total, stream);
  if (all_last)
    return fused_backward_mode<TG, TS, CT, BWD_LAST>(L, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n, count,
                                                     b1, b2, eps_root, total, stream);
  return fused_backward_mode<TG, TS, CT, BWD_GENERIC>(L, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n, count,
                                                      b1, b2, eps_root, total, stream);
}

// ---- whole-step kernels (Adam + scale(step_size) + apply_updates) ----------------------------------
template <typename TG, typename TS, typename CT, bool INPLACE>
struct StepFwdH : StepFwd<TG, TS, CT, INPLACE> {
  static void advance(void **p, long long k) {
    const size_t sz[8] = {sizeof(TG), sizeof(TG), sizeof(TS), sizeof(TS), sizeof(TG), sizeof(TS), sizeof(TS), sizeof(TG)};
    for (int a = 0; a < 8; ++a)
      if (p[a]) p[a] = (char *)p[a] + (size_t)k * sz[a];
  }
};
template <typename TG, typename TS, typename CT, int MODE>
struct StepBwdH : StepBwd<TG, TS, CT, MODE> {
  static void advance(void **p, long long k) {
    const size_t sz[12] = {sizeof(TG), sizeof(TG), sizeof(TS), sizeof(TS), sizeof(TG), sizeof(TS),
                           sizeof(TS), sizeof(TG), sizeof(TS), sizeof(TS), sizeof(TG), sizeof(TG)};
    for (int a = 0; a < 12; ++a)
      if (p[a]) p[a] = (char *)p[a] + (size_t)k * sz[a];
  }
};

// Capturable launches: per-leaf DEVICE step counts + device bias tables instead of host counts (count == NULL then).
struct DeviceCounts {
  const int64_t *const *count_dev = nullptr;
  const void *c1 = nullptr, *c2 = nullptr, *o1 = nullptr, *c2e = nullptr;
  int64_t len = 0;
  bool on() const { return count_dev != nullptr; }
};

template <typename TG, typename TS, typename T, bool INPLACE>
static int step_forward_impl(int64_t L, const void *const *p, const void *const *g, const void *const *mu,
                             const void *const *nu, void *const *p_out, void *const *mu_out, void *const *nu_out,
                             void *const *us_out, const int64_t *n, const uint64_t *count, double b1, double b2,
                             double eps, double eps_root, double step_size, double wd_l2, double wd_decoupled,
                             int maximize, cudaStream_t stream, const DeviceCounts &dc = DeviceCounts()) {
  using Op = StepFwdH<TG, TS, T, INPLACE>;
  const T B1 = (T)b1, B2 = (T)b2;
  const T gs[10] = {B1, T(1) - B1, B2, T(1) - B2, (T)eps_root, (T)eps, (T)step_size, (T)wd_l2, (T)wd_decoupled,
                    maximize ? T(-1) : T(1)};
  long long total = 0;
  for (int64_t i = 0; i < L; ++i) {
    if (n[i] < 0) return B200_ADAM_E_ARG;
    if (n[i] > 0 && (!p[i] || !g[i] || !mu[i] || !nu[i] || !p_out[i] || !mu_out[i] || !nu_out[i]))
      return B200_ADAM_E_ARG;
    if (n[i] > 0 && dc.on() && !dc.count_dev[i]) return B200_ADAM_E_ARG;
    total += n[i];
  }
  uint64_t last_count = ~0ull;
  T c1 = 0, c2 = 0;
  auto next = [&](long long i, HostLeaf<Op> &hl) {
    if (n[i] == 0) return false;
    if (dc.on()) {  // corrections come from the device tables, indexed by the device count
      hl.ptr[8] = const_cast<int64_t *>(dc.count_dev[i]);
      hl.ptr[9] = const_cast<void *>(dc.c1), hl.ptr[10] = const_cast<void *>(dc.c2);
    } else if (count[i] != last_count) {
      last_count = count[i];
      c1 = (T)inv_one_minus_pow(b1, last_count);
      c2 = (T)inv_one_minus_pow(b2, last_count);
    }
    const void *in[4] = {p[i], g[i], mu[i], nu[i]};
    for (int a = 0; a < 4; ++a) hl.ptr[a] = const_cast<void *>(in[a]);
    hl.ptr[4] = p_out[i], hl.ptr[5] = mu_out[i], hl.ptr[6] = nu_out[i], hl.ptr[7] = us_out ? us_out[i] : nullptr;
    hl.n = n[i];
    hl.ls[0] = c1, hl.ls[1] = c2;
    return true;
  };
  using C = Cfg<TG, TS>;
  return run_fwd_sized<Op, C::P_FWD, C::U_FWD>(L, total, (int)dc.len, gs, next, stream);
}

template <typename TG, typename TS, typename T>
static int step_capturable_impl(int64_t L, void *const *p, void *const *g, void *const *mu, void *const *nu,
                                const int64_t *n, const int64_t *const *count_dev, const void *c1_table,
                                const void *c2_table, int64_t table_len, double b1, double b2, double eps,
                                double eps_root, double step_size, double wd_l2, double wd_decoupled, int maximize,
                                int overwrite_grads, cudaStream_t stream) {
  DeviceCounts dc;
  dc.count_dev = count_dev, dc.c1 = c1_table, dc.c2 = c2_table, dc.len = table_len;
  return step_forward_impl<TG, TS, T, true>(L, p, g, mu, nu, p, mu, nu, overwrite_grads ? g : nullptr, n, nullptr, b1, b2,
                                            eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, stream, dc);
}

// number of table entries after which 1/(1-b^count) is exactly 1 in T for every larger count (the value decreases
// monotonically to 1); 0 if that does not happen within `cap` counts
template <typename T>
static int64_t bias_table_len_one(double b, int64_t cap) {
  auto is_one = [&](int64_t c) { return (T)inv_one_minus_pow(b, (uint64_t)c) == T(1); };
  if (!is_one(cap)) return 0;
  int64_t lo = 1, hi = cap;  // smallest c with is_one(c)
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (is_one(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// smallest count from which `value(c)` equals `limit` (value is monotone in c and converges to limit); 0 if that does
// not happen within `cap` counts
template <class F>
static int64_t first_count_at_limit(F &&at_limit, int64_t cap) {
  if (!at_limit(cap)) return 0;
  int64_t lo = 1, hi = cap;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (at_limit(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// length of the FOUR bias tables (c1, c2 for the forward; o1, c2e in addition for the backward): the count from which
// every correction has reached its limit value in T
template <typename T>
static int64_t bias_table_len_all(double b1, double b2, double eps_root, int64_t cap) {
  const T c2e_limit = (T)(1 / (1 - 0.0 + eps_root));
  const int64_t a = first_count_at_limit([&](int64_t c) { return (T)inv_one_minus_pow(b1, (uint64_t)c) == T(1); }, cap);
  const int64_t b = first_count_at_limit([&](int64_t c) { return (T)inv_one_minus_pow(b2, (uint64_t)c) == T(1); }, cap);
  const int64_t d = first_count_at_limit([&](int64_t c) { return (T)one_minus_pow(b1, (uint64_t)c) == T(1); }, cap);
  const int64_t e = first_count_at_limit(
      [&](int64_t c) { return (T)inv_one_minus_pow_eps(b2, eps_root, (uint64_t)c) == c2e_limit; }, cap);
  if (a == 0 || b == 0 || d == 0 || e == 0) return 0;
  int64_t m = a > b ? a : b;
  m = d > m ? d : m;
  return e > m ? e : m;
}

template <typename T>
static void fill_bias_tables(double b1, double b2, double eps_root, int64_t len, void *c1, void *c2, void *o1, void *c2e) {
  for (int64_t c = 1; c <= len; ++c) {
    ((T *)c1)[c - 1] = (T)inv_one_minus_pow(b1, (uint64_t)c);
    ((T *)c2)[c - 1] = (T)inv_one_minus_pow(b2, (uint64_t)c);
    if (o1) ((T *)o1)[c - 1] = (T)one_minus_pow(b1, (uint64_t)c);
    if (c2e) ((T *)c2e)[c - 1] = (T)inv_one_minus_pow_eps(b2, eps_root, (uint64_t)c);
  }
}

template <typename TG, typename TS, typename T, int MODE>
static int step_backward_mode(int64_t L, const void *const *dp, const void *const *dus, const void *const *new_mu,
                              const void *const *new_nu, const void *const *g, const void *const *dmu_ext,
                              const void *const *dnu_ext, void *const *dg, void *const *dmu, void *const *dnu,
                              const void *const *p, void *const *dp_out, const int64_t *n, const uint64_t *count,
                              double b1, double b2, double eps, double eps_root, double step_size, double wd_l2,
                              double wd_decoupled, int maximize, long long total, cudaStream_t stream,
                              const DeviceCounts &dc) {
  using Op = StepBwdH<TG, TS, T, MODE>;
  const T B1 = (T)b1, B2 = (T)b2;
  const T gs[10] = {B1, T(1) - B1, B2, T(2) * (T(1) - B2), (T)eps_root, (T)eps, (T)step_size, (T)wd_l2,
                    (T)wd_decoupled, maximize ? T(-1) : T(1)};
  uint64_t last_count = ~0ull;
  T o1 = 0, c2e = 0, c1 = 0, c2 = 0;
  auto next = [&](long long i, HostLeaf<Op> &hl) {
    if (n[i] == 0) return false;
    void *dpo = dp_out ? dp_out[i] : nullptr;
    if (!dg[i] && !dmu[i] && !dnu[i] && !dpo) return false;
    if (dc.on()) {
      hl.ptr[12] = const_cast<int64_t *>(dc.count_dev[i]);
      hl.ptr[13] = const_cast<void *>(dc.c1), hl.ptr[14] = const_cast<void *>(dc.c2);
      hl.ptr[15] = const_cast<void *>(dc.o1), hl.ptr[16] = const_cast<void *>(dc.c2e);
    } else if (count[i] != last_count) {
      last_count = count[i];
      o1 = (T)one_minus_pow(b1, last_count);
      c2e = (T)inv_one_minus_pow_eps(b2, eps_root, last_count);
      c1 = (T)inv_one_minus_pow(b1, last_count);
      c2 = (T)inv_one_minus_pow(b2, last_count);
    }
    const void *q[12] = {dp[i], dus ? dus[i] : nullptr, new_mu[i], new_nu[i], g[i], dmu_ext[i], dnu_ext[i],
                         dg[i], dmu[i], dnu[i], p ? p[i] : nullptr, dpo};
    for (int a = 0; a < 12; ++a) hl.ptr[a] = const_cast<void *>(q[a]);
    hl.n = n[i];
    hl.ls[0] = o1, hl.ls[1] = c2e, hl.ls[2] = c1, hl.ls[3] = c2;
    return true;
  };
  using C = Cfg<TG, TS>;
  if constexpr (MODE == BWD_GENERIC) {
    return run_op<Op, C::P_BWD / 2, 1, 128>(L, total, (int)dc.len, gs, next, stream);
  } else {
    // The step backward recomputes u (one more IEEE div + sqrt) and carries one more stream (p) than FusedBwd, so it
    // is the most instruction-heavy kernel of the family.  Measured (profiles/r01_tune_step.md): U=2 needs 142
    // registers and halves the resident warps (4.4-4.6 TB/s); U=1 with SMALL CTAs is best -- 128 threads x 92
    // registers = 5 CTAs (20 warps) per SM, whose compute phases interleave with each other's loads: 7.13 TB/s.
    return run_op<Op, C::P_BWD, B200_STEP_U_BWD, B200_STEP_BLOCK_BWD>(L, total, (int)dc.len, gs, next, stream);
  }
}

template <typename TG, typename TS, typename T>
static int step_backward_impl(int64_t L, const void *const *dp, const void *const *dus, const void *const *new_mu,
                              const void *const *new_nu, const void *const *g, const void *const *dmu_ext,
                              const void *const *dnu_ext, void *const *dg, void *const *dmu, void *const *dnu,
                              const void *const *p, void *const *dp_out, const int64_t *n, const uint64_t *count,
                              double b1, double b2, double eps, double eps_root, double step_size, double wd_l2,
                              double wd_decoupled, int maximize, cudaStream_t stream,
                              const DeviceCounts &dc = DeviceCounts()) {
  bool all_full = true, all_last = true;
  long long total = 0;
  for (int64_t i = 0; i < L; ++i) {
    if (n[i] < 0) return B200_ADAM_E_ARG;
    if (n[i] == 0) continue;
    total += n[i];
    if (dc.on() && !dc.count_dev[i]) return B200_ADAM_E_ARG;
    const bool hdp = dp[i] != nullptr, hdus = dus && dus[i] != nullptr;
    const bool hme = dmu_ext[i] != nullptr, hne = dnu_ext[i] != nullptr;
    if ((hdp || hdus) && (!new_mu[i] || !new_nu[i])) return B200_ADAM_E_ARG;
    const bool want_dg2 = dg[i] || (wd_l2 != 0.0 && dp_out && dp_out[i]);  // dp_wd = wd_l2 * dg2 needs dg2 too
    if (want_dg2 && (hdp || hdus || hne) && !g[i]) return B200_ADAM_E_ARG;
    if (wd_l2 != 0.0 && g[i] && !(p && p[i])) return B200_ADAM_E_ARG;  // L2 decay: g2 is re-derived from g and p
    const bool outs = dg[i] && dmu[i] && dnu[i];
    all_full = all_full && hdp && !hdus && hme && hne && outs;
    all_last = all_last && hdp && !hdus && !hme && !hne && outs;
  }
  if (total == 0) return 0;
#define B200_STEP_BWD(MODE)                                                                                      \
  step_backward_mode<TG, TS, T, MODE>(L, dp, dus, new_mu, new_nu, g, dmu_ext, dnu_ext, dg, dmu, dnu, p, dp_out, n, count, b1, \
                              b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, total, stream, dc)
  if (all_full) return B200_STEP_BWD(BWD_FULL);
  if (all_last) return B200_STEP_BWD(BWD_LAST);
  return B200_STEP_BWD(BWD_GENERIC);
#undef B200_STEP_BWD
}

// ---- un-fused single-tensor operators --------------------------------------------------------------
template <typename T, int KIND>
static int unfused_impl(const void *in0, const void *in1, const void *in2, void *out0, void *out1, int64_t n,
                        const T (&gs)[4], cudaStream_t stream) {
  using Op = UnfH<T, KIND>;
  if (n < 0) return B200_ADAM_E_ARG;
  if (n == 0) return 0;
  const void *ins[3] = {in0, in1, in2};
  void *outs[2] = {out0, out1};
  for (int a = 0; a < Op::NIN; ++a)
    if (!ins[a]) return B200_ADAM_E_ARG;
  for (int a = 0; a < Op::NOUT; ++a)
    if (!outs[a]) return B200_ADAM_E_ARG;
  auto next = [&](long long, HostLeaf<Op> &hl) {
    for (int a = 0; a < Op::NIN; ++a) hl.ptr[a] = const_cast<void *>(ins[a]);
    for (int a = 0; a < Op::NOUT; ++a) hl.ptr[Op::NIN + a] = outs[a];
    hl.n = n;
    hl.ls[0] = hl.ls[1] = T(0);
    return true;
  };
  using C = Cfg<T, T>;
  return run_op<Op, C::P_UNF, C::U_UNF>(1, n, 0, gs, next, stream);
}

#ifdef B200_ADAM_TUNE_F32_ONLY
#define B200_UNFUSED_DISPATCH(KIND, in0, in1, in2, out0, out1, s0, s1, s2, s3) return B200_ADAM_E_DTYPE;
#else
#define B200_UNFUSED_DISPATCH(KIND, in0, in1, in2, out0, out1, s0, s1, s2, s3)                          \
  switch (dtype) {                                                                                      \
    case B200_ADAM_F32: {                                                                               \
      const float gs[4] = {(float)(s0), (float)(s1), (float)(s2), (float)(s3)};                         \
      return unfused_impl<float, KIND>(in0, in1, in2, out0, out1, n, gs, (cudaStream_t)stream);         \
    }                                                                                                   \
    case B200_ADAM_F64: {                                                                               \
      const double gs[4] = {(double)(s0), (double)(s1), (double)(s2), (double)(s3)};                    \
      return unfused_impl<double, KIND>(in0, in1, in2, out0, out1, n, gs, (cudaStream_t)stream);        \
    }                                                                                                   \
    default:                                                                                            \
      return B200_ADAM_E_DTYPE;                                                                         \
  }
#endif

// `1 - rn_T(b)` must be formed in T from the rounded beta (adam_op_impl_cpu.cpp:105,140,223,261).
template <typename T>
static inline T omb(double b) {
  return T(1) - (T)b;
}

}  // namespace b200adam

using namespace b200adam;

extern "C" {

int b200_adam_abi_version(void) { return B200_ADAM_ABI_VERSION; }

const char *b200_adam_error_string(int code) {
  switch (code) {
    case B200_ADAM_OK: return "ok";
    case B200_ADAM_E_DTYPE: return "b200_adam: dtype not supported by this entry point";
    case B200_ADAM_E_ARG: return "b200_adam: invalid argument (NULL array, negative size or missing saved tensor)";
    case B200_ADAM_E_NO_DEVICE: return "b200_adam: the current CUDA device is not a compute-capability-10.x (B200) device";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "b200_adam: unknown error";
}

uint64_t b200_adam_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int b200_adam_set_tuning(int ctas_per_sm, int tiles_per_chunk) {
  if (ctas_per_sm < 0 || tiles_per_chunk < 0) return B200_ADAM_E_ARG;
  g_ctas_per_sm.store(ctas_per_sm);
  g_tiles_per_chunk.store(tiles_per_chunk);
  return 0;
}

int b200_adam_set_pdl(int enabled) {
  g_pdl.store(enabled ? 1 : 0);
  return 0;
}

int b200_adam_query_launch(int dtype, int backward, int64_t n, int32_t out[5]) {
  if (!out || n < 0) return B200_ADAM_E_ARG;
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  int block = B200_ADAM_BLOCK, P, U;
  switch (dtype) {
    case B200_ADAM_F32: P = backward ? Cfg<float, float>::P_BWD : Cfg<float, float>::P_FWD,
                        U = backward ? Cfg<float, float>::U_BWD : Cfg<float, float>::U_FWD; break;
#ifndef B200_ADAM_TUNE_F32_ONLY
    case B200_ADAM_F64: P = backward ? Cfg<double, double>::P_BWD : Cfg<double, double>::P_FWD,
                        U = backward ? Cfg<double, double>::U_BWD : Cfg<double, double>::U_FWD; break;
    case B200_ADAM_BF16_F32: P = backward ? Cfg<bf16_t, float>::P_BWD : Cfg<bf16_t, float>::P_FWD,
                             U = backward ? Cfg<bf16_t, float>::U_BWD : Cfg<bf16_t, float>::U_FWD; break;
#endif
    default: return B200_ADAM_E_DTYPE;
  }
  // the size-dependent shape of run_fwd_sized / run_bwd_sized
  if (!backward && block > 256 && n < kFwdSmallBlockBelow) block = 256;
  if (backward && U > 1 && n < kBwdSingleVectorBelow) U = 1;
  int tiles = g_tiles_per_chunk.load(std::memory_order_relaxed);
  const long long tile = (long long)block * P * U;
  if (tiles <= 0 || n < tile * tiles * di.sms * 4) tiles = 1;
  const int chunk = (int)(tile * tiles);
  const int per_sm = g_ctas_per_sm.load();
  const long long chunks = (n + chunk - 1) / chunk;
  long long grid = chunks;
  if (per_sm > 0 && (long long)di.sms * per_sm < grid) grid = (long long)di.sms * per_sm;
  out[0] = (int32_t)grid, out[1] = block, out[2] = chunk, out[3] = P, out[4] = U;
  return 0;
}

int b200_adam_fused_forward(int dtype, int64_t num_leaves, const void *const *g, const void *const *mu,
                            const void *const *nu, void *const *mu_out, void *const *nu_out, void *const *u_out,
                            const int64_t *n, const uint64_t *count, double b1, double b2, double eps,
                            double eps_root, void *stream) {
  return fused_forward_dispatch<false>(dtype, num_leaves, g, mu, nu, mu_out, nu_out, u_out, n, count, b1, b2, eps,
                                       eps_root, stream);
}

int b200_adam_forward_inplace_mt(int dtype, int64_t num_leaves, void *const *updates, void *const *mu,
                                 void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2,
                                 double eps, double eps_root, void *stream) {
  return fused_forward_dispatch<true>(dtype, num_leaves, updates, mu, nu, mu, nu, updates, n, count, b1, b2, eps,
                                      eps_root, stream);
}

int b200_adam_forward_inplace(int dtype, void *updates, void *mu, void *nu, int64_t n, double b1, double b2,
                              double eps, double eps_root, uint64_t count, void *stream) {
  return b200_adam_forward_inplace_mt(dtype, 1, &updates, &mu, &nu, &n, &count, b1, b2, eps, eps_root, stream);
}

int b200_adam_fused_backward(int dtype, int64_t num_leaves, const void *const *du, const void *const *u,
                             const void *const *new_mu, const void *const *g, const void *const *dmu_ext,
                             const void *const *dnu_ext, void *const *dg, void *const *dmu, void *const *dnu,
                             const int64_t *n, const uint64_t *count, double b1, double b2, double eps_root,
                             void *stream) {
  if (num_leaves < 0) return B200_ADAM_E_ARG;
  if (num_leaves == 0) return 0;
  if (!du || !u || !new_mu || !g || !dmu_ext || !dnu_ext || !dg || !dmu || !dnu || !n || !count)
    return B200_ADAM_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case B200_ADAM_F32:
      return fused_backward_impl<float, float, float>(num_leaves, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n,
                                                      count, b1, b2, eps_root, s);
#ifndef B200_ADAM_TUNE_F32_ONLY
    case B200_ADAM_F64:
      return fused_backward_impl<double, double, double>(num_leaves, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu,
                                                         n, count, b1, b2, eps_root, s);
    case B200_ADAM_BF16_F32:
      return fused_backward_impl<bf16_t, float, float>(num_leaves, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n,
                                                       count, b1, b2, eps_root, s);
#endif
    default:
      return B200_ADAM_E_DTYPE;
  }
}

int b200_adam_forward_mu(int dtype, const void *updates, const void *mu, void *mu_out, int64_t n, double b1,
                         void *stream) {
  B200_UNFUSED_DISPATCH(OP_FWD_MU, updates, mu, nullptr, mu_out, nullptr, b1,
                        (dtype == B200_ADAM_F32 ? (double)omb<float>(b1) : omb<double>(b1)), 0, 0)
}

int b200_adam_forward_nu(int dtype, const void *updates, const void *nu, void *nu_out, int64_t n, double b2,
                         void *stream) {
  B200_UNFUSED_DISPATCH(OP_FWD_NU, updates, nu, nullptr, nu_out, nullptr, b2,
                        (dtype == B200_ADAM_F32 ? (double)omb<float>(b2) : omb<double>(b2)), 0, 0)
}

int b200_adam_forward_updates(int dtype, const void *new_mu, const void *new_nu, void *updates_out, int64_t n,
                              double b1, double b2, double eps, double eps_root, uint64_t count, void *stream) {
  const double c1 = inv_one_minus_pow(b1, count), c2 = inv_one_minus_pow(b2, count);
  B200_UNFUSED_DISPATCH(OP_FWD_UPD, new_mu, new_nu, nullptr, updates_out, nullptr, c1, c2, eps_root, eps)
}

int b200_adam_backward_mu(int dtype, const void *dmu, void *dupdates_out, void *dmu_out, int64_t n, double b1,
                          void *stream) {
  B200_UNFUSED_DISPATCH(OP_BWD_MU, dmu, nullptr, nullptr, dupdates_out, dmu_out, b1,
                        (dtype == B200_ADAM_F32 ? (double)omb<float>(b1) : omb<double>(b1)), 0, 0)
}

int b200_adam_backward_nu(int dtype, const void *dnu, const void *updates, void *dupdates_out, void *dnu_out,
                          int64_t n, double b2, void *stream) {
  // TWO_OMB2 = T(2) * (T(1) - rn_T(b2)) is exact in T
  B200_UNFUSED_DISPATCH(OP_BWD_NU, dnu, updates, nullptr, dupdates_out, dnu_out, b2,
                        (dtype == B200_ADAM_F32 ? (double)(2.0f * omb<float>(b2)) : 2.0 * omb<double>(b2)), 0, 0)
}

int b200_adam_backward_updates(int dtype, const void *dupdates, const void *updates, const void *new_mu,
                               void *dnew_mu_out, void *dnew_nu_out, int64_t n, double b1, double b2,
                               double eps_root, uint64_t count, void *stream) {
  (void)b2;
  const double o1 = one_minus_pow(b1, count), c2e = inv_one_minus_pow_eps(b2, eps_root, count);
  B200_UNFUSED_DISPATCH(OP_BWD_UPD, dupdates, updates, new_mu, dnew_mu_out, dnew_nu_out, o1, c2e, 0, 0)
}


static int step_forward_dispatch(bool inplace, int dtype, int64_t L, const void *const *p, const void *const *g,
                                 const void *const *mu, const void *const *nu, void *const *p_out,
                                 void *const *mu_out, void *const *nu_out, void *const *us_out, const int64_t *n,
                                 const uint64_t *count, double b1, double b2, double eps, double eps_root,
                                 double step_size, double wd_l2, double wd_decoupled, int maximize, void *stream) {
  if (L < 0) return B200_ADAM_E_ARG;
  if (L == 0) return 0;
  if (!p || !g || !mu || !nu || !p_out || !mu_out || !nu_out || !n || !count) return B200_ADAM_E_ARG;
  if (wd_l2 != 0.0 && wd_decoupled != 0.0) return B200_ADAM_E_ARG;  // one recipe at a time (adam xor adamw)
  cudaStream_t s = (cudaStream_t)stream;
#define B200_STEP_FWD(TG, TS, CT)                                                                                   \
  return inplace ? step_forward_impl<TG, TS, CT, true>(L, p, g, mu, nu, p_out, mu_out, nu_out, us_out, n, count, b1, b2, \
                                                       eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, s)  \
                 : step_forward_impl<TG, TS, CT, false>(L, p, g, mu, nu, p_out, mu_out, nu_out, us_out, n, count, b1,   \
                                                        b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, s)
#ifndef B200_ADAM_TUNE_F32_ONLY
  if (dtype == B200_ADAM_F64) B200_STEP_FWD(double, double, double);
  if (dtype == B200_ADAM_BF16_F32) B200_STEP_FWD(bf16_t, float, float);
#endif
  if (dtype == B200_ADAM_F32) B200_STEP_FWD(float, float, float);
#undef B200_STEP_FWD
  return B200_ADAM_E_DTYPE;
}

int b200_adam_step_forward(int dtype, int64_t num_leaves, const void *const *p, const void *const *g,
                           const void *const *mu, const void *const *nu, void *const *p_out, void *const *mu_out,
                           void *const *nu_out, void *const *us_out, const int64_t *n, const uint64_t *count,
                           double b1, double b2, double eps, double eps_root, double step_size, double wd_l2,
                           double wd_decoupled, int maximize, void *stream) {
  return step_forward_dispatch(false, dtype, num_leaves, p, g, mu, nu, p_out, mu_out, nu_out, us_out, n, count, b1, b2,
                               eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, stream);
}

int b200_adam_step_inplace(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                           void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2, double eps,
                           double eps_root, double step_size, double wd_l2, double wd_decoupled, int maximize,
                           void *stream) {
  return step_forward_dispatch(true, dtype, num_leaves, p, g, mu, nu, p, mu, nu, g, n, count, b1, b2, eps, eps_root,
                               step_size, wd_l2, wd_decoupled, maximize, stream);
}

int b200_adam_step_inplace_ex(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                              void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2,
                              double eps, double eps_root, double step_size, double wd_l2, double wd_decoupled,
                              int maximize, int overwrite_grads, void *stream) {
  return step_forward_dispatch(true, dtype, num_leaves, p, g, mu, nu, p, mu, nu, overwrite_grads ? g : nullptr, n, count,
                               b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, stream);
}

int64_t b200_adam_bias_table_len(int dtype, double b1, double b2) {
  if (!(b1 >= 0.0 && b1 < 1.0 && b2 >= 0.0 && b2 < 1.0)) return B200_ADAM_E_ARG;
  const int64_t cap = int64_t(1) << 24;
  int64_t a, b;
  if (dtype == B200_ADAM_F64) {
    a = bias_table_len_one<double>(b1, cap), b = bias_table_len_one<double>(b2, cap);
  } else if (dtype == B200_ADAM_F32 || dtype == B200_ADAM_BF16_F32) {
    a = bias_table_len_one<float>(b1, cap), b = bias_table_len_one<float>(b2, cap);
  } else {
    return B200_ADAM_E_DTYPE;
  }
  if (a == 0 || b == 0) return B200_ADAM_E_ARG;  // a beta so close to 1 that the table would exceed 2^24 entries
  return a > b ? a : b;
}

int b200_adam_bias_table(int dtype, double b1, double b2, int64_t len, void *c1_host, void *c2_host) {
  if (len <= 0 || !c1_host || !c2_host) return B200_ADAM_E_ARG;
  if (dtype == B200_ADAM_F64) {
    for (int64_t c = 1; c <= len; ++c) {
      ((double *)c1_host)[c - 1] = inv_one_minus_pow(b1, (uint64_t)c);
      ((double *)c2_host)[c - 1] = inv_one_minus_pow(b2, (uint64_t)c);
    }
    return 0;
  }
  if (dtype == B200_ADAM_F32 || dtype == B200_ADAM_BF16_F32) {
    for (int64_t c = 1; c <= len; ++c) {
      ((float *)c1_host)[c - 1] = (float)inv_one_minus_pow(b1, (uint64_t)c);
      ((float *)c2_host)[c - 1] = (float)inv_one_minus_pow(b2, (uint64_t)c);
    }
    return 0;
  }
  return B200_ADAM_E_DTYPE;
}

int b200_adam_step_inplace_capturable(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                                      void *const *nu, const int64_t *n, const int64_t *const *count_dev,
                                      const void *c1_table_dev, const void *c2_table_dev, int64_t table_len, double b1,
                                      double b2, double eps, double eps_root, double step_size, double wd_l2,
                                      double wd_decoupled, int maximize, int overwrite_grads, void *stream) {
  if (num_leaves < 0) return B200_ADAM_E_ARG;
  if (num_leaves == 0) return 0;
  if (!p || !g || !mu || !nu || !n || !count_dev || !c1_table_dev || !c2_table_dev || table_len <= 0 ||
      table_len > (int64_t(1) << 24))
    return B200_ADAM_E_ARG;
  if (wd_l2 != 0.0 && wd_decoupled != 0.0) return B200_ADAM_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
#define B200_STEP_CAP(TG, TS, CT)                                                                                     \
  return step_capturable_impl<TG, TS, CT>(num_leaves, p, g, mu, nu, n, count_dev, c1_table_dev, c2_table_dev, table_len, \
                                          b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize,              \
                                          overwrite_grads, s)
#ifndef B200_ADAM_TUNE_F32_ONLY
  if (dtype == B200_ADAM_F64) B200_STEP_CAP(double, double, double);
  if (dtype == B200_ADAM_BF16_F32) B200_STEP_CAP(bf16_t, float, float);
#endif
  if (dtype == B200_ADAM_F32) B200_STEP_CAP(float, float, float);
#undef B200_STEP_CAP
  return B200_ADAM_E_DTYPE;
}

int b200_adam_step_backward(int dtype, int64_t num_leaves, const void *const *dp, const void *const *dus,
                            const void *const *new_mu, const void *const *new_nu, const void *const *g,
                            const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg, void *const *dmu,
                            void *const *dnu, const void *const *p, void *const *dp_out, const int64_t *n,
                            const uint64_t *count, double b1, double b2, double eps, double eps_root, double step_size,
                            double wd_l2, double wd_decoupled, int maximize, void *stream) {
  if (num_leaves < 0) return B200_ADAM_E_ARG;
  if (num_leaves == 0) return 0;
  if (!dp || !new_mu || !new_nu || !g || !dmu_ext || !dnu_ext || !dg || !dmu || !dnu || !n || !count)
    return B200_ADAM_E_ARG;
  if (wd_l2 != 0.0 && wd_decoupled != 0.0) return B200_ADAM_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
#define B200_STEP_BWD_DT(TG, TS, CT)                                                                                 \
  return step_backward_impl<TG, TS, CT>(num_leaves, dp, dus, new_mu, new_nu, g, dmu_ext, dnu_ext, dg, dmu, dnu, p,     \
                                        dp_out, n, count, b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled,      \
                                        maximize, s)
#ifndef B200_ADAM_TUNE_F32_ONLY
  if (dtype == B200_ADAM_F64) B200_STEP_BWD_DT(double, double, double);
  if (dtype == B200_ADAM_BF16_F32) B200_STEP_BWD_DT(bf16_t, float, float);
#endif
  if (dtype == B200_ADAM_F32) B200_STEP_BWD_DT(float, float, float);
#undef B200_STEP_BWD_DT
  return B200_ADAM_E_DTYPE;
}

// ---- prepared launches: the argument block of a fused forward + backward pair kept on the library side ------------
struct PreparedPair {
  int dtype;
  int64_t L;
  std::vector<const void *> g, mu, nu, du, u, new_mu, dmu_ext, dnu_ext;
  std::vector<void *> mu_out, nu_out, u_out, dg, dmu, dnu;
  std::vector<int64_t> n;
  std::vector<uint64_t> count;
  double b1, b2, eps, eps_root;
};

void *b200_adam_prepare_fused(int dtype, int64_t num_leaves, const void *const *g, const void *const *mu,
                              const void *const *nu, void *const *mu_out, void *const *nu_out, void *const *u_out,
                              const void *const *du, const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg,
                              void *const *dmu, void *const *dnu, const int64_t *n, const uint64_t *count, double b1,
                              double b2, double eps, double eps_root) {
  if (num_leaves <= 0 || !g || !mu || !nu || !mu_out || !nu_out || !u_out || !du || !dmu_ext || !dnu_ext || !dg || !dmu ||
      !dnu || !n || !count)
    return nullptr;
  auto *p = new PreparedPair();
  p->dtype = dtype, p->L = num_leaves, p->b1 = b1, p->b2 = b2, p->eps = eps, p->eps_root = eps_root;
  const size_t L = (size_t)num_leaves;
  p->g.assign(g, g + L), p->mu.assign(mu, mu + L), p->nu.assign(nu, nu + L);
  p->mu_out.assign(mu_out, mu_out + L), p->nu_out.assign(nu_out, nu_out + L), p->u_out.assign(u_out, u_out + L);
  p->du.assign(du, du + L), p->dmu_ext.assign(dmu_ext, dmu_ext + L), p->dnu_ext.assign(dnu_ext, dnu_ext + L);
  p->dg.assign(dg, dg + L), p->dmu.assign(dmu, dmu + L), p->dnu.assign(dnu, dnu + L);
  p->n.assign(n, n + L), p->count.assign(count, count + L);
  // the backward reads what the forward wrote: u and new_mu are the forward's outputs
  p->u.assign(u_out, u_out + L), p->new_mu.assign(mu_out, mu_out + L);
  return p;
}

int b200_adam_launch_prepared(void *handle, int backward, void *stream) {
  if (!handle) return B200_ADAM_E_ARG;
  const PreparedPair &p = *(const PreparedPair *)handle;
  if (!backward)
    return b200_adam_fused_forward(p.dtype, p.L, p.g.data(), p.mu.data(), p.nu.data(), p.mu_out.data(), p.nu_out.data(),
                                   p.u_out.data(), p.n.data(), p.count.data(), p.b1, p.b2, p.eps, p.eps_root, stream);
  return b200_adam_fused_backward(p.dtype, p.L, p.du.data(), p.u.data(), p.new_mu.data(), p.g.data(), p.dmu_ext.data(),
                                  p.dnu_ext.data(), p.dg.data(), p.dmu.data(), p.dnu.data(), p.n.data(), p.count.data(), p.b1,
                                  p.b2, p.eps_root, stream);
}

int b200_adam_release_prepared(void *handle) {
  delete (PreparedPair *)handle;
  return 0;
}

int64_t b200_adam_bias_table_len_ex(int dtype, double b1, double b2, double eps_root) {
  if (!(b1 >= 0.0 && b1 < 1.0 && b2 >= 0.0 && b2 < 1.0) || !(eps_root >= 0.0)) return B200_ADAM_E_ARG;
  const int64_t cap = int64_t(1) << 24;
  int64_t len;
  if (dtype == B200_ADAM_F64) len = bias_table_len_all<double>(b1, b2, eps_root, cap);
  else if (dtype == B200_ADAM_F32 || dtype == B200_ADAM_BF16_F32) len = bias_table_len_all<float>(b1, b2, eps_root, cap);
  else return B200_ADAM_E_DTYPE;
  return len == 0 ? (int64_t)B200_ADAM_E_ARG : len;
}

int b200_adam_bias_table_ex(int dtype, double b1, double b2, double eps_root, int64_t len, void *c1_host, void *c2_host,
                            void *o1_host, void *c2e_host) {
  if (len <= 0 || !c1_host || !c2_host || !o1_host || !c2e_host) return B200_ADAM_E_ARG;
  if (dtype == B200_ADAM_F64) fill_bias_tables<double>(b1, b2, eps_root, len, c1_host, c2_host, o1_host, c2e_host);
  else if (dtype == B200_ADAM_F32 || dtype == B200_ADAM_BF16_F32)
    fill_bias_tables<float>(b1, b2, eps_root, len, c1_host, c2_host, o1_host, c2e_host);
  else return B200_ADAM_E_DTYPE;
  return 0;
}

int b200_adam_step_forward_capturable(int dtype, int64_t num_leaves, const void *const *p, const void *const *g,
                                      const void *const *mu, const void *const *nu, void *const *p_out,
                                      void *const *mu_out, void *const *nu_out, void *const *us_out, const int64_t *n,
                                      const int64_t *const *count_dev, const void *c1_table_dev, const void *c2_table_dev,
                                      int64_t table_len, double b1, double b2, double eps, double eps_root,
                                      double step_size, double wd_l2, double wd_decoupled, int maximize, void *stream) {
  if (num_leaves < 0) return B200_ADAM_E_ARG;
  if (num_leaves == 0) return 0;
  if (!p || !g || !mu || !nu || !p_out || !mu_out || !nu_out || !n || !count_dev || !c1_table_dev || !c2_table_dev ||
      table_len <= 0 || table_len > (int64_t(1) << 24))
    return B200_ADAM_E_ARG;
  if (wd_l2 != 0.0 && wd_decoupled != 0.0) return B200_ADAM_E_ARG;
  DeviceCounts dc;
  dc.count_dev = count_dev, dc.c1 = c1_table_dev, dc.c2 = c2_table_dev, dc.len = table_len;
  cudaStream_t s = (cudaStream_t)stream;
#define B200_STEP_FWD_CAP(TG, TS, CT)                                                                                 \
  return step_forward_impl<TG, TS, CT, false>(num_leaves, p, g, mu, nu, p_out, mu_out, nu_out, us_out, n, nullptr, b1, b2, \
                                              eps, eps_root, step_size, wd_l2, wd_decoupled, maximize, s, dc)
#ifndef B200_ADAM_TUNE_F32_ONLY
  if (dtype == B200_ADAM_F64) B200_STEP_FWD_CAP(double, double, double);
  if (dtype == B200_ADAM_BF16_F32) B200_STEP_FWD_CAP(bf16_t, float, float);
#endif
  if (dtype == B200_ADAM_F32) B200_STEP_FWD_CAP(float, float, float);
#undef B200_STEP_FWD_CAP
  return B200_ADAM_E_DTYPE;
}

int b200_adam_step_backward_capturable(int dtype, int64_t num_leaves, const void *const *dp, const void *const *dus,
                                       const void *const *new_mu, const void *const *new_nu, const void *const *g,
                                       const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg,
                                       void *const *dmu, void *const *dnu, const void *const *p, void *const *dp_out,
                                       const int64_t *n, const int64_t *const *count_dev, const void *c1_table_dev,
                                       const void *c2_table_dev, const void *o1_table_dev, const void *c2e_table_dev,
                                       int64_t table_len, double b1, double b2, double eps, double eps_root,
                                       double step_size, double wd_l2, double wd_decoupled, int maximize, void *stream) {
  if (num_leaves < 0) return B200_ADAM_E_ARG;
  if (num_leaves == 0) return 0;
  if (!dp || !new_mu || !new_nu || !g || !dmu_ext || !dnu_ext || !dg || !dmu || !dnu || !n || !count_dev ||
      !c1_table_dev || !c2_table_dev || !o1_table_dev || !c2e_table_dev || table_len <= 0 ||
      table_len > (int64_t(1) << 24))
    return B200_ADAM_E_ARG;
  if (wd_l2 != 0.0 && wd_decoupled != 0.0) return B200_ADAM_E_ARG;
  DeviceCounts dc;
  dc.count_dev = count_dev, dc.c1 = c1_table_dev, dc.c2 = c2_table_dev, dc.o1 = o1_table_dev, dc.c2e = c2e_table_dev;
  dc.len = table_len;
  cudaStream_t s = (cudaStream_t)stream;
#define B200_STEP_BWD_CAP(TG, TS, CT)                                                                                  \
  return step_backward_impl<TG, TS, CT>(num_leaves, dp, dus, new_mu, new_nu, g, dmu_ext, dnu_ext, dg, dmu, dnu, p,       \
                                        dp_out, n, nullptr, b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled,       \
                                        maximize, s, dc)
#ifndef B200_ADAM_TUNE_F32_ONLY
  if (dtype == B200_ADAM_F64) B200_STEP_BWD_CAP(double, double, double);
  if (dtype == B200_ADAM_BF16_F32) B200_STEP_BWD_CAP(bf16_t, float, float);
#endif
  if (dtype == B200_ADAM_F32) B200_STEP_BWD_CAP(float, float, float);
#undef B200_STEP_BWD_CAP
  return B200_ADAM_E_DTYPE;
}

}  // extern "C"
