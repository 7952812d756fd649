// adam_kernels.cuh -- device code of the B200 (sm_100a) accelerated differentiable Adam operator.
//
// One streaming engine (`stream_kernel`) + one small policy struct per operator.  Every operator of the
// path is an element-wise HBM stream (arithmetic intensity < 1 flop/B, no contraction => no tensor cores),
// so the engine is built for bytes in flight:
//   * 128-bit or 256-bit (LDG.E.256 / STG.E.256, new on sm_100) vector accesses, fully coalesced: lane k of a
//     warp touches vector k of a 32-vector run; the read-only streams use ld.global.nc + L1::no_allocate;
//   * U independent vector loads per array are issued front-batched before any use (MLP = U x arrays);
//   * the work unit is a "chunk" (one or a few block-tiles of ONE leaf); CTAs walk chunks in grid-stride order, so
//     the CTAs resident at any instant cover one contiguous window of every array.  By default the grid has one
//     CTA per chunk (measured best on B200: the hardware CTA scheduler back-fills SMs and de-phases the load and
//     store bursts of neighbouring CTAs, +6-9% over a lock-step persistent grid); a persistent grid of k CTAs per
//     SM is a run-time option (b200_adam_set_tuning);
//   * multi-tensor: the leaf table (pointers, sizes, per-leaf bias corrections, chunk prefix sums) rides in
//     the kernel parameters (__grid_constant__, up to 32 KB on sm_100) -- no device-side table, no extra copy,
//     graph-capturable; a single flat tensor is simply the 1-leaf table.
//
// Arithmetic contract (reference loops in /root/reference/src/adam_op/adam_op_impl_cpu.cpp):
// every * + / sqrt is one correctly rounded IEEE operation in the tensor dtype, in the reference's source order,
// never contracted into FMA (explicit __f*_rn / __d*_rn intrinsics), subnormals kept.  Results are therefore
// bit-identical to the reference CPU build.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200adam {

// ------------------------------------------------------------------------------------------------
// element types
// ------------------------------------------------------------------------------------------------
struct bf16_t {
  uint16_t bits;
};

template <typename CT>
struct Math;

template <>
struct Math<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  // tail of backward_updates: the reference's literal 0.5 is a double, so (t*0.5*c2e*den) is evaluated in
  // fp64 and rounded to float once (adam_op_impl_cpu.cpp:315-316).
  static __device__ __forceinline__ float dnu_tail(float t, float c2e, float den) {
    const double x = __dmul_rn(__dmul_rn(__dmul_rn((double)t, 0.5), (double)c2e), (double)den);
    return __double2float_rn(x);
  }
};

template <>
struct Math<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double dnu_tail(double t, double c2e, double den) {
    return __dmul_rn(__dmul_rn(__dmul_rn(t, 0.5), c2e), den);
  }
};

template <typename CT, typename T>
__device__ __forceinline__ CT up(T v) {
  return (CT)v;
}
template <>
__device__ __forceinline__ float up<float, bf16_t>(bf16_t v) {
  return __uint_as_float(((uint32_t)v.bits) << 16);  // exact
}

template <typename T, typename CT>
struct Down {
  static __device__ __forceinline__ T cvt(CT v) { return (T)v; }
};
template <>
struct Down<bf16_t, float> {
  // round-to-nearest-even float -> bf16, NaN kept quiet (same result as __float2bfloat16_rn)
  static __device__ __forceinline__ bf16_t cvt(float f) {
    uint32_t x = __float_as_uint(f);
    bf16_t r;
    if ((x & 0x7fffffffu) > 0x7f800000u) {
      r.bits = 0x7fffu;
    } else {
      x += 0x7fffu + ((x >> 16) & 1u);
      r.bits = (uint16_t)(x >> 16);
    }
    return r;
  }
};

// ------------------------------------------------------------------------------------------------
// vector global-memory access: BYTES in {8, 16, 32}; NC = read-only (non-coherent) path allowed
// ------------------------------------------------------------------------------------------------
// L2 eviction priority of the streams (every byte is touched exactly once per launch); overridable for sweeps.
#ifndef B200_LD_L2
#define B200_LD_L2 ""
#endif
#ifndef B200_ST_L2
#define B200_ST_L2 ""
#endif

template <int BYTES, bool NC>
struct VecIO;

template <bool NC>
struct VecIO<32, NC> {
  static __device__ __forceinline__ void ld(uint32_t (&w)[8], const void *p) {
    if (NC)
      asm volatile("ld.global.nc.L1::no_allocate" B200_LD_L2 ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(p));
    else
      asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(p)
                   : "memory");
  }
  static __device__ __forceinline__ void st(void *p, const uint32_t (&w)[8]) {
    asm volatile("st.global.L1::no_allocate" B200_ST_L2 ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  }
};

template <bool NC>
struct VecIO<16, NC> {
  static __device__ __forceinline__ void ld(uint32_t (&w)[4], const void *p) {
    if (NC)
      asm volatile("ld.global.nc.L1::no_allocate" B200_LD_L2 ".v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "l"(p));
    else
      asm volatile("ld.global.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "l"(p)
                   : "memory");
  }
  static __device__ __forceinline__ void st(void *p, const uint32_t (&w)[4]) {
    asm volatile("st.global.L1::no_allocate" B200_ST_L2 ".v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]),
                 "r"(w[2]), "r"(w[3])
                 : "memory");
  }
};

template <bool NC>
struct VecIO<8, NC> {
  static __device__ __forceinline__ void ld(uint32_t (&w)[2], const void *p) {
    if (NC)
      asm volatile("ld.global.nc.L1::no_allocate" B200_LD_L2 ".v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    else
      asm volatile("ld.global.L1::no_allocate.v2.b32 {%0,%1}, [%2];"
                   : "=r"(w[0]), "=r"(w[1])
                   : "l"(p)
                   : "memory");
  }
  static __device__ __forceinline__ void st(void *p, const uint32_t (&w)[2]) {
    asm volatile("st.global.L1::no_allocate" B200_ST_L2 ".v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
  }
};

// P elements of type T held in registers; loaded/stored as ONE vector access when P > 1.
template <typename T, int P>
union Pack {
  T v[P];
  uint32_t w[(sizeof(T) * P + 3) / 4];
  __device__ __forceinline__ Pack() {}
};

template <bool NC, typename T, int P>
__device__ __forceinline__ void load_pack(Pack<T, P> &r, const T *p) {
  if constexpr (P == 1) {
    r.v[0] = *p;
  } else {
    VecIO<(int)sizeof(T) * P, NC>::ld(r.w, p);
  }
}

// streams that are absent in a launch leave their Pack untouched; zero it so the compiler keeps it in registers
template <typename T, int P>
__device__ __forceinline__ void zero_pack(Pack<T, P> &r) {
  if constexpr (sizeof(T) * P < 4) {
    // a sub-word pack (one bf16 element on the scalar path): zero it through its own type -- writing the 32-bit word
    // of the union and reading 16 bits back makes the compiler park the pack in local memory
#pragma unroll
    for (int k = 0; k < P; ++k) r.v[k] = T();
  } else {
#pragma unroll
    for (int k = 0; k < (int)((sizeof(T) * P + 3) / 4); ++k) r.w[k] = 0u;
  }
}

template <bool NC, typename T, int P>
__device__ __forceinline__ void load_or_zero(bool present, Pack<T, P> &r, const T *p) {
  if (present) {
    load_pack<NC>(r, p);
  } else {
    zero_pack(r);
  }
}

template <typename T, int P>
__device__ __forceinline__ void store_pack(T *p, const Pack<T, P> &r) {
  if constexpr (P == 1) {
    *p = r.v[0];
  } else {
    VecIO<(int)sizeof(T) * P, false>::st(p, r.w);
  }
}

template <typename T, int P>
__device__ __forceinline__ bool is_aligned(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & (uintptr_t)(sizeof(T) * P - 1)) == 0;
}

// ------------------------------------------------------------------------------------------------
// leaf table (kernel parameter)
// ------------------------------------------------------------------------------------------------
template <class Op, int MAXL>
struct Table {
  int num_leaves;
  int chunk_elems;   // elements per chunk; a multiple of the block tile (hence of the vector width)
  int total_chunks;  // == chunk_start[num_leaves]
  int flags;         // operator specific
  typename Op::CT gs[Op::NGS];  // scalars shared by all leaves (betas, eps, ...)
  int chunk_start[MAXL + 1];    // exclusive prefix sum of per-leaf chunk counts
  long long n[MAXL];
  void *ptr[Op::NPTR][MAXL];
  typename Op::CT ls[Op::NLS > 0 ? Op::NLS : 1][MAXL];  // per-leaf scalars (bias corrections: depend on count)
};

// ------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------
// forward of one differentiable Adam step: reads g, mu, nu ; writes mu', nu', u.
//   mu' = (B1*mu) + (OMB1*g)                    adam_op_impl_cpu.cpp:105
//   nu' = (B2*nu) + ((OMB2*g)*g)                adam_op_impl_cpu.cpp:140
//   u   = (mu'*C1) / (sqrt((nu'*C2)+ER) + EPS)  adam_op_impl_cpu.cpp:176-179
// INPLACE = outputs may alias inputs (forward_, adam_op_impl_cpu.cpp:46-60): coherent loads, and the stores
// are issued in the reference's order mu, nu, updates so that fully aliased arguments end up holding u.
template <typename TG_, typename TS_, typename CT_, bool INPLACE>
struct FusedFwd {
  using TG = TG_;
  using TS = TS_;
  using CT = CT_;
  static constexpr int NPTR = 6, NGS = 6, NLS = 2;
  struct Leaf {
    const TG *g;
    const TS *mu, *nu;
    TS *mu_o, *nu_o;
    TG *u_o;
    CT B1, OMB1, B2, OMB2, ER, EPS, C1, C2;
  };
  template <class Tab>
  static __device__ __forceinline__ Leaf leaf(const Tab &t, int l) {
    Leaf L;
    L.g = (const TG *)t.ptr[0][l];
    L.mu = (const TS *)t.ptr[1][l];
    L.nu = (const TS *)t.ptr[2][l];
    L.mu_o = (TS *)t.ptr[3][l];
    L.nu_o = (TS *)t.ptr[4][l];
    L.u_o = (TG *)t.ptr[5][l];
    L.B1 = t.gs[0], L.OMB1 = t.gs[1], L.B2 = t.gs[2], L.OMB2 = t.gs[3], L.ER = t.gs[4], L.EPS = t.gs[5];
    L.C1 = t.ls[0][l], L.C2 = t.ls[1][l];
    return L;
  }
  template <int P>
  static __device__ __forceinline__ bool aligned(const Leaf &L) {
    return is_aligned<TG, P>(L.g) && is_aligned<TS, P>(L.mu) && is_aligned<TS, P>(L.nu) &&
           is_aligned<TS, P>(L.mu_o) && is_aligned<TS, P>(L.nu_o) && is_aligned<TG, P>(L.u_o);
  }
  template <int P>
  struct Regs {
    Pack<TG, P> g;  // becomes u
    Pack<TS, P> mu, nu;
  };
  template <int P>
  static __device__ __forceinline__ void load(Regs<P> &r, const Leaf &L, long long i) {
    load_pack<!INPLACE>(r.g, L.g + i);
    load_pack<!INPLACE>(r.mu, L.mu + i);
    load_pack<!INPLACE>(r.nu, L.nu + i);
  }
  template <int P>
  static __device__ __forceinline__ void compute(Regs<P> &r, const Leaf &L) {
    using M = Math<CT>;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const CT g = up<CT>(r.g.v[k]), m = up<CT>(r.mu.v[k]), v = up<CT>(r.nu.v[k]);
      const CT m_out = M::add(M::mul(L.B1, m), M::mul(L.OMB1, g));
      const CT v_out = M::add(M::mul(L.B2, v), M::mul(M::mul(L.OMB2, g), g));
      const CT m_hat = M::mul(m_out, L.C1);
      const CT v_hat = M::mul(v_out, L.C2);
      const CT u = M::div(m_hat, M::add(M::sqrt(M::add(v_hat, L.ER)), L.EPS));
      r.mu.v[k] = Down<TS, CT>::cvt(m_out);
      r.nu.v[k] = Down<TS, CT>::cvt(v_out);
      r.g.v[k] = Down<TG, CT>::cvt(u);
    }
  }
  template <int P>
  static __device__ __forceinline__ void store(const Regs<P> &r, const Leaf &L, long long i) {
    store_pack(L.mu_o + i, r.mu);
    store_pack(L.nu_o + i, r.nu);
    store_pack(L.u_o + i, r.g);
  }
};

// backward of one differentiable Adam step (see include/b200_adam.h, b200_adam_fused_backward).
// MODE: 0 = every leaf has du, dmu_ext, dnu_ext and wants dg, dmu, dnu (general unroll step, 36 B/elem)
//       1 = every leaf has du only and wants dg, dmu, dnu         (last unroll step, 28 B/elem)
//       2 = per-leaf presence decided at run time from NULL pointers (uniform per chunk)
enum { BWD_FULL = 0, BWD_LAST = 1, BWD_GENERIC = 2 };

template <typename TG_, typename TS_, typename CT_, int MODE>
struct FusedBwd {
  using TG = TG_;
  using TS = TS_;
  using CT = CT_;
  static constexpr int NPTR = 9, NGS = 4, NLS = 2;
  struct Leaf {
    const TG *du, *u, *g;
    const TS *m, *dmu_ext, *dnu_ext;
    TG *dg;
    TS *dmu, *dnu;
    CT B1, OMB1, B2, TWO_OMB2, O1, C2E;
  };
  template <class Tab>
  static __device__ __forceinline__ Leaf leaf(const Tab &t, int l) {
    Leaf L;
    L.du = (const TG *)t.ptr[0][l];
    L.u = (const TG *)t.ptr[1][l];
    L.m = (const TS *)t.ptr[2][l];
    L.g = (const TG *)t.ptr[3][l];
    L.dmu_ext = (const TS *)t.ptr[4][l];
    L.dnu_ext = (const TS *)t.ptr[5][l];
    L.dg = (TG *)t.ptr[6][l];
    L.dmu = (TS *)t.ptr[7][l];
    L.dnu = (TS *)t.ptr[8][l];
    L.B1 = t.gs[0], L.OMB1 = t.gs[1], L.B2 = t.gs[2], L.TWO_OMB2 = t.gs[3];
    L.O1 = t.ls[0][l], L.C2E = t.ls[1][l];
    return L;
  }
  static __device__ __forceinline__ bool has_du(const Leaf &L) { return MODE != BWD_GENERIC || L.du != nullptr; }
  static __device__ __forceinline__ bool has_dmu_ext(const Leaf &L) {
    return MODE == BWD_FULL || (MODE == BWD_GENERIC && L.dmu_ext != nullptr);
  }
  static __device__ __forceinline__ bool has_dnu_ext(const Leaf &L) {
    return MODE == BWD_FULL || (MODE == BWD_GENERIC && L.dnu_ext != nullptr);
  }
  static __device__ __forceinline__ bool want_dg(const Leaf &L) { return MODE != BWD_GENERIC || L.dg != nullptr; }
  static __device__ __forceinline__ bool want_dmu(const Leaf &L) { return MODE != BWD_GENERIC || L.dmu != nullptr; }
  static __device__ __forceinline__ bool want_dnu(const Leaf &L) { return MODE != BWD_GENERIC || L.dnu != nullptr; }
  // g is read only if a dnu_tot exists and dg is wanted
  static __device__ __forceinline__ bool need_g(const Leaf &L) {
    return want_dg(L) && (has_du(L) || has_dnu_ext(L));
  }

  template <int P>
  static __device__ __forceinline__ bool aligned(const Leaf &L) {
    // NULL pointers are aligned by construction
    return is_aligned<TG, P>(L.du) && is_aligned<TG, P>(L.u) && is_aligned<TS, P>(L.m) &&
           is_aligned<TG, P>(L.g) && is_aligned<TS, P>(L.dmu_ext) && is_aligned<TS, P>(L.dnu_ext) &&
           is_aligned<TG, P>(L.dg) && is_aligned<TS, P>(L.dmu) && is_aligned<TS, P>(L.dnu);
  }
  template <int P>
  struct Regs {
    Pack<TG, P> du, u, g;      // g becomes dg
    Pack<TS, P> m, dme, dne;   // dme/dne become dmu/dnu
  };
  template <int P>
  static __device__ __forceinline__ void load(Regs<P> &r, const Leaf &L, long long i) {
    load_or_zero<true>(has_du(L), r.du, L.du + i);
    load_or_zero<true>(has_du(L), r.u, L.u + i);
    load_or_zero<true>(has_du(L), r.m, L.m + i);
    load_or_zero<true>(need_g(L), r.g, L.g + i);
    load_or_zero<true>(has_dmu_ext(L), r.dme, L.dmu_ext + i);
    load_or_zero<true>(has_dnu_ext(L), r.dne, L.dnu_ext + i);
  }
  template <int P>
  static __device__ __forceinline__ void compute(Regs<P> &r, const Leaf &L) {
    using M = Math<CT>;
    const bool hdu = has_du(L), hme = has_dmu_ext(L), hne = has_dnu_ext(L), ng = need_g(L);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      CT dmu_tot = CT(0), dnu_tot = CT(0);
      if (hdu) {
        // backward_updates, adam_op_impl_cpu.cpp:298-316
        const CT du = up<CT>(r.du.v[k]), u = up<CT>(r.u.v[k]), m = up<CT>(r.m.v[k]);
        CT dm_new = CT(0), dn_new = CT(0);
        if (m != CT(0)) {
          const CT q = M::div(u, m);
          const CT den = M::mul(q, L.O1);
          dm_new = M::mul(du, q);
          const CT t = M::mul(M::mul(-du, u), den);
          dn_new = M::dnu_tail(t, L.C2E, den);
        }
        // autograd-engine accumulation with the gradients arriving from later steps (two addends: exact order-free)
        dmu_tot = hme ? M::add(up<CT>(r.dme.v[k]), dm_new) : dm_new;
        dnu_tot = hne ? M::add(up<CT>(r.dne.v[k]), dn_new) : dn_new;
      } else {
        if (hme) dmu_tot = up<CT>(r.dme.v[k]);
        if (hne) dnu_tot = up<CT>(r.dne.v[k]);
      }
      const bool have_m = hdu || hme, have_n = hdu || hne;
      // backward_mu, adam_op_impl_cpu.cpp:223-224 ; backward_nu, :261-262
      const CT dg_mu = M::mul(L.OMB1, dmu_tot);
      const CT dg_nu = ng ? M::mul(M::mul(L.TWO_OMB2, up<CT>(r.g.v[k])), dnu_tot) : CT(0);
      const CT dg = (have_m && have_n) ? M::add(dg_mu, dg_nu) : (have_m ? dg_mu : dg_nu);
      r.dme.v[k] = Down<TS, CT>::cvt(have_m ? M::mul(L.B1, dmu_tot) : CT(0));
      r.dne.v[k] = Down<TS, CT>::cvt(have_n ? M::mul(L.B2, dnu_tot) : CT(0));
      r.g.v[k] = Down<TG, CT>::cvt(dg);
    }
  }
  template <int P>
  static __device__ __forceinline__ void store(const Regs<P> &r, const Leaf &L, long long i) {
    if (want_dg(L)) store_pack(L.dg + i, r.g);
    if (want_dmu(L)) store_pack(L.dmu + i, r.dme);
    if (want_dnu(L)) store_pack(L.dnu + i, r.dne);
  }
};

// ------------------------------------------------------------------------------------------------
// "whole optimizer step" kernels: Adam update + scale(step_size) + apply_updates in one pass
// ------------------------------------------------------------------------------------------------
// forward:  mu', nu', u as FusedFwd ; us = u * S (transform/scale.py:86-105: g.mul(step_size)) ;
//           p' = p + us (update.py:78-79: p.add(u)).  Reads p, g, mu, nu (16 B/elem fp32), writes p', mu', nu'
//           (12 B) and, only if the caller wants the update tensor, us (4 B).  u itself is never stored: the
//           backward recomputes it from mu', nu' with the identical instruction sequence (same bits).
// Optional recipe terms, each skipped entirely when its coefficient is zero, like the reference
// drops the identity transformation (alias/utils.py:86-87, transform/add_decayed_weights.py:206-207):
//   maximize      g1 = -g                 (alias/utils.py:133-191: g.neg())
//   L2 decay      g2 = g1 + WD1 * p       (alias/utils.py:116-126: g.add(p, alpha=wd))         -- `adam(weight_decay=)`
//   decoupled     u2 = u  + WD2 * p       (transform/add_decayed_weights.py:237-247)            -- `adamw`
// `x.add(y, alpha=a)` on CUDA is ATen's AddFunctor `x + a * y` compiled with FMA contraction, i.e. ONE rounding:
// fma(a, y, x).  The kernels use the explicit fma so that they match the torch chain bit-for-bit
// (tests/test_step_gpu.py compares against torch itself).
// INPLACE:  p, g, mu, nu updated in place (g <- us), for the non-differentiable Optimizer.step path.
// Mixed precision (TG = bf16 for p, g, us and their gradients; TS = float moments; CT = float arithmetic): the chain
// materialises g2, u, u2, us and p' as bf16 TENSORS, so the fused kernel rounds to TG at exactly those points
// (`rd`); for TG = float / double `rd` is the identity.
template <typename TG, typename CT>
__device__ __forceinline__ CT rd(CT x) {
  return up<CT>(Down<TG, CT>::cvt(x));
}

// Capturable form of the step kernels (for a step recorded in a CUDA graph: the in-place training step of Adam /
// FlatAdam(capturable=True) and the differentiable step of MetaAdam / FuncAdam(capturable=True)): nothing about the
// step count may be baked into the launch, so
//   * the count is read from DEVICE memory (an int64 the caller increments on the device), and
//   * the bias corrections C1 = 1/(1-b1^c), C2 = 1/(1-b2^c) (forward, backward) and O1 = 1-b1^c,
//     C2E = 1/(1-b2^c+eps_root) (backward) come from device TABLES indexed by the count.  The tables are filled on the
//     host with the very expressions the host-scalar launchers use (std::pow in double, one rounding to CT), so results
//     stay bit-identical; they end where every correction has reached its limit value in CT (count 16 628 for
//     b2 = 0.999 in fp32), and larger counts clamp to the last entry.  `flags` of the table carries the table length.
__device__ __forceinline__ int bias_index(const void *count_ptr, int len) {
  const long long c = *(const long long *)count_ptr;
  return c < 1 ? 0 : (c > (long long)len ? len - 1 : (int)c - 1);
}

template <typename TG_, typename TS_, typename CT_, bool INPLACE>
struct StepFwd {
  using TG = TG_;
  using TS = TS_;
  using CT = CT_;
  // ptr[8..10] (optional, NULL = host scalars): capturable form -- see bias_index below
  static constexpr int NPTR = 11, NGS = 10, NLS = 2;
  struct Leaf {
    const TG *p, *g;
    const TS *mu, *nu;
    TG *p_o, *us_o;
    TS *mu_o, *nu_o;
    CT B1, OMB1, B2, OMB2, ER, EPS, S, WD1, WD2, SGN, C1, C2;
  };
  template <class Tab>
  static __device__ __forceinline__ Leaf leaf(const Tab &t, int l) {
    Leaf L;
    L.p = (const TG *)t.ptr[0][l], L.g = (const TG *)t.ptr[1][l];
    L.mu = (const TS *)t.ptr[2][l], L.nu = (const TS *)t.ptr[3][l];
    L.p_o = (TG *)t.ptr[4][l], L.mu_o = (TS *)t.ptr[5][l], L.nu_o = (TS *)t.ptr[6][l], L.us_o = (TG *)t.ptr[7][l];
    L.B1 = t.gs[0], L.OMB1 = t.gs[1], L.B2 = t.gs[2], L.OMB2 = t.gs[3], L.ER = t.gs[4], L.EPS = t.gs[5];
    L.S = t.gs[6], L.WD1 = t.gs[7], L.WD2 = t.gs[8], L.SGN = t.gs[9];
    if (t.ptr[8][l] != nullptr) {
      const int idx = bias_index(t.ptr[8][l], t.flags);
      L.C1 = ((const CT *)t.ptr[9][l])[idx], L.C2 = ((const CT *)t.ptr[10][l])[idx];
    } else {
      L.C1 = t.ls[0][l], L.C2 = t.ls[1][l];
    }
    return L;
  }
  template <int P>
  static __device__ __forceinline__ bool aligned(const Leaf &L) {
    return is_aligned<TG, P>(L.p) && is_aligned<TG, P>(L.g) && is_aligned<TS, P>(L.mu) && is_aligned<TS, P>(L.nu) &&
           is_aligned<TG, P>(L.p_o) && is_aligned<TS, P>(L.mu_o) && is_aligned<TS, P>(L.nu_o) &&
           is_aligned<TG, P>(L.us_o);
  }
  template <int P>
  struct Regs {
    Pack<TG, P> p, g;    // become p', us
    Pack<TS, P> mu, nu;  // become mu', nu'
  };
  template <int P>
  static __device__ __forceinline__ void load(Regs<P> &r, const Leaf &L, long long i) {
    load_pack<!INPLACE>(r.p, L.p + i);
    load_pack<!INPLACE>(r.g, L.g + i);
    load_pack<!INPLACE>(r.mu, L.mu + i);
    load_pack<!INPLACE>(r.nu, L.nu + i);
  }
  template <int P>
  static __device__ __forceinline__ void compute(Regs<P> &r, const Leaf &L) {
    using M = Math<CT>;
    const bool neg = L.SGN < CT(0), l2 = L.WD1 != CT(0), dec = L.WD2 != CT(0);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const CT p = up<CT>(r.p.v[k]), m = up<CT>(r.mu.v[k]), v = up<CT>(r.nu.v[k]);
      CT g = neg ? -up<CT>(r.g.v[k]) : up<CT>(r.g.v[k]);
      if (l2) g = rd<TG>(M::fma(L.WD1, p, g));
      const CT m_out = M::add(M::mul(L.B1, m), M::mul(L.OMB1, g));
      const CT v_out = M::add(M::mul(L.B2, v), M::mul(M::mul(L.OMB2, g), g));
      CT u = rd<TG>(M::div(M::mul(m_out, L.C1), M::add(M::sqrt(M::add(M::mul(v_out, L.C2), L.ER)), L.EPS)));
      if (dec) u = rd<TG>(M::fma(L.WD2, p, u));
      const CT us = rd<TG>(M::mul(u, L.S));
      r.mu.v[k] = Down<TS, CT>::cvt(m_out);
      r.nu.v[k] = Down<TS, CT>::cvt(v_out);
      r.g.v[k] = Down<TG, CT>::cvt(us);
      r.p.v[k] = Down<TG, CT>::cvt(M::add(p, us));
    }
  }
  template <int P>
  static __device__ __forceinline__ void store(const Regs<P> &r, const Leaf &L, long long i) {
    store_pack(L.mu_o + i, r.mu);
    store_pack(L.nu_o + i, r.nu);
    if (L.us_o != nullptr) store_pack(L.us_o + i, r.g);
    store_pack(L.p_o + i, r.p);
  }
};

// backward of the whole step.  Incoming: dp' (gradient of the new parameters), dmu_ext, dnu_ext (gradients of the new
// moments from later steps) and optionally dus (gradient of the returned scaled update, normally absent).
//   du_s  = dp' (+ dus)            AddBackward of p' = p + us: both inputs receive dp'      (dp = dp' is returned by
//                                  the binding as the same tensor -- no bytes move)
//   du    = du_s * S               MulBackward of us = u * S
//   u     = recomputed from mu', nu' (bit-identical to the forward's u)
//   then exactly FusedBwd: backward_updates, two-addend accumulations, backward_mu, backward_nu, dg = dg_mu + dg_nu
// Reads dp', mu', nu', g, dmu_ext, dnu_ext (24 B/elem fp32) and writes dg, dmu, dnu (12 B).
// With weight decay the parameter also feeds the update, so its gradient is no longer the pass-through dp':
//   decoupled: du2 = (dp'+dus)*S ; du = du2 ; dp_wd = du2 * WD2        (AddBackward with alpha: grad * alpha)
//   L2       : dg1 = dg2 ; dp_wd = dg2 * WD1 ;  maximize: dg = -dg1     (NegBackward)
//   dp = dp' + dp_wd (two addends) is written to `dp_out` (+4 B/elem).  backward_nu needs the gradient that ENTERED
//   the moment update, g2 = fma(WD1, p, +-g); it is re-derived from the saved raw g and p (+4 B/elem read, L2 only).
// MODE as for FusedBwd: BWD_FULL (all three incoming present), BWD_LAST (dp' only), BWD_GENERIC (run-time NULLs).
template <typename TG_, typename TS_, typename CT_, int MODE>
struct StepBwd {
  using TG = TG_;
  using TS = TS_;
  using CT = CT_;
  // ptr[12..16] (optional, NULL = host scalars): device step count + c1 / c2 / o1 / c2e tables (capturable form)
  static constexpr int NPTR = 17, NGS = 10, NLS = 4;
  static constexpr int MAXL_BIG = 160;  // 17 pointers + 4 scalars per leaf: 160 leaves keep the table under 32 KB
  struct Leaf {
    const TG *dp, *dus, *g, *p;
    const TS *m, *v, *dmu_ext, *dnu_ext;
    TG *dg, *dp_o;
    TS *dmu, *dnu;
    CT B1, OMB1, B2, TWO_OMB2, ER, EPS, S, WD1, WD2, SGN, O1, C2E, C1, C2;
  };
  template <class Tab>
  static __device__ __forceinline__ Leaf leaf(const Tab &t, int l) {
    Leaf L;
    L.dp = (const TG *)t.ptr[0][l], L.dus = (const TG *)t.ptr[1][l], L.m = (const TS *)t.ptr[2][l];
    L.v = (const TS *)t.ptr[3][l], L.g = (const TG *)t.ptr[4][l];
    L.dmu_ext = (const TS *)t.ptr[5][l], L.dnu_ext = (const TS *)t.ptr[6][l];
    L.dg = (TG *)t.ptr[7][l], L.dmu = (TS *)t.ptr[8][l], L.dnu = (TS *)t.ptr[9][l];
    L.p = (const TG *)t.ptr[10][l], L.dp_o = (TG *)t.ptr[11][l];
    L.B1 = t.gs[0], L.OMB1 = t.gs[1], L.B2 = t.gs[2], L.TWO_OMB2 = t.gs[3], L.ER = t.gs[4], L.EPS = t.gs[5];
    L.S = t.gs[6], L.WD1 = t.gs[7], L.WD2 = t.gs[8], L.SGN = t.gs[9];
    if (t.ptr[12][l] != nullptr) {
      const int idx = bias_index(t.ptr[12][l], t.flags);
      L.C1 = ((const CT *)t.ptr[13][l])[idx], L.C2 = ((const CT *)t.ptr[14][l])[idx];
      L.O1 = ((const CT *)t.ptr[15][l])[idx], L.C2E = ((const CT *)t.ptr[16][l])[idx];
    } else {
      L.O1 = t.ls[0][l], L.C2E = t.ls[1][l], L.C1 = t.ls[2][l], L.C2 = t.ls[3][l];
    }
    return L;
  }
  // raw p is read only when the effective gradient g2 = fma(WD1, p, +-g) has to be re-derived for backward_nu
  static __device__ __forceinline__ bool need_p(const Leaf &L) { return L.WD1 != CT(0) && L.p != nullptr; }
  static __device__ __forceinline__ bool want_dp(const Leaf &L) { return L.dp_o != nullptr; }
  static __device__ __forceinline__ bool has_dp(const Leaf &L) { return MODE != BWD_GENERIC || L.dp != nullptr; }
  static __device__ __forceinline__ bool has_dus(const Leaf &L) { return MODE == BWD_GENERIC && L.dus != nullptr; }
  static __device__ __forceinline__ bool has_du(const Leaf &L) { return has_dp(L) || has_dus(L); }
  static __device__ __forceinline__ bool has_dmu_ext(const Leaf &L) {
    return MODE == BWD_FULL || (MODE == BWD_GENERIC && L.dmu_ext != nullptr);
  }
  static __device__ __forceinline__ bool has_dnu_ext(const Leaf &L) {
    return MODE == BWD_FULL || (MODE == BWD_GENERIC && L.dnu_ext != nullptr);
  }
  static __device__ __forceinline__ bool want_dg(const Leaf &L) { return MODE != BWD_GENERIC || L.dg != nullptr; }
  static __device__ __forceinline__ bool want_dmu(const Leaf &L) { return MODE != BWD_GENERIC || L.dmu != nullptr; }
  static __device__ __forceinline__ bool want_dnu(const Leaf &L) { return MODE != BWD_GENERIC || L.dnu != nullptr; }
  // dg2 (the gradient w.r.t. the decayed gradient g2) is needed for dg itself AND, with L2 decay, for the parameter
  // gradient dp_wd = WD1 * dg2 -- also when the caller does not want dg (detached first-order gradients)
  static __device__ __forceinline__ bool need_dg2(const Leaf &L) {
    return want_dg(L) || (L.WD1 != CT(0) && want_dp(L));
  }
  static __device__ __forceinline__ bool need_g(const Leaf &L) {
    return need_dg2(L) && (has_du(L) || has_dnu_ext(L));
  }
  template <int P>
  static __device__ __forceinline__ bool aligned(const Leaf &L) {
    return is_aligned<TG, P>(L.dp) && is_aligned<TG, P>(L.dus) && is_aligned<TS, P>(L.m) && is_aligned<TS, P>(L.v) &&
           is_aligned<TG, P>(L.g) && is_aligned<TS, P>(L.dmu_ext) && is_aligned<TS, P>(L.dnu_ext) &&
           is_aligned<TG, P>(L.dg) && is_aligned<TS, P>(L.dmu) && is_aligned<TS, P>(L.dnu) && is_aligned<TG, P>(L.p) &&
           is_aligned<TG, P>(L.dp_o);
  }
  template <int P>
  struct Regs {
    Pack<TG, P> dp, dus, g, p;      // g -> dg, dp -> dp_out
    Pack<TS, P> m, v, dme, dne;     // dme -> dmu, dne -> dnu
  };
  template <int P>
  static __device__ __forceinline__ void load(Regs<P> &r, const Leaf &L, long long i) {
    load_or_zero<true>(has_dp(L), r.dp, L.dp + i);
    load_or_zero<true>(has_dus(L), r.dus, L.dus + i);
    load_or_zero<true>(has_du(L), r.m, L.m + i);
    load_or_zero<true>(has_du(L), r.v, L.v + i);
    load_or_zero<true>(need_g(L), r.g, L.g + i);
    load_or_zero<true>(need_g(L) && need_p(L), r.p, L.p + i);
    load_or_zero<true>(has_dmu_ext(L), r.dme, L.dmu_ext + i);
    load_or_zero<true>(has_dnu_ext(L), r.dne, L.dnu_ext + i);
  }
  template <int P>
  static __device__ __forceinline__ void compute(Regs<P> &r, const Leaf &L) {
    using M = Math<CT>;
    const bool hdp = has_dp(L), hdus = has_dus(L), hdu = has_du(L), hme = has_dmu_ext(L), hne = has_dnu_ext(L);
    const bool ng = need_g(L), np = need_p(L), neg = L.SGN < CT(0), l2 = L.WD1 != CT(0), dec = L.WD2 != CT(0);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      CT dmu_tot = CT(0), dnu_tot = CT(0), dp_wd = CT(0);
      const CT dp_in = hdp ? up<CT>(r.dp.v[k]) : CT(0);
      if (hdu) {
        const CT dus = (hdp && hdus) ? rd<TG>(M::add(dp_in, up<CT>(r.dus.v[k]))) : (hdp ? dp_in : up<CT>(r.dus.v[k]));
        const CT du = rd<TG>(M::mul(dus, L.S));
        if (dec) dp_wd = rd<TG>(M::mul(du, L.WD2));
        const CT m = up<CT>(r.m.v[k]);
        const CT u = rd<TG>(
            M::div(M::mul(m, L.C1), M::add(M::sqrt(M::add(M::mul(up<CT>(r.v.v[k]), L.C2), L.ER)), L.EPS)));
        CT dm_new = CT(0), dn_new = CT(0);
        if (m != CT(0)) {
          const CT q = M::div(u, m);
          const CT den = M::mul(q, L.O1);
          dm_new = M::mul(du, q);
          const CT t = M::mul(M::mul(-du, u), den);
          dn_new = M::dnu_tail(t, L.C2E, den);
        }
        dmu_tot = hme ? M::add(up<CT>(r.dme.v[k]), dm_new) : dm_new;
        dnu_tot = hne ? M::add(up<CT>(r.dne.v[k]), dn_new) : dn_new;
      } else {
        if (hme) dmu_tot = up<CT>(r.dme.v[k]);
        if (hne) dnu_tot = up<CT>(r.dne.v[k]);
      }
      const bool have_m = hdu || hme, have_n = hdu || hne;
      const CT dg_mu = M::mul(L.OMB1, dmu_tot);
      CT dg_nu = CT(0);
      if (ng) {
        CT g = neg ? -up<CT>(r.g.v[k]) : up<CT>(r.g.v[k]);  // the gradient that entered the moment update
        if (np) g = rd<TG>(M::fma(L.WD1, up<CT>(r.p.v[k]), g));
        dg_nu = M::mul(M::mul(L.TWO_OMB2, g), dnu_tot);
      }
      const CT dg2 = rd<TG>((have_m && have_n) ? M::add(dg_mu, dg_nu) : (have_m ? dg_mu : dg_nu));
      if (l2) dp_wd = rd<TG>(M::mul(dg2, L.WD1));
      r.g.v[k] = Down<TG, CT>::cvt(neg ? -dg2 : dg2);
      r.dme.v[k] = Down<TS, CT>::cvt(have_m ? M::mul(L.B1, dmu_tot) : CT(0));
      r.dne.v[k] = Down<TS, CT>::cvt(have_n ? M::mul(L.B2, dnu_tot) : CT(0));
      r.dp.v[k] = Down<TG, CT>::cvt(hdp ? M::add(dp_in, dp_wd) : dp_wd);  // only stored when dp_out is requested
    }
  }
  template <int P>
  static __device__ __forceinline__ void store(const Regs<P> &r, const Leaf &L, long long i) {
    if (want_dg(L)) store_pack(L.dg + i, r.g);
    if (want_dmu(L)) store_pack(L.dmu + i, r.dme);
    if (want_dnu(L)) store_pack(L.dnu + i, r.dne);
    if (want_dp(L)) store_pack(L.dp_o + i, r.dp);
  }
};

// ---- the six out-of-place operators of the reference surface (single dtype T for every array) ----
// Generic 2-in / 2-out shape; unused slots are compiled away through the NIN/NOUT constants.
enum { OP_FWD_MU = 0, OP_FWD_NU, OP_FWD_UPD, OP_BWD_MU, OP_BWD_NU, OP_BWD_UPD };

template <typename T, int KIND>
struct UnfusedOp {
  using CT = T;
  static constexpr int NIN = (KIND == OP_BWD_MU) ? 1 : (KIND == OP_BWD_UPD ? 3 : 2);
  static constexpr int NOUT = (KIND <= OP_FWD_UPD) ? 1 : 2;
  static constexpr int NPTR = NIN + NOUT, NGS = 4, NLS = 2;
  struct Leaf {
    const T *in[3];
    T *out[2];
    CT s0, s1, s2, s3;
  };
  template <class Tab>
  static __device__ __forceinline__ Leaf leaf(const Tab &t, int l) {
    Leaf L;
#pragma unroll
    for (int a = 0; a < NIN; ++a) L.in[a] = (const T *)t.ptr[a][l];
#pragma unroll
    for (int a = 0; a < NOUT; ++a) L.out[a] = (T *)t.ptr[NIN + a][l];
    L.s0 = t.gs[0], L.s1 = t.gs[1], L.s2 = t.gs[2], L.s3 = t.gs[3];
    return L;
  }
  template <int P>
  static __device__ __forceinline__ bool aligned(const Leaf &L) {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < NIN; ++a) ok = ok && is_aligned<T, P>(L.in[a]);
#pragma unroll
    for (int a = 0; a < NOUT; ++a) ok = ok && is_aligned<T, P>(L.out[a]);
    return ok;
  }
  template <int P>
  struct Regs {
    Pack<T, P> a[3];  // inputs; a[0], a[1] are overwritten with the outputs
  };
  template <int P>
  static __device__ __forceinline__ void load(Regs<P> &r, const Leaf &L, long long i) {
#pragma unroll
    for (int a = 0; a < NIN; ++a) load_pack<true>(r.a[a], L.in[a] + i);
  }
  template <int P>
  static __device__ __forceinline__ void compute(Regs<P> &r, const Leaf &L) {
    using M = Math<CT>;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if constexpr (KIND == OP_FWD_MU) {  // in: g, mu ; s0=B1 s1=OMB1                 (cpu.cpp:105)
        r.a[0].v[k] = M::add(M::mul(L.s0, r.a[1].v[k]), M::mul(L.s1, r.a[0].v[k]));
      } else if constexpr (KIND == OP_FWD_NU) {  // in: g, nu ; s0=B2 s1=OMB2          (cpu.cpp:140)
        const T g = r.a[0].v[k];
        r.a[0].v[k] = M::add(M::mul(L.s0, r.a[1].v[k]), M::mul(M::mul(L.s1, g), g));
      } else if constexpr (KIND == OP_FWD_UPD) {  // in: mu', nu' ; s0=C1 s1=C2 s2=ER s3=EPS (cpu.cpp:176-179)
        const T m_hat = M::mul(r.a[0].v[k], L.s0);
        const T v_hat = M::mul(r.a[1].v[k], L.s1);
        r.a[0].v[k] = M::div(m_hat, M::add(M::sqrt(M::add(v_hat, L.s2)), L.s3));
      } else if constexpr (KIND == OP_BWD_MU) {  // in: dmu ; s0=B1 s1=OMB1 ; out: dg, dmu (cpu.cpp:223-224)
        const T d = r.a[0].v[k];
        r.a[0].v[k] = M::mul(L.s1, d);
        r.a[1].v[k] = M::mul(L.s0, d);
      } else if constexpr (KIND == OP_BWD_NU) {  // in: dnu, g ; s0=B2 s1=TWO_OMB2 ; out: dg, dnu (cpu.cpp:261-262)
        const T d = r.a[0].v[k];
        r.a[0].v[k] = M::mul(M::mul(L.s1, r.a[1].v[k]), d);
        r.a[1].v[k] = M::mul(L.s0, d);
      } else {  // OP_BWD_UPD  in: du, u, mu' ; s0=O1 s1=C2E ; out: dmu', dnu'         (cpu.cpp:298-316)
        const T du = r.a[0].v[k], u = r.a[1].v[k], m = r.a[2].v[k];
        T dm = T(0), dn = T(0);
        if (m != T(0)) {
          const T q = M::div(u, m);
          const T den = M::mul(q, L.s0);
          dm = M::mul(du, q);
          const T t = M::mul(M::mul(-du, u), den);
          dn = M::dnu_tail(t, L.s1, den);
        }
        r.a[0].v[k] = dm;
        r.a[1].v[k] = dn;
      }
    }
  }
  template <int P>
  static __device__ __forceinline__ void store(const Regs<P> &r, const Leaf &L, long long i) {
#pragma unroll
    for (int a = 0; a < NOUT; ++a) store_pack(L.out[a] + i, r.a[a]);
  }
};

// ------------------------------------------------------------------------------------------------
// the streaming engine
// ------------------------------------------------------------------------------------------------
template <class Op, int P, int U, int BLOCK>
__device__ __forceinline__ void process_chunk(const typename Op::Leaf &L, long long off, int len) {
  constexpr int TILE = BLOCK * P * U;
  const int tid = threadIdx.x;
  long long base = off;
  int rem = len;
  // full tiles: U front-batched vector loads per array, no bounds checks
  for (; rem >= TILE; rem -= TILE, base += TILE) {
    typename Op::template Regs<P> r[U];
#pragma unroll
    for (int j = 0; j < U; ++j) Op::template load<P>(r[j], L, base + (long long)(j * BLOCK + tid) * P);
#pragma unroll
    for (int j = 0; j < U; ++j) Op::template compute<P>(r[j], L);
#pragma unroll
    for (int j = 0; j < U; ++j) Op::template store<P>(r[j], L, base + (long long)(j * BLOCK + tid) * P);
  }
  // ragged end of the leaf: whole vectors, then < P single elements
  const int nvec = rem / P;
  for (int v = tid; v < nvec; v += BLOCK) {
    typename Op::template Regs<P> r;
    Op::template load<P>(r, L, base + (long long)v * P);
    Op::template compute<P>(r, L);
    Op::template store<P>(r, L, base + (long long)v * P);
  }
  if constexpr (P > 1) {
    const int tail = rem - nvec * P;
    if (tid < tail) {
      typename Op::template Regs<1> r;
      const long long i = base + (long long)nvec * P + tid;
      Op::template load<1>(r, L, i);
      Op::template compute<1>(r, L);
      Op::template store<1>(r, L, i);
    }
  }
}

// Programmatic dependent launch (PDL): every launch carries cudaLaunchAttributeProgrammaticStreamSerialization
// (b200_adam.cu:launch_table).  `griddepcontrol.wait` blocks until the grids this one depends on have completed and
// flushed -- it precedes every global access, so correctness does not depend on what the previous kernel was --
// and `griddepcontrol.launch_dependents` lets the NEXT grid's CTAs become resident as soon as all of this grid's CTAs
// have started, i.e. while the last wave drains: the ~3 us launch gap between the back-to-back kernels of an unroll
// disappears (profiles/r01_pdl_experiment.md: +4 % at 2^24 elements; r02_leaf_sweep.md).  Both are no-ops for a grid
// launched without the attribute.
__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <class Op, int P, int U, int BLOCK, int MAXL>
__global__ void __launch_bounds__(BLOCK) stream_kernel(const __grid_constant__ Table<Op, MAXL> tab) {
  pdl_wait_then_release();
  for (int c = blockIdx.x; c < tab.total_chunks; c += gridDim.x) {
    int l = 0;
    if constexpr (MAXL > 1) {
      // largest l with chunk_start[l] <= c (uniform across the CTA; constant-bank reads)
      int lo = 0, hi = tab.num_leaves;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab.chunk_start[mid] <= c) lo = mid; else hi = mid;
      }
      l = lo;
    }
    const typename Op::Leaf L = Op::leaf(tab, l);
    const long long off = (long long)(c - tab.chunk_start[l]) * tab.chunk_elems;
    const long long left = tab.n[l] - off;
    const int len = left < (long long)tab.chunk_elems ? (int)left : tab.chunk_elems;
    if (Op::template aligned<P>(L)) {
      process_chunk<Op, P, U, BLOCK>(L, off, len);
    } else {
      process_chunk<Op, 1, 1, BLOCK>(L, off, len);  // unaligned views: element-wise path
    }
  }
}

}  // namespace b200adam
