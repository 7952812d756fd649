/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("O2") of the reference's accelerated Adam operator.
 *
 * This file is the parity checker for the CUDA kernels in torchopt_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here bit-for-bit against
 * (a) tests/golden/*.npz, produced by the reference's own compiled adam_op (oracle/_ref/_C.so, built
 * from the unmodified sources by oracle/build_ref.py; generator: tests/golden/make_golden.py) and
 * (b) oracle/_ref/_C.so itself whenever it is present.
 *
 * Every function cites the reference loop it restates (paths relative to /root/reference).
 * Arithmetic contract: each * + - / sqrt is ONE correctly rounded IEEE operation
 * in the tensor dtype T, in source order, with no FMA contraction (compile with -ffp-contract=off and
 * no -march/-ffast-math); the host scalars are derived in double and rounded to T once.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared adam_oracle.c -o liboracle_adam.so -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Threading policy of the reference: src/adam_op/adam_op_impl_cpu.cpp:30,43-45
 * (serial when n <= 1000, else min(n/1000, omp_get_num_procs()) threads). */
static int oracle_threads(size_t n) {
#ifdef _OPENMP
  if (n <= 1000) return 1;
  size_t t = n / 1000, p = (size_t)omp_get_num_procs();
  return (int)(t < p ? t : p);
#else
  (void)n;
  return 1;
#endif
}

int oracle_adam_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* Host-side scalar preparation.
 * forward_/forward_updates: adam_op_impl_cpu.cpp:72-74,190-192 ; backward_updates: :327-329. */
static double inv_one_minus_pow(double b, uint64_t count) { return 1 / (1 - pow(b, (double)count)); }
static double one_minus_pow(double b, uint64_t count) { return 1 - pow(b, (double)count); }
static double inv_one_minus_pow_eps(double b, double eps_root, uint64_t count) {
  return 1 / (1 - pow(b, (double)count) + eps_root);
}

#define DEFINE_ORACLE(T, SUF, SQRT)                                                                       \
  /* adamForwardInplaceCPUKernel: adam_op_impl_cpu.cpp:32-62 (loop body :46-60) */                        \
  void oracle_adam_forward_inplace_##SUF(T *updates, T *mu, T *nu, size_t n, double b1d, double b2d,      \
                                         double epsd, double eps_rootd, uint64_t count) {                 \
    const T b1 = (T)b1d, b2 = (T)b2d, eps = (T)epsd, eps_root = (T)eps_rootd;                             \
    const T c1 = (T)inv_one_minus_pow(b1d, count), c2 = (T)inv_one_minus_pow(b2d, count);                 \
    const T omb1 = 1 - b1, omb2 = 1 - b2;                                                                 \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T g = updates[i], m = mu[i], v = nu[i];                                                       \
      const T m_out = (T)((T)(b1 * m) + (T)(omb1 * g));                                                   \
      const T v_out = (T)((T)(b2 * v) + (T)((T)(omb2 * g) * g));                                          \
      const T m_hat = (T)(m_out * c1);                                                                    \
      const T v_hat = (T)(v_out * c2);                                                                    \
      const T u_out = (T)(m_hat / (T)((T)SQRT((T)(v_hat + eps_root)) + eps));                             \
      mu[i] = m_out;                                                                                      \
      nu[i] = v_out;                                                                                      \
      updates[i] = u_out;                                                                                 \
    }                                                                                                     \
  }                                                                                                       \
  /* adamForwardMuCPUKernel: adam_op_impl_cpu.cpp:93-108 */                                               \
  void oracle_adam_forward_mu_##SUF(const T *updates, const T *mu, T *mu_out, size_t n, double b1d) {     \
    const T b1 = (T)b1d;                                                                                  \
    const T omb1 = 1 - b1;                                                                                \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) mu_out[i] = (T)((T)(b1 * mu[i]) + (T)(omb1 * updates[i]));             \
  }                                                                                                       \
  /* adamForwardNuCPUKernel: adam_op_impl_cpu.cpp:127-143 ; (1-b2)*g*g associates left to right */        \
  void oracle_adam_forward_nu_##SUF(const T *updates, const T *nu, T *nu_out, size_t n, double b2d) {     \
    const T b2 = (T)b2d;                                                                                  \
    const T omb2 = 1 - b2;                                                                                \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T g = updates[i];                                                                             \
      nu_out[i] = (T)((T)(b2 * nu[i]) + (T)((T)(omb2 * g) * g));                                          \
    }                                                                                                     \
  }                                                                                                       \
  /* adamForwardUpdatesCPUKernel: adam_op_impl_cpu.cpp:162-181 */                                         \
  void oracle_adam_forward_updates_##SUF(const T *new_mu, const T *new_nu, T *updates_out, size_t n,      \
                                         double b1d, double b2d, double epsd, double eps_rootd,           \
                                         uint64_t count) {                                                \
    const T eps = (T)epsd, eps_root = (T)eps_rootd;                                                       \
    const T c1 = (T)inv_one_minus_pow(b1d, count), c2 = (T)inv_one_minus_pow(b2d, count);                 \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T m_hat = (T)(new_mu[i] * c1);                                                                \
      const T v_hat = (T)(new_nu[i] * c2);                                                                \
      updates_out[i] = (T)(m_hat / (T)((T)SQRT((T)(v_hat + eps_root)) + eps));                            \
    }                                                                                                     \
  }                                                                                                       \
  /* adamBackwardMuCPUKernel: adam_op_impl_cpu.cpp:211-226 */                                             \
  void oracle_adam_backward_mu_##SUF(const T *dmu, T *dupdates_out, T *dmu_out, size_t n, double b1d) {   \
    const T b1 = (T)b1d;                                                                                  \
    const T omb1 = 1 - b1;                                                                                \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T d = dmu[i];                                                                                 \
      dupdates_out[i] = (T)(omb1 * d);                                                                    \
      dmu_out[i] = (T)(b1 * d);                                                                           \
    }                                                                                                     \
  }                                                                                                       \
  /* adamBackwardNuCPUKernel: adam_op_impl_cpu.cpp:247-264 ; 2*(1-b2)*g*dnu associates left to right */  \
  void oracle_adam_backward_nu_##SUF(const T *dnu, const T *updates, T *dupdates_out, T *dnu_out,         \
                                     size_t n, double b2d) {                                              \
    const T b2 = (T)b2d;                                                                                  \
    const T two_omb2 = (T)(2 * (T)(1 - b2));                                                              \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T d = dnu[i];                                                                                 \
      dupdates_out[i] = (T)((T)(two_omb2 * updates[i]) * d);                                              \
      dnu_out[i] = (T)(b2 * d);                                                                           \
    }                                                                                                     \
  }                                                                                                       \
  /* adamBackwardUpdatesCPUKernel: adam_op_impl_cpu.cpp:286-317.  The literal 0.5 is a C double, so the   \
   * product is promoted to double from that factor on and rounded to T once at the store (:315-316). */  \
  void oracle_adam_backward_updates_##SUF(const T *dupdates, const T *updates, const T *new_mu,           \
                                          T *dnew_mu_out, T *dnew_nu_out, size_t n, double b1d,           \
                                          double b2d, double eps_rootd, uint64_t count) {                 \
    const T o1 = (T)one_minus_pow(b1d, count);                                                            \
    const T c2e = (T)inv_one_minus_pow_eps(b2d, eps_rootd, count);                                        \
    const int nt = oracle_threads(n);                                                                     \
    (void)nt;                                                                                             \
    _Pragma("omp parallel for num_threads(nt) if (n > 1000)")                                             \
    for (size_t i = 0; i < n; ++i) {                                                                      \
      const T du = dupdates[i], u = updates[i], m = new_mu[i];                                            \
      if (m == 0) {                                                                                       \
        dnew_mu_out[i] = 0;                                                                               \
        dnew_nu_out[i] = 0;                                                                               \
        continue;                                                                                         \
      }                                                                                                   \
      const T r = (T)(u / m);                                                                             \
      const T den = (T)(r * o1);                                                                          \
      dnew_mu_out[i] = (T)(du * r);                                                                       \
      const T t = (T)((T)((T)(-du) * u) * den);                                                           \
      dnew_nu_out[i] = (T)((((double)t * 0.5) * (double)c2e) * (double)den);                              \
    }                                                                                                     \
  }

DEFINE_ORACLE(float, f32, sqrtf)
DEFINE_ORACLE(double, f64, sqrt)

/* Host-scalar helpers exported so tests can pin the scalar derivation separately. */
double oracle_adam_inv_one_minus_pow(double b, uint64_t count) { return inv_one_minus_pow(b, count); }
double oracle_adam_one_minus_pow(double b, uint64_t count) { return one_minus_pow(b, count); }
double oracle_adam_inv_one_minus_pow_eps(double b, double eps_root, uint64_t count) {
  return inv_one_minus_pow_eps(b, eps_root, count);
}
