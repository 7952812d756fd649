/*
 * b200_adam.h -- C ABI of libb200adam.so: the accelerated differentiable Adam operator for NVIDIA B200
 * (sm_100a).  Plain pointers and sizes only; no torch / pybind types cross this boundary.
 *
 * This is the drop-in boundary for ONE path of metaopt/torchopt (v0.7.3): the native functions behind
 * `torchopt._C.adam_op` (reference declarations: include/adam_op/adam_op.h:30-72, dispatcher
 * src/adam_op/adam_op.cpp:32-153, CUDA launchers src/adam_op/adam_op_impl_cuda.cu:65-490).  Each entry
 * point below names the reference launcher it replaces.  The reference-side binding a maintainer would
 * add (the body of src/adam_op/adam_op.cpp calling these) is shown in INTEGRATION.md and shipped as
 * torchopt_b200/csrc/torch_shim.cpp.
 *
 * Conventions
 *  - Every array argument is a DEVICE pointer to `n` densely packed elements (element i of every argument
 *    of one call is paired, exactly like the reference's data_ptr()[i] indexing, include/utils.h:27-34).
 *  - `dtype` selects the element types:
 *        B200_ADAM_F32       all arrays float            (arithmetic: IEEE fp32, no FMA contraction)
 *        B200_ADAM_F64       all arrays double           (arithmetic: IEEE fp64, no FMA contraction)
 *        B200_ADAM_BF16_F32  gradient-side arrays (updates / new_updates / dupdates in and out) are
 *                            bfloat16, moment-side arrays (mu, nu, their grads) are float; arithmetic is
 *                            fp32 on the up-cast inputs with ONE final rounding of bf16 outputs.
 *                            Accepted by the fused, in-place and whole-step entry points (not by the six
 *                            out-of-place reference operators).
 *  - Scalars arrive as the reference's pyfloat_t=double / pyuint_t=size_t (include/common.h:23-24).  The
 *    bias corrections are derived on the host exactly as the reference does (std::pow in double, one
 *    rounding to the tensor dtype: adam_op_impl_cuda.cu:73-75,253-255,452-454).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is enqueued
 *    asynchronously on it; nothing synchronises.  The CUDA *current device* must own the pointers.
 *  - Return value: 0 on success; a positive value is a cudaError_t; negative values are B200_ADAM_E_*.
 *    n == 0 (or num_leaves == 0) is a successful no-op (the reference CPU path returns an empty tensor;
 *    its CUDA path divides by zero on the host, adam_op_impl_cuda.cu:79-80).
 *  - Results are bit-identical to the reference's CPU build (adam_op_impl_cpu.cpp) for F32 and F64,
 *    including the mu'==0 branch, the fp64 tail of backward_updates, tails of any length and unaligned
 *    pointers (which take a slower scalar path).
 *  - Thread safety: all entry points are re-entrant; the library keeps no per-call state.
 */
#ifndef B200_ADAM_H_
#define B200_ADAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 1: round 1.  2: additive -- b200_adam_set_pdl, b200_adam_step_forward_capturable / _backward_capturable,
 * b200_adam_bias_table_len_ex / _ex, b200_adam_prepare_fused / _launch_prepared / _release_prepared (nothing of version 1
 * changed its signature or meaning). */
#define B200_ADAM_ABI_VERSION 2

enum b200_adam_dtype { B200_ADAM_F32 = 0, B200_ADAM_F64 = 1, B200_ADAM_BF16_F32 = 2 };

enum b200_adam_error {
  B200_ADAM_OK = 0,
  B200_ADAM_E_DTYPE = -1,    /* dtype not supported by this entry point */
  B200_ADAM_E_ARG = -2,      /* NULL pointer where an array is required, negative size, ... */
  B200_ADAM_E_NO_DEVICE = -3 /* the current CUDA device is not compute capability 10.x (B200); no fallback exists */
};

/* ---- the seven operators of torchopt._C.adam_op (single tensor) -------------------------------- */

/* forward_ : replaces adamForwardInplaceCUDA (adam_op_impl_cuda.cu:65-112).  In place: updates <- u,
 * mu <- mu', nu <- nu'.  The three pointers may alias each other (is_available() passes one tensor three
 * times, accelerated_op/__init__.py:46-47); per element all reads precede the writes mu, nu, updates. */
int b200_adam_forward_inplace(int dtype, void *updates, void *mu, void *nu, int64_t n, double b1, double b2,
                              double eps, double eps_root, uint64_t count, void *stream);

/* forward_mu : replaces adamForwardMuCUDA (adam_op_impl_cuda.cu:135-165).  mu_out = b1*mu + (1-b1)*updates */
int b200_adam_forward_mu(int dtype, const void *updates, const void *mu, void *mu_out, int64_t n, double b1,
                         void *stream);

/* forward_nu : replaces adamForwardNuCUDA (:189-219).  nu_out = b2*nu + ((1-b2)*updates)*updates */
int b200_adam_forward_nu(int dtype, const void *updates, const void *nu, void *nu_out, int64_t n, double b2,
                         void *stream);

/* forward_updates : replaces adamForwardUpdatesCUDA (:246-291).
 * updates_out = (new_mu*c1) / (sqrt(new_nu*c2 + eps_root) + eps) */
int b200_adam_forward_updates(int dtype, const void *new_mu, const void *new_nu, void *updates_out, int64_t n,
                              double b1, double b2, double eps, double eps_root, uint64_t count, void *stream);

/* backward_mu : replaces adamBackwardMuCUDA (:314-346).  dupdates_out = (1-b1)*dmu ; dmu_out = b1*dmu */
int b200_adam_backward_mu(int dtype, const void *dmu, void *dupdates_out, void *dmu_out, int64_t n, double b1,
                          void *stream);

/* backward_nu : replaces adamBackwardNuCUDA (:371-405).
 * dupdates_out = ((2*(1-b2))*updates)*dnu ; dnu_out = b2*dnu */
int b200_adam_backward_nu(int dtype, const void *dnu, const void *updates, void *dupdates_out, void *dnu_out,
                          int64_t n, double b2, void *stream);

/* backward_updates : replaces adamBackwardUpdatesCUDA (:444-490); CPU semantics (adam_op_impl_cpu.cpp:298-316):
 * every element is processed (`continue`, not `return`, on new_mu == 0) and no tail is dropped. */
int b200_adam_backward_updates(int dtype, const void *dupdates, const void *updates, const void *new_mu,
                               void *dnew_mu_out, void *dnew_nu_out, int64_t n, double b1, double b2,
                               double eps_root, uint64_t count, void *stream);

/* ---- fused, multi-tensor entry points (new; one launch covers all pytree leaves) ----------------- */

/* Differentiable forward of one Adam step for `num_leaves` leaves in ONE launch (24 B/element fp32):
 *   mu_out = forward_mu(g, mu) ; nu_out = forward_nu(g, nu) ; u_out = forward_updates(mu_out, nu_out, count)
 * i.e. what AdamOp.__call__(inplace=False) computes with three launches per leaf
 * (accelerated_op/adam_op.py:168-180).  Arrays of `num_leaves` device pointers / sizes / counts live in
 * HOST memory and are consumed before the call returns.  Outputs must not alias inputs (use
 * b200_adam_forward_inplace_mt for that).  Leaves with n[i] == 0 are skipped. */
int b200_adam_fused_forward(int dtype, int64_t num_leaves, const void *const *g, const void *const *mu,
                            const void *const *nu, void *const *mu_out, void *const *nu_out, void *const *u_out,
                            const int64_t *n, const uint64_t *count, double b1, double b2, double eps,
                            double eps_root, void *stream);

/* Multi-tensor forward_ (in place over g/mu/nu of every leaf; aliasing allowed as in forward_). */
int b200_adam_forward_inplace_mt(int dtype, int64_t num_leaves, void *const *updates, void *const *mu,
                                 void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2,
                                 double eps, double eps_root, void *stream);

/* Differentiable backward of one Adam step for `num_leaves` leaves in ONE launch (36 B/element fp32).
 * Equals UpdatesOp.backward -> engine accumulation with the incoming moment gradients -> MuOp.backward and
 * NuOp.backward -> engine accumulation of the two dupdates (accelerated_op/adam_op.py:105-121,53-60,79-86):
 *   (dmu', dnu')   = backward_updates(du, u, new_mu, count)          [skipped when du[i] == NULL]
 *   dmu_tot        = dmu_ext + dmu'   (either may be absent)         dnu_tot likewise
 *   (dg_mu, dmu)   = backward_mu(dmu_tot)                            [absent when dmu_tot is absent]
 *   (dg_nu, dnu)   = backward_nu(dnu_tot, g)
 *   dg             = dg_mu + dg_nu    (0 when both absent)
 * Per leaf, du[i], dmu_ext[i], dnu_ext[i] may be NULL (= no gradient flows in), and dg[i], dmu[i], dnu[i] may
 * be NULL (= that output is not needed and is not written).  When an output is requested but its value is
 * "absent" (no contributing gradient) zeros are written.  u, new_mu are required when du[i] != NULL; g is
 * required when dnu_tot exists and dg[i] != NULL. */
int b200_adam_fused_backward(int dtype, int64_t num_leaves, const void *const *du, const void *const *u,
                             const void *const *new_mu, const void *const *g, const void *const *dmu_ext,
                             const void *const *dnu_ext, void *const *dg, void *const *dmu, void *const *dnu,
                             const int64_t *n, const uint64_t *count, double b1, double b2, double eps_root,
                             void *stream);

/* ---- whole optimizer step: recipe + Adam update + scale(step_size) + apply_updates in one launch ------------------
 * Replaces, for all leaves at once, the chains of alias/adam.py:116-137 and alias/adamw.py:140-151 followed by
 * apply_updates (update.py:78-79):
 *   g1 = maximize ? -g : g                                  flip_sign (alias/utils.py:133-191)
 *   g2 = g1 + wd_l2 * p                 [wd_l2 != 0]        L2 weight decay, `adam(weight_decay=)` (alias/utils.py:116-126)
 *   mu', nu', u = Adam(g2, mu, nu)                          as b200_adam_fused_forward
 *   u2 = u + wd_decoupled * p           [wd_decoupled != 0] `adamw` (transform/add_decayed_weights.py:237-247)
 *   us = u2 * step_size                                     scale(-lr) (transform/scale.py:103), step_size = -lr
 *   p_out = p + us                                          apply_updates
 * All three dtype families are accepted; with B200_ADAM_BF16_F32 the parameter-side arrays (p, g, us and their
 * gradients) are bfloat16, the moments float, and every tensor the chain would materialise in bf16 (g2, u, u2, us, p')
 * is rounded to bf16 at that point, so the result equals the chain evaluated with torch's bf16 ops.
 * `x.add(y, alpha=a)` is evaluated like ATen does on CUDA: one rounding, fma(a, y, x).  At most one of wd_l2 /
 * wd_decoupled may be non-zero.  u is never stored (the backward recomputes it bit-identically from mu', nu');
 * `us_out` may be NULL, or hold NULL entries, when the scaled update tensor is not needed.
 * 28 B/element fp32 (32 with us_out). */
int b200_adam_step_forward(int dtype, int64_t num_leaves, const void *const *p, const void *const *g,
                           const void *const *mu, const void *const *nu, void *const *p_out, void *const *mu_out,
                           void *const *nu_out, void *const *us_out, const int64_t *n, const uint64_t *count,
                           double b1, double b2, double eps, double eps_root, double step_size, double wd_l2,
                           double wd_decoupled, int maximize, void *stream);

/* In-place, non-differentiable variant (Optimizer.step, optim/base.py:99-124): p += us, g <- us, mu <- mu', nu <- nu'. */
int b200_adam_step_inplace(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                           void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2, double eps,
                           double eps_root, double step_size, double wd_l2, double wd_decoupled, int maximize,
                           void *stream);

/* Same, with a choice about the gradient arrays: overwrite_grads != 0 is b200_adam_step_inplace (the reference's side
 * effect: with inplace=True its chain turns every gradient tensor into the scaled update, transform/scale.py:98-101,
 * update.py:76-77); overwrite_grads == 0 leaves g untouched -- p, mu, nu are still updated in place -- which saves the
 * 4 B/element store: 28 instead of 32 B/element fp32.  For optimizers that own their gradient buffer (FlatAdam). */
int b200_adam_step_inplace_ex(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                              void *const *nu, const int64_t *n, const uint64_t *count, double b1, double b2,
                              double eps, double eps_root, double step_size, double wd_l2, double wd_decoupled,
                              int maximize, int overwrite_grads, void *stream);

/* Capturable in-place step: for an optimizer step that is recorded in a CUDA graph together with the rest of a training
 * iteration (the reference cannot be recorded at all: its binding reads `count` back to the host on every call,
 * accelerated_op/adam_op.py:98-101).  Nothing about the step count is baked into the launch:
 *   count_dev[i]   DEVICE pointer to the int64 step count of leaf i, already incremented by the caller on the device;
 *   c1/c2_table    DEVICE arrays of `table_len` bias corrections 1/(1-b^count) for count = 1..table_len in the
 *                  arithmetic type (float for F32 and BF16_F32, double for F64), produced on the host by
 *                  b200_adam_bias_table -- the same std::pow expression and single rounding as every other entry point,
 *                  hence bit-identical results.  b200_adam_bias_table_len gives the length after which both corrections
 *                  are exactly 1 (16 628 entries for b2 = 0.999 in fp32); larger counts use the last entry.
 * Otherwise identical to b200_adam_step_inplace_ex. */
int64_t b200_adam_bias_table_len(int dtype, double b1, double b2);
int b200_adam_bias_table(int dtype, double b1, double b2, int64_t len, void *c1_host, void *c2_host);
int b200_adam_step_inplace_capturable(int dtype, int64_t num_leaves, void *const *p, void *const *g, void *const *mu,
                                      void *const *nu, const int64_t *n, const int64_t *const *count_dev,
                                      const void *c1_table_dev, const void *c2_table_dev, int64_t table_len, double b1,
                                      double b2, double eps, double eps_root, double step_size, double wd_l2,
                                      double wd_decoupled, int maximize, int overwrite_grads, void *stream);

/* Backward of b200_adam_step_forward.  Incoming per leaf (each may be NULL): dp = gradient of p_out, dus = gradient of
 * us_out (the array argument itself may be NULL), dmu_ext / dnu_ext = gradients of mu_out / nu_out.
 *   du2 = (dp + dus) * step_size ; then exactly b200_adam_fused_backward with u recomputed from new_mu, new_nu and the
 *   gradient that entered the moments re-derived as g2 = fma(wd_l2, p, +-g)  (`p` is required only when wd_l2 != 0);
 *   dg = maximize ? -dg2 : dg2.
 * Gradient w.r.t. the parameter: without weight decay it is dp itself and nothing is written (pass dp_out = NULL);
 * with weight decay dp_out[i] = dp + wd * (du2 | dg2) is written when dp_out[i] != NULL (+4 B/element).
 * 36 B/element fp32. */
int b200_adam_step_backward(int dtype, int64_t num_leaves, const void *const *dp, const void *const *dus,
                            const void *const *new_mu, const void *const *new_nu, const void *const *g,
                            const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg, void *const *dmu,
                            void *const *dnu, const void *const *p, void *const *dp_out, const int64_t *n,
                            const uint64_t *count, double b1, double b2, double eps, double eps_root, double step_size,
                            double wd_l2, double wd_decoupled, int maximize, void *stream);

/* Capturable DIFFERENTIABLE step (forward + backward): b200_adam_step_forward / b200_adam_step_backward with the step
 * counts in device memory and all bias corrections in device tables, so that a differentiable optimizer step over a
 * PERSISTENT state (MetaAdam / FuncAdam(capturable=True)) can be recorded in a CUDA graph and every replay is one real
 * step (the reference reads `count` back to the host on every call, accelerated_op/adam_op.py:98-101,110).
 *   count_dev[i]   DEVICE pointer to the int64 step count of leaf i (already incremented); the backward must be given
 *                  the same value the forward saw (the binding snapshots it per step);
 *   c1 / c2 tables as for b200_adam_step_inplace_capturable; the backward additionally needs o1 = 1-b1^count and
 *                  c2e = 1/(1-b2^count+eps_root).  b200_adam_bias_table_len_ex gives the length from which ALL FOUR have
 *                  reached their limit value, b200_adam_bias_table_ex fills the four host tables with the same std::pow
 *                  expressions (one rounding) as the host-scalar entry points: results are bit-identical to those. */
int64_t b200_adam_bias_table_len_ex(int dtype, double b1, double b2, double eps_root);
int b200_adam_bias_table_ex(int dtype, double b1, double b2, double eps_root, int64_t len, void *c1_host, void *c2_host,
                            void *o1_host, void *c2e_host);
int b200_adam_step_forward_capturable(int dtype, int64_t num_leaves, const void *const *p, const void *const *g,
                                      const void *const *mu, const void *const *nu, void *const *p_out,
                                      void *const *mu_out, void *const *nu_out, void *const *us_out, const int64_t *n,
                                      const int64_t *const *count_dev, const void *c1_table_dev, const void *c2_table_dev,
                                      int64_t table_len, double b1, double b2, double eps, double eps_root,
                                      double step_size, double wd_l2, double wd_decoupled, int maximize, void *stream);
int b200_adam_step_backward_capturable(int dtype, int64_t num_leaves, const void *const *dp, const void *const *dus,
                                       const void *const *new_mu, const void *const *new_nu, const void *const *g,
                                       const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg,
                                       void *const *dmu, void *const *dnu, const void *const *p, void *const *dp_out,
                                       const int64_t *n, const int64_t *const *count_dev, const void *c1_table_dev,
                                       const void *c2_table_dev, const void *o1_table_dev, const void *c2e_table_dev,
                                       int64_t table_len, double b1, double b2, double eps, double eps_root,
                                       double step_size, double wd_l2, double wd_decoupled, int maximize, void *stream);

/* ---- prepared launches ---------------------------------------------------------------------------------------------
 * For a caller that issues the SAME fused forward + backward pair over and over (a steady-state loop over fixed buffers
 * that is not recorded in a CUDA graph): the whole argument block of b200_adam_fused_forward / b200_adam_fused_backward is
 * copied into the library once, and each launch is then a two-argument call.  The backward of the pair reads the
 * forward's outputs (u = u_out, new_mu = mu_out).  Returns NULL on invalid arguments.  Nothing about the launches
 * themselves changes (same kernels, same results); this only removes per-call argument marshalling in bindings where
 * that costs more than the launch (ctypes: ~1.5 us of 17 converted arguments). */
void *b200_adam_prepare_fused(int dtype, int64_t num_leaves, const void *const *g, const void *const *mu,
                              const void *const *nu, void *const *mu_out, void *const *nu_out, void *const *u_out,
                              const void *const *du, const void *const *dmu_ext, const void *const *dnu_ext, void *const *dg,
                              void *const *dmu, void *const *dnu, const int64_t *n, const uint64_t *count, double b1,
                              double b2, double eps, double eps_root);
int b200_adam_launch_prepared(void *handle, int backward, void *stream);
int b200_adam_release_prepared(void *handle);

/* ---- host-buffer entry point (what an application holding HOST arrays calls; used for the e2e number) */

/* One fused differentiable forward + backward over HOST arrays of n elements (dtype F32 only):
 * inputs g, mu, nu, du, dmu_ext, dnu_ext; outputs mu_out, nu_out, u_out, dg, dmu, dnu (all host pointers,
 * ideally page-locked).  The range is cut into slices that are copied in, processed by the two fused kernels
 * and copied out on three streams so H2D, compute and D2H overlap.  Blocks until the outputs are complete.
 * `slice_elems` <= 0 selects the default slice. */
int b200_adam_host_step_f32(const float *g, const float *mu, const float *nu, const float *du,
                            const float *dmu_ext, const float *dnu_ext, float *mu_out, float *nu_out,
                            float *u_out, float *dg, float *dmu, float *dnu, int64_t n, double b1, double b2,
                            double eps, double eps_root, uint64_t count, int64_t slice_elems);

/* Frees the device staging buffers / streams cached by b200_adam_host_step_f32 (optional). */
int b200_adam_host_release(void);

/* ---- utilities ------------------------------------------------------------------------------------ */

int b200_adam_abi_version(void);

/* Human-readable text for a return code of this library (never NULL). */
const char *b200_adam_error_string(int code);

/* Number of CUDA kernels this library has launched in this process so far (for bench.py's gpu_launches). */
uint64_t b200_adam_kernel_launches(void);

/* Launch tuning.  ctas_per_sm: 0 = default, one CTA per chunk (non-persistent grid); k > 0 = persistent grid of
 * k CTAs per SM walking the chunks in grid-stride order.  tiles_per_chunk: block-tiles per chunk (0 = default, 1).
 * Process-wide; intended for the bench/tuning harness. */
int b200_adam_set_tuning(int ctas_per_sm, int tiles_per_chunk);

/* Programmatic dependent launch (default on): every kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, waits (griddepcontrol.wait) for the work before it in the stream
 * before its first global access and lets the next kernel's CTAs become resident while its own last wave drains -- the
 * launch gap between the back-to-back kernels of an unroll (~3 us each) is hidden.  enabled = 0 restores plain stream
 * serialisation.  Process-wide; results are identical either way. */
int b200_adam_set_pdl(int enabled);

/* Writes the launch geometry the library would use for an `n`-element single-leaf fused launch:
 * out[0]=grid, out[1]=block, out[2]=elements per chunk, out[3]=vector width in elements, out[4]=unroll. */
int b200_adam_query_launch(int dtype, int backward, int64_t n, int32_t out[5]);

#ifdef __cplusplus
}
#endif
#endif /* B200_ADAM_H_ */
