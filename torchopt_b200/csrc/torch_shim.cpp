// torch_shim.cpp -- the CPython module `_C` (submodule `adam_op`) over the C ABI of libb200adam.so.
//
// This is the reference-side binding: it has the Python-visible surface of the reference's pybind module
// (/root/reference/src/adam_op/adam_op.cpp:155-215, src/extension.cpp:20 -- same function names, keyword
// names incl. backward_nu's misnamed `b1`, list return types, size_t `count` accepting a 0-dim tensor through
// __index__) and the dispatcher's behaviour (/root/reference/src/adam_op/adam_op.cpp:32-153: .contiguous()
// on the incoming gradient of the three backward ops only, fresh empty_like outputs, no shape/layout checks,
// `throw std::runtime_error("Not implemented")` for devices without a kernel).  It owns nothing numeric:
// it resolves device guard + current stream, allocates outputs and passes raw pointers across the C ABI.
//
// This build has NO CPU kernels: CPU (and meta, ...) tensors raise "Not implemented".
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <torch/extension.h>
#include <torch/csrc/autograd/custom_function.h>
#include <ATen/ExpandUtils.h>

#include <array>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "b200_adam.h"

namespace py = pybind11;
using at::Tensor;
using OptTensor = std::optional<Tensor>;

namespace {

[[noreturn]] void fail(int rc, const char *what) {
  throw std::runtime_error(std::string(what) + ": " + b200_adam_error_string(rc));
}

inline void check(int rc, const char *what) {
  if (rc != 0) fail(rc, what);
}

inline void require_cuda(const Tensor &t) {
  // same contract as the reference dispatcher for a device it has no kernel for (adam_op.cpp:48)
  if (!t.device().is_cuda()) throw std::runtime_error("Not implemented");
}

// element-type family of a call: `grad_like` is an updates/dupdates-side tensor, `state_like` a moment-side one
int family(const Tensor &grad_like, const Tensor &state_like, const char *name, bool allow_mixed) {
  const auto g = grad_like.scalar_type(), s = state_like.scalar_type();
  if (g == at::kFloat && s == at::kFloat) return B200_ADAM_F32;
  if (g == at::kDouble && s == at::kDouble) return B200_ADAM_F64;
  if (allow_mixed && g == at::kBFloat16 && s == at::kFloat) return B200_ADAM_BF16_F32;
  if (g != s) {
    throw std::runtime_error(std::string("expected scalar type ") + c10::toString(g) + " but found " +
                             c10::toString(s));
  }
  // message format of the ATen dispatch macro the reference relies on (include/common.h:26-37)
  throw std::runtime_error(std::string("\"") + name + "\" not implemented for '" + c10::toString(g) + "'");
}

inline void same_type(const Tensor &a, const Tensor &b) {
  if (a.scalar_type() != b.scalar_type())
    throw std::runtime_error(std::string("expected scalar type ") + c10::toString(a.scalar_type()) +
                             " but found " + c10::toString(b.scalar_type()));
}

struct Ctx {
  c10::cuda::CUDAGuard guard;
  void *stream;
  explicit Ctx(const Tensor &t)
      : guard(t.device()), stream((void *)c10::cuda::getCurrentCUDAStream(t.device().index()).stream()) {}
};

// ---- the seven reference operators ---------------------------------------------------------------
std::array<Tensor, 3> forward_inplace(const Tensor &updates, const Tensor &mu, const Tensor &nu, double b1,
                                      double b2, double eps, double eps_root, size_t count) {
  require_cuda(updates);
  const int dt = family(updates, mu, "adamForwardInplaceCUDA", true);
  same_type(mu, nu);
  Ctx c(updates);
  check(b200_adam_forward_inplace(dt, updates.data_ptr(), mu.data_ptr(), nu.data_ptr(), updates.numel(), b1, b2,
                                  eps, eps_root, count, c.stream),
        "forward_");
  return {updates, mu, nu};
}

Tensor forward_mu(const Tensor &updates, const Tensor &mu, double b1) {
  require_cuda(updates);
  const int dt = family(updates, mu, "adamForwardMuCUDA", false);
  Ctx c(updates);
  Tensor out = at::empty_like(mu);
  check(b200_adam_forward_mu(dt, updates.data_ptr(), mu.data_ptr(), out.data_ptr(), updates.numel(), b1, c.stream),
        "forward_mu");
  return out;
}

Tensor forward_nu(const Tensor &updates, const Tensor &nu, double b2) {
  require_cuda(updates);
  const int dt = family(updates, nu, "adamForwardNuCUDA", false);
  Ctx c(updates);
  Tensor out = at::empty_like(nu);
  check(b200_adam_forward_nu(dt, updates.data_ptr(), nu.data_ptr(), out.data_ptr(), updates.numel(), b2, c.stream),
        "forward_nu");
  return out;
}

Tensor forward_updates(const Tensor &new_mu, const Tensor &new_nu, double b1, double b2, double eps,
                       double eps_root, size_t count) {
  require_cuda(new_mu);
  const int dt = family(new_mu, new_nu, "adamForwardUpdatesCUDA", false);
  Ctx c(new_mu);
  Tensor out = at::empty_like(new_mu);
  check(b200_adam_forward_updates(dt, new_mu.data_ptr(), new_nu.data_ptr(), out.data_ptr(), new_mu.numel(), b1, b2,
                                  eps, eps_root, count, c.stream),
        "forward_updates");
  return out;
}

std::array<Tensor, 2> backward_mu(const Tensor &dmu_in, const Tensor &updates, const Tensor &mu, double b1) {
  require_cuda(dmu_in);
  const Tensor dmu = dmu_in.contiguous();
  const int dt = family(dmu, dmu, "adamBackwardMuCUDA", false);
  Ctx c(dmu);
  Tensor dupdates_out = at::empty_like(updates), dmu_out = at::empty_like(mu);
  same_type(dmu, dupdates_out), same_type(dmu, dmu_out);
  check(b200_adam_backward_mu(dt, dmu.data_ptr(), dupdates_out.data_ptr(), dmu_out.data_ptr(), dmu.numel(), b1,
                              c.stream),
        "backward_mu");
  return {dupdates_out, dmu_out};
}

std::array<Tensor, 2> backward_nu(const Tensor &dnu_in, const Tensor &updates, const Tensor &nu, double b2) {
  require_cuda(dnu_in);
  const Tensor dnu = dnu_in.contiguous();
  const int dt = family(dnu, updates, "adamForwardNuCUDA" /* sic: adam_op_impl_cuda.cu:382 */, false);
  Ctx c(dnu);
  Tensor dupdates_out = at::empty_like(updates), dnu_out = at::empty_like(nu);
  same_type(dnu, dnu_out);
  check(b200_adam_backward_nu(dt, dnu.data_ptr(), updates.data_ptr(), dupdates_out.data_ptr(), dnu_out.data_ptr(),
                              dnu.numel(), b2, c.stream),
        "backward_nu");
  return {dupdates_out, dnu_out};
}

std::array<Tensor, 2> backward_updates(const Tensor &dupdates_in, const Tensor &updates, const Tensor &new_mu,
                                       const Tensor &new_nu, double b1, double b2, double eps_root, size_t count) {
  require_cuda(dupdates_in);
  const Tensor dupdates = dupdates_in.contiguous();
  const int dt = family(dupdates, updates, "adamBackwardUpdatesCUDA", false);
  same_type(dupdates, new_mu);
  Ctx c(dupdates);
  Tensor dmu_out = at::empty_like(new_mu), dnu_out = at::empty_like(new_nu);
  same_type(dupdates, dnu_out);
  check(b200_adam_backward_updates(dt, dupdates.data_ptr(), updates.data_ptr(), new_mu.data_ptr(), dmu_out.data_ptr(),
                                   dnu_out.data_ptr(), dupdates.numel(), b1, b2, eps_root, count, c.stream),
        "backward_updates");
  return {dmu_out, dnu_out};
}

// ---- leaf layout (multi-tensor entry points only; the seven reference operators keep the reference's raw behaviour) --
// The kernels walk every array of a leaf linearly (data_ptr()[i], include/utils.h:27-34), so the tensors of ONE leaf
// must pair element i with element i: non-overlapping, dense, identical strides.  Parameters and their zeros_like
// moments satisfy that by construction, but gradients handed back by autograd need not (an expanded stride-0 gradient
// of sum(), a transposed or channels_last view): walking those linearly reads out of bounds or pairs the wrong
// elements.  `conform` returns `t` itself when it already has the donor's layout and otherwise ONE strided copy of it.
inline Tensor dense_donor(const Tensor &t) { return t.is_non_overlapping_and_dense() ? t : t.contiguous(); }

inline bool conforms(const Tensor &t, const Tensor &donor) {
  if (t.sizes() == donor.sizes()) return t.numel() <= 1 || t.strides() == donor.strides();
  // different shapes with equal numel: flat storage-order pairing like the reference -- fine if both are row-major
  return t.numel() == donor.numel() && t.is_contiguous() && donor.is_contiguous();
}

inline Tensor conform(const Tensor &t, const Tensor &donor) {  // `donor` is non-overlapping and dense
  if (conforms(t, donor)) return t;
  if (t.sizes() == donor.sizes()) {
    Tensor out = at::empty_strided(donor.sizes(), donor.strides(), t.options());
    out.copy_(t);
    return out;
  }
  if (t.numel() != donor.numel())
    throw std::runtime_error("torchopt_b200: the tensors of one leaf have different numbers of elements");
  return t.contiguous();
}

inline void require_inplace_layout(const Tensor &t, const Tensor &donor, const char *what) {
  if (!donor.is_non_overlapping_and_dense() || !conforms(t, donor))
    throw std::runtime_error(std::string("torchopt_b200: in-place step: ") + what +
                             " must be non-overlapping, dense and laid out like the leaf's other tensors");
}

// ---- flat output allocation -------------------------------------------------------
// The multi-tensor entry points return up to 4 tensors per leaf.  Allocating them one by one costs ~1.4 us of
// caching-allocator work each (186 allocations = ~255 us for a ResNet-18-sized pytree, 4x the kernel time).
// FlatAlloc hands out every output of a call from ONE buffer per (device, dtype): each output has its donor's sizes
// and (for dense donors) strides -- exactly what empty_like gives -- and starts on a 128-byte boundary.
// The outputs are independent tensors sharing a storage (Tensor::set_), not autograd views of each other.
// Only SMALL outputs are pooled (< kPoolMaxBytes = 256 KiB each): a pooled output keeps the whole pool alive, which is harmless
// for the many tiny leaves the pool exists for (41 of ResNet-18's 62 leaves are < 4096 elements) but would let one
// surviving 64-element tensor pin hundreds of MB if large leaves shared its storage.  Large outputs get their own
// allocation (the allocator cost is noise next to their kernel time).
class FlatAlloc {
 public:
  static constexpr int64_t kPoolMaxBytes = 256 << 10;
  size_t request(const Tensor &donor) {
    reqs_.push_back(&donor);
    return reqs_.size() - 1;
  }
  void commit() {
    out_.resize(reqs_.size());
    if (reqs_.size() <= 3) {  // single-leaf calls: plain empty_like is cheaper than the bookkeeping
      for (size_t i = 0; i < reqs_.size(); ++i) out_[i] = at::empty_like(*reqs_[i]);
      return;
    }
    struct Group {
      int64_t total = 0;
      std::vector<std::pair<size_t, int64_t>> items;  // (request index, element offset)
    };
    std::map<std::tuple<int, int, int>, Group> groups;
    for (size_t i = 0; i < reqs_.size(); ++i) {
      const Tensor &d = *reqs_[i];
      if (d.numel() * (int64_t)d.element_size() >= kPoolMaxBytes) {
        out_[i] = at::empty_like(d);
        continue;
      }
      const int64_t align = 128 / (int64_t)d.element_size();
      Group &gr = groups[{(int)d.device().type(), (int)d.device().index(), (int)d.scalar_type()}];
      gr.total = (gr.total + align - 1) / align * align;
      gr.items.emplace_back(i, gr.total);
      gr.total += d.numel();
    }
    for (auto &kv : groups) {
      const Tensor &first = *reqs_[kv.second.items.front().first];
      Tensor flat = at::empty({kv.second.total}, first.options().memory_format(c10::nullopt));
      for (auto &it : kv.second.items) {
        const Tensor &d = *reqs_[it.first];
        // build the TensorImpl directly (no dispatcher round trips): shares flat's storage, own sizes/strides/offset
        auto impl = c10::make_intrusive<c10::TensorImpl>(c10::Storage(flat.storage()), flat.key_set(), flat.dtype());
        impl->set_storage_offset(it.second);
        if (d.is_non_overlapping_and_dense()) {
          impl->set_sizes_and_strides(d.sizes(), d.strides());
        } else {
          impl->set_sizes_and_strides(d.sizes(), at::infer_dense_strides(d.sizes(), d.strides()));
        }
        out_[it.first] = Tensor(std::move(impl));
      }
    }
  }
  Tensor &operator[](size_t i) { return out_[i]; }

 private:
  std::vector<const Tensor *> reqs_;
  std::vector<Tensor> out_;
};

// ---- fused multi-tensor operators (new) ------------------------------------------------------------
struct GroupKey {
  int device, dtype;
  bool operator<(const GroupKey &o) const { return std::tie(device, dtype) < std::tie(o.device, o.dtype); }
};

// (new_mu, new_nu, new_updates) for every leaf; leaves whose `updates` is None yield (mu, nu, None), like
// AdamOp.__call__ (accelerated_op/adam_op.py:148-149).  One launch per (device, dtype family).
std::tuple<std::vector<Tensor>, std::vector<Tensor>, std::vector<OptTensor>> fused_forward(
    const std::vector<OptTensor> &updates, const std::vector<Tensor> &mu, const std::vector<Tensor> &nu, double b1,
    double b2, double eps, double eps_root, const std::vector<size_t> &count, bool inplace) {
  const size_t L = updates.size();
  if (mu.size() != L || nu.size() != L || count.size() != L)
    throw std::runtime_error("fused_forward: updates, mu, nu and count must have the same number of leaves");
  std::vector<Tensor> new_mu(L), new_nu(L);
  std::vector<OptTensor> new_u(L);
  std::map<GroupKey, std::vector<size_t>> groups;
  for (size_t i = 0; i < L; ++i) {
    if (!updates[i].has_value()) {
      new_mu[i] = mu[i], new_nu[i] = nu[i];
      continue;
    }
    const Tensor &g = *updates[i];
    require_cuda(g);
    const int dt = family(g, mu[i], inplace ? "adamForwardInplaceCUDA" : "adamForwardMuCUDA", true);
    same_type(mu[i], nu[i]);
    groups[GroupKey{(int)g.device().index(), dt}].push_back(i);
  }
  for (auto &kv : groups) {
    const auto &idx = kv.second;
    const size_t G = idx.size();
    std::vector<const void *> pg(G), pmu(G), pnu(G);
    std::vector<void *> omu(G), onu(G), ou(G);
    std::vector<int64_t> n(G);
    std::vector<uint64_t> cnt(G);
    Ctx c(*updates[idx[0]]);
    // every tensor of a leaf in ONE layout (the first moment's): see `conform`
    std::vector<Tensor> cg(G), cmu(G), cnu(G);
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      if (inplace) {
        require_inplace_layout(nu[i], mu[i], "nu");
        cmu[k] = mu[i], cnu[k] = nu[i];
      } else {
        cmu[k] = dense_donor(mu[i]), cnu[k] = conform(nu[i], cmu[k]);
      }
      cg[k] = conform(*updates[i], cmu[k]);
    }
    FlatAlloc fa;
    if (!inplace) {
      for (size_t k = 0; k < G; ++k) fa.request(cmu[k]), fa.request(cnu[k]), fa.request(cg[k]);
      fa.commit();
    }
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      if (inplace) {
        new_mu[i] = mu[i], new_nu[i] = nu[i], new_u[i] = cg[k];
      } else {
        new_mu[i] = fa[3 * k], new_nu[i] = fa[3 * k + 1], new_u[i] = fa[3 * k + 2];
      }
      pg[k] = cg[k].data_ptr(), pmu[k] = cmu[k].data_ptr(), pnu[k] = cnu[k].data_ptr();
      omu[k] = new_mu[i].data_ptr(), onu[k] = new_nu[i].data_ptr(), ou[k] = new_u[i]->data_ptr();
      n[k] = cg[k].numel(), cnt[k] = count[i];
    }
    if (inplace) {
      check(b200_adam_forward_inplace_mt(kv.first.dtype, (int64_t)G, ou.data(), omu.data(), onu.data(), n.data(),
                                         cnt.data(), b1, b2, eps, eps_root, c.stream),
            "forward_ (multi-tensor)");
      for (size_t k = 0; k < G; ++k) {  // a gradient that had to be re-laid-out: write the update back into it
        const Tensor &g = *updates[idx[k]];
        if (!cg[k].is_same(g)) const_cast<Tensor &>(g).copy_(cg[k]);
        new_u[idx[k]] = g;
      }
    } else {
      check(b200_adam_fused_forward(kv.first.dtype, (int64_t)G, pg.data(), pmu.data(), pnu.data(), omu.data(),
                                    onu.data(), ou.data(), n.data(), cnt.data(), b1, b2, eps, eps_root, c.stream),
            "fused_forward");
    }
  }
  return {new_mu, new_nu, new_u};
}

// Per leaf i: gradients (dupdates[i], dmu_ext[i], dnu_ext[i]) of (new_updates, new_mu, new_nu) -- any may be
// None -- and the saved (new_updates[i], new_mu[i], g[i]).  need_* say which input gradients autograd wants.
// Returns (dg, dmu, dnu) lists with None where not needed or where no gradient flows at all.
std::tuple<std::vector<OptTensor>, std::vector<OptTensor>, std::vector<OptTensor>> fused_backward(
    const std::vector<OptTensor> &dupdates, const std::vector<OptTensor> &dmu_ext,
    const std::vector<OptTensor> &dnu_ext, const std::vector<Tensor> &new_updates, const std::vector<Tensor> &new_mu,
    const std::vector<Tensor> &g, const std::vector<bool> &need_dg, const std::vector<bool> &need_dmu,
    const std::vector<bool> &need_dnu, double b1, double b2, double eps_root, const std::vector<size_t> &count) {
  const size_t L = g.size();
  if (dupdates.size() != L || dmu_ext.size() != L || dnu_ext.size() != L || new_updates.size() != L ||
      new_mu.size() != L || need_dg.size() != L || need_dmu.size() != L || need_dnu.size() != L || count.size() != L)
    throw std::runtime_error("fused_backward: all per-leaf lists must have the same length");
  std::vector<OptTensor> dg(L), dmu(L), dnu(L);
  std::vector<long> slot_dg(L, -1), slot_dmu(L, -1), slot_dnu(L, -1);
  FlatAlloc fa;
  // every array of a leaf in the layout of its saved new_mu (`conform`: incoming gradients may be expanded / permuted
  // views; the reference only makes them row-major, adam_op.cpp:107,123,144, which mis-pairs a channels_last leaf)
  std::vector<Tensor> cdu(L), cme(L), cne(L), cu(L), cm(L), cg(L);  // kept alive until after the launch
  std::map<GroupKey, std::vector<size_t>> groups;
  for (size_t i = 0; i < L; ++i) {
    const bool hdu = dupdates[i].has_value(), hme = dmu_ext[i].has_value(), hne = dnu_ext[i].has_value();
    if (!hdu && !hme && !hne) continue;  // nothing flows: all three stay None
    const bool wg = need_dg[i] && (hdu || hme || hne), wm = need_dmu[i] && (hdu || hme), wn = need_dnu[i] && (hdu || hne);
    if (!wg && !wm && !wn) continue;
    require_cuda(g[i]);
    const int dt = family(g[i], new_mu[i], "adamBackwardUpdatesCUDA", true);
    cm[i] = dense_donor(new_mu[i]), cu[i] = conform(new_updates[i], cm[i]), cg[i] = conform(g[i], cm[i]);
    if (hdu) cdu[i] = conform(*dupdates[i], cm[i]), same_type(cdu[i], g[i]);
    if (hme) cme[i] = conform(*dmu_ext[i], cm[i]), same_type(cme[i], new_mu[i]);
    if (hne) cne[i] = conform(*dnu_ext[i], cm[i]), same_type(cne[i], new_mu[i]);
    if (wg) slot_dg[i] = (long)fa.request(cg[i]);
    if (wm) slot_dmu[i] = (long)fa.request(cm[i]);
    if (wn) slot_dnu[i] = (long)fa.request(cm[i]);
    groups[GroupKey{(int)g[i].device().index(), dt}].push_back(i);
  }
  fa.commit();
  for (size_t i = 0; i < L; ++i) {
    if (slot_dg[i] >= 0) dg[i] = fa[slot_dg[i]];
    if (slot_dmu[i] >= 0) dmu[i] = fa[slot_dmu[i]];
    if (slot_dnu[i] >= 0) dnu[i] = fa[slot_dnu[i]];
  }
  for (auto &kv : groups) {
    const auto &idx = kv.second;
    const size_t G = idx.size();
    std::vector<const void *> pdu(G), pu(G), pm(G), pg(G), pme(G), pne(G);
    std::vector<void *> odg(G), odm(G), odn(G);
    std::vector<int64_t> n(G);
    std::vector<uint64_t> cnt(G);
    Ctx c(g[idx[0]]);
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      pdu[k] = cdu[i].defined() ? cdu[i].data_ptr() : nullptr;
      pme[k] = cme[i].defined() ? cme[i].data_ptr() : nullptr;
      pne[k] = cne[i].defined() ? cne[i].data_ptr() : nullptr;
      pu[k] = cu[i].data_ptr(), pm[k] = cm[i].data_ptr(), pg[k] = cg[i].data_ptr();
      odg[k] = dg[i] ? dg[i]->data_ptr() : nullptr;
      odm[k] = dmu[i] ? dmu[i]->data_ptr() : nullptr;
      odn[k] = dnu[i] ? dnu[i]->data_ptr() : nullptr;
      n[k] = cg[i].numel(), cnt[k] = count[i];
    }
    check(b200_adam_fused_backward(kv.first.dtype, (int64_t)G, pdu.data(), pu.data(), pm.data(), pg.data(), pme.data(),
                                   pne.data(), odg.data(), odm.data(), odn.data(), n.data(), cnt.data(), b1, b2,
                                   eps_root, c.stream),
          "fused_backward");
  }
  return {dg, dmu, dnu};
}

// ---- whole optimizer step: recipe + Adam + scale(step_size) + apply_updates (fp32 / fp64 / bf16-param+fp32-state) ---
// Differentiable form: returns (new_params, new_mu, new_nu, scaled_updates?) for leaves whose gradient is not None;
// leaves without a gradient pass (param, mu, nu, None) through.  `inplace` mutates params/grads/mu/nu instead.
std::tuple<std::vector<Tensor>, std::vector<Tensor>, std::vector<Tensor>, std::vector<OptTensor>> fused_step(
    const std::vector<Tensor> &params, const std::vector<OptTensor> &grads, const std::vector<Tensor> &mu,
    const std::vector<Tensor> &nu, double b1, double b2, double eps, double eps_root, double step_size,
    const std::vector<size_t> &count, bool inplace, bool want_updates, double wd_l2, double wd_decoupled,
    bool maximize, bool keep_grads = false) {
  const size_t L = params.size();
  if (grads.size() != L || mu.size() != L || nu.size() != L || count.size() != L)
    throw std::runtime_error("fused_step: params, grads, mu, nu and count must have the same number of leaves");
  std::vector<Tensor> new_p(L), new_mu(L), new_nu(L);
  std::vector<OptTensor> new_us(L);
  std::map<GroupKey, std::vector<size_t>> groups;
  for (size_t i = 0; i < L; ++i) {
    if (!grads[i].has_value()) {
      new_p[i] = params[i], new_mu[i] = mu[i], new_nu[i] = nu[i];
      continue;
    }
    require_cuda(params[i]);
    const int dt = family(*grads[i], mu[i], "adamStepCUDA", true);
    same_type(mu[i], nu[i]), same_type(*grads[i], params[i]);
    groups[GroupKey{(int)params[i].device().index(), dt}].push_back(i);
  }
  for (auto &kv : groups) {
    const auto &idx = kv.second;
    const size_t G = idx.size();
    std::vector<const void *> pp(G), pg(G), pmu(G), pnu(G);
    std::vector<void *> op(G), omu(G), onu(G), ous(G);
    std::vector<int64_t> n(G);
    std::vector<uint64_t> cnt(G);
    Ctx c(params[idx[0]]);
    // every tensor of a leaf in ONE layout (the parameter's): see `conform`
    std::vector<Tensor> cp(G), cg(G), cmu(G), cnu(G);
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      if (inplace) {
        require_inplace_layout(mu[i], params[i], "mu"), require_inplace_layout(nu[i], params[i], "nu");
        cp[k] = params[i], cmu[k] = mu[i], cnu[k] = nu[i];
      } else {
        cp[k] = dense_donor(params[i]), cmu[k] = conform(mu[i], cp[k]), cnu[k] = conform(nu[i], cp[k]);
      }
      cg[k] = conform(*grads[i], cp[k]);
    }
    FlatAlloc fa;
    const size_t per_leaf = want_updates ? 4 : 3;
    if (!inplace) {
      for (size_t k = 0; k < G; ++k) {
        fa.request(cp[k]), fa.request(cmu[k]), fa.request(cnu[k]);
        if (want_updates) fa.request(cg[k]);
      }
      fa.commit();
    }
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      if (inplace) {
        new_p[i] = params[i], new_mu[i] = mu[i], new_nu[i] = nu[i];
        if (!keep_grads) new_us[i] = cg[k];   // reference side effect: the gradient tensor becomes the scaled update
      } else {
        new_p[i] = fa[per_leaf * k], new_mu[i] = fa[per_leaf * k + 1], new_nu[i] = fa[per_leaf * k + 2];
        if (want_updates) new_us[i] = fa[per_leaf * k + 3];
      }
      pp[k] = cp[k].data_ptr(), pg[k] = cg[k].data_ptr(), pmu[k] = cmu[k].data_ptr(), pnu[k] = cnu[k].data_ptr();
      op[k] = new_p[i].data_ptr(), omu[k] = new_mu[i].data_ptr(), onu[k] = new_nu[i].data_ptr();
      ous[k] = new_us[i] ? new_us[i]->data_ptr() : nullptr;
      n[k] = cg[k].numel(), cnt[k] = count[i];
    }
    if (inplace) {
      std::vector<void *> gio(G);
      for (size_t k = 0; k < G; ++k) gio[k] = const_cast<void *>(pg[k]);
      check(b200_adam_step_inplace_ex(kv.first.dtype, (int64_t)G, op.data(), gio.data(), omu.data(), onu.data(), n.data(),
                                      cnt.data(), b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled,
                                      maximize ? 1 : 0, keep_grads ? 0 : 1, c.stream),
            "fused_step (inplace)");
      if (!keep_grads)
        for (size_t k = 0; k < G; ++k) {  // a gradient that had to be re-laid-out: write the update back into it
          const Tensor &g = *grads[idx[k]];
          if (!cg[k].is_same(g)) const_cast<Tensor &>(g).copy_(cg[k]);
          new_us[idx[k]] = g;
        }
    } else {
      check(b200_adam_step_forward(kv.first.dtype, (int64_t)G, pp.data(), pg.data(), pmu.data(), pnu.data(), op.data(),
                                   omu.data(), onu.data(), ous.data(), n.data(), cnt.data(), b1, b2, eps, eps_root,
                                   step_size, wd_l2, wd_decoupled, maximize ? 1 : 0, c.stream),
            "fused_step");
    }
  }
  return {new_p, new_mu, new_nu, new_us};
}

// Backward of fused_step: incoming gradients of (new_params, scaled_updates, new_mu, new_nu), saved (new_mu, new_nu,
// grads [, params when wd_l2 != 0]).  Returns (dgrads, dmu, dnu, dparams).  Without weight decay the gradient w.r.t.
// params is the incoming dparams tensor itself (returned as is, nothing is written); with weight decay it is a new tensor.
std::tuple<std::vector<OptTensor>, std::vector<OptTensor>, std::vector<OptTensor>, std::vector<OptTensor>>
fused_step_backward(
    const std::vector<OptTensor> &dparams, const std::vector<OptTensor> &dus, const std::vector<OptTensor> &dmu_ext,
    const std::vector<OptTensor> &dnu_ext, const std::vector<Tensor> &new_mu, const std::vector<Tensor> &new_nu,
    const std::vector<Tensor> &g, const std::vector<bool> &need_dg, const std::vector<bool> &need_dmu,
    const std::vector<bool> &need_dnu, double b1, double b2, double eps, double eps_root, double step_size,
    const std::vector<size_t> &count, const std::vector<OptTensor> &params, const std::vector<bool> &need_dp,
    double wd_l2, double wd_decoupled, bool maximize) {
  const size_t L = g.size();
  if (dparams.size() != L || dus.size() != L || dmu_ext.size() != L || dnu_ext.size() != L || new_mu.size() != L ||
      new_nu.size() != L || need_dg.size() != L || need_dmu.size() != L || need_dnu.size() != L || count.size() != L ||
      params.size() != L || need_dp.size() != L)
    throw std::runtime_error("fused_step_backward: all per-leaf lists must have the same length");
  const bool decay = wd_l2 != 0.0 || wd_decoupled != 0.0;
  std::vector<OptTensor> dg(L), dmu(L), dnu(L), dp(L);
  std::vector<long> slot_dg(L, -1), slot_dmu(L, -1), slot_dnu(L, -1), slot_dp(L, -1);
  FlatAlloc fa;
  std::vector<Tensor> cdp(L), cdus(L), cme(L), cne(L), cm(L), cv(L), cg(L), cpar(L);  // see `conform`
  std::map<GroupKey, std::vector<size_t>> groups;
  for (size_t i = 0; i < L; ++i) {
    const bool hdu = dparams[i].has_value() || dus[i].has_value();
    const bool hme = dmu_ext[i].has_value(), hne = dnu_ext[i].has_value();
    if (!hdu && !hme && !hne) continue;
    const bool wg = need_dg[i], wm = need_dmu[i] && (hdu || hme), wn = need_dnu[i] && (hdu || hne);
    if (need_dp[i] && !decay) dp[i] = dparams[i];  // p' = p + us: pass-through, no bytes move
    const bool wp = need_dp[i] && decay;
    if (!wg && !wm && !wn && !wp) continue;
    require_cuda(g[i]);
    const int dt = family(g[i], new_mu[i], "adamStepBackwardCUDA", true);
    if (wd_l2 != 0.0 && !params[i].has_value())
      throw std::runtime_error("fused_step_backward: params are required for L2 weight decay");
    cm[i] = dense_donor(new_mu[i]), cv[i] = conform(new_nu[i], cm[i]), cg[i] = conform(g[i], cm[i]);
    if (params[i]) cpar[i] = conform(*params[i], cm[i]);
    if (wp) slot_dp[i] = (long)fa.request(cg[i]);
    if (dparams[i]) cdp[i] = conform(*dparams[i], cm[i]), same_type(cdp[i], g[i]);
    if (dus[i]) cdus[i] = conform(*dus[i], cm[i]), same_type(cdus[i], g[i]);
    if (hme) cme[i] = conform(*dmu_ext[i], cm[i]), same_type(cme[i], new_mu[i]);
    if (hne) cne[i] = conform(*dnu_ext[i], cm[i]), same_type(cne[i], new_mu[i]);
    if (wg) slot_dg[i] = (long)fa.request(cg[i]);
    if (wm) slot_dmu[i] = (long)fa.request(cm[i]);
    if (wn) slot_dnu[i] = (long)fa.request(cv[i]);
    groups[GroupKey{(int)g[i].device().index(), dt}].push_back(i);
  }
  fa.commit();
  for (size_t i = 0; i < L; ++i) {
    if (slot_dg[i] >= 0) dg[i] = fa[slot_dg[i]];
    if (slot_dmu[i] >= 0) dmu[i] = fa[slot_dmu[i]];
    if (slot_dnu[i] >= 0) dnu[i] = fa[slot_dnu[i]];
    if (slot_dp[i] >= 0) dp[i] = fa[slot_dp[i]];
  }
  for (auto &kv : groups) {
    const auto &idx = kv.second;
    const size_t G = idx.size();
    std::vector<const void *> pdp(G), pdus(G), pm(G), pv(G), pg(G), pme(G), pne(G), pp(G);
    std::vector<void *> odg(G), odm(G), odn(G), odp(G);
    std::vector<int64_t> n(G);
    std::vector<uint64_t> cnt(G);
    Ctx c(g[idx[0]]);
    for (size_t k = 0; k < G; ++k) {
      const size_t i = idx[k];
      pdp[k] = cdp[i].defined() ? cdp[i].data_ptr() : nullptr;
      pdus[k] = cdus[i].defined() ? cdus[i].data_ptr() : nullptr;
      pme[k] = cme[i].defined() ? cme[i].data_ptr() : nullptr;
      pne[k] = cne[i].defined() ? cne[i].data_ptr() : nullptr;
      pm[k] = cm[i].data_ptr(), pv[k] = cv[i].data_ptr(), pg[k] = cg[i].data_ptr();
      odg[k] = dg[i] ? dg[i]->data_ptr() : nullptr;
      odm[k] = dmu[i] ? dmu[i]->data_ptr() : nullptr;
      odn[k] = dnu[i] ? dnu[i]->data_ptr() : nullptr;
      pp[k] = cpar[i].defined() ? cpar[i].data_ptr() : nullptr;
      odp[k] = (decay && dp[i]) ? dp[i]->data_ptr() : nullptr;
      n[k] = cg[i].numel(), cnt[k] = count[i];
    }
    check(b200_adam_step_backward(kv.first.dtype, (int64_t)G, pdp.data(), pdus.data(), pm.data(), pv.data(), pg.data(),
                                  pme.data(), pne.data(), odg.data(), odm.data(), odn.data(), pp.data(), odp.data(),
                                  n.data(), cnt.data(), b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled,
                                  maximize ? 1 : 0, c.stream),
          "fused_step_backward");
  }
  return {dg, dmu, dnu, dp};
}

// ---- capturable in-place step (device-resident step counts + host-computed bias-correction tables) -------------
// (c1_table, c2_table) on the CPU for a dtype family: 0 = float32 tables (fp32 and bf16-param/fp32-state), 1 = float64
std::tuple<Tensor, Tensor> bias_tables(int dtype_family, double b1, double b2) {
  const int64_t len = b200_adam_bias_table_len(dtype_family, b1, b2);
  if (len <= 0)
    throw std::runtime_error("bias_tables: betas this close to 1 need more than 2^24 table entries; not capturable");
  auto opts = at::TensorOptions().dtype(dtype_family == B200_ADAM_F64 ? at::kDouble : at::kFloat).device(at::kCPU);
  Tensor c1 = at::empty({len}, opts), c2 = at::empty({len}, opts);
  check(b200_adam_bias_table(dtype_family, b1, b2, len, c1.data_ptr(), c2.data_ptr()), "bias_tables");
  return {c1, c2};
}

// In-place Adam step over all leaves whose step counts live on the DEVICE (0-dim int64 CUDA tensors, already
// incremented) -- safe to record in a CUDA graph.  Leaves must share one device and one dtype family.
void fused_step_capturable(const std::vector<Tensor> &params, const std::vector<Tensor> &grads,
                           const std::vector<Tensor> &mu, const std::vector<Tensor> &nu,
                           const std::vector<Tensor> &count, const Tensor &c1_table, const Tensor &c2_table, double b1,
                           double b2, double eps, double eps_root, double step_size, double wd_l2, double wd_decoupled,
                           bool maximize, bool keep_grads) {
  const size_t L = params.size();
  if (L == 0) return;
  if (grads.size() != L || mu.size() != L || nu.size() != L || count.size() != L)
    throw std::runtime_error("fused_step_capturable: params, grads, mu, nu and count must have the same number of leaves");
  require_cuda(params[0]);
  const int dt = family(grads[0], mu[0], "adamStepCUDA", true);
  const auto table_type = dt == B200_ADAM_F64 ? at::kDouble : at::kFloat;
  if (c1_table.scalar_type() != table_type || c2_table.scalar_type() != table_type || !c1_table.is_cuda() ||
      !c2_table.is_cuda() || !c1_table.is_contiguous() || !c2_table.is_contiguous() ||
      c1_table.numel() != c2_table.numel() || c1_table.numel() == 0)
    throw std::runtime_error("fused_step_capturable: bias tables must be contiguous CUDA tensors of the arithmetic type");
  std::vector<void *> pp(L), pg(L), pmu(L), pnu(L);
  std::vector<const int64_t *> pc(L);
  std::vector<int64_t> n(L);
  for (size_t i = 0; i < L; ++i) {
    require_cuda(params[i]);
    if (family(grads[i], mu[i], "adamStepCUDA", true) != dt || params[i].device() != params[0].device())
      throw std::runtime_error("fused_step_capturable: all leaves must share one device and one dtype family");
    same_type(mu[i], nu[i]), same_type(grads[i], params[i]);
    if (!count[i].is_cuda() || count[i].scalar_type() != at::kLong || count[i].numel() != 1)
      throw std::runtime_error("fused_step_capturable: step counts must be one-element int64 CUDA tensors");
    // a recorded step cannot hide a re-layout copy: every tensor of the leaf must already share the parameter's layout
    require_inplace_layout(grads[i], params[i], "the gradient");
    require_inplace_layout(mu[i], params[i], "mu"), require_inplace_layout(nu[i], params[i], "nu");
    pp[i] = params[i].data_ptr(), pg[i] = grads[i].data_ptr(), pmu[i] = mu[i].data_ptr(), pnu[i] = nu[i].data_ptr();
    pc[i] = count[i].data_ptr<int64_t>(), n[i] = grads[i].numel();
  }
  Ctx c(params[0]);
  check(b200_adam_step_inplace_capturable(dt, (int64_t)L, pp.data(), pg.data(), pmu.data(), pnu.data(), n.data(),
                                          pc.data(), c1_table.data_ptr(), c2_table.data_ptr(), c1_table.numel(), b1, b2,
                                          eps, eps_root, step_size, wd_l2, wd_decoupled, maximize ? 1 : 0,
                                          keep_grads ? 0 : 1, c.stream),
        "fused_step_capturable");
}

// ---- native autograd nodes ----------------------------------------------------------------------------
// The Python autograd.Function wrappers (accelerated_op/adam_op.py:FusedOp, adam_step.py:AdamStep.Op) spend ~200 us per
// call on a 62-leaf pytree in THPFunction bookkeeping (248 inputs / 186 outputs), 3x the launch itself.  These C++
// nodes do the same thing -- same kernels, same saved tensors, same None handling -- without the interpreter, and their
// backward runs on the autograd engine's thread without taking the GIL.
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

inline std::vector<size_t> to_counts(const std::vector<int64_t> &c) { return std::vector<size_t>(c.begin(), c.end()); }
inline OptTensor opt(const Tensor &t) { return t.defined() ? OptTensor(t) : OptTensor(); }
inline Tensor undef(const OptTensor &t) { return t.has_value() ? *t : Tensor(); }

// tensors = g_0.., mu_0.., nu_0.. ; hyper = b1, b2, eps, eps_root ; outputs mu'_0.., nu'_0.., u_0..
struct FusedNode : public torch::autograd::Function<FusedNode> {
  static variable_list forward(AutogradContext *ctx, at::TensorList tensors, std::vector<int64_t> counts,
                               std::vector<double> hyper) {
    const size_t L = counts.size();
    std::vector<OptTensor> g(L);
    std::vector<Tensor> mu(L), nu(L);
    for (size_t i = 0; i < L; ++i) g[i] = tensors[i], mu[i] = tensors[L + i], nu[i] = tensors[2 * L + i];
    auto out = fused_forward(g, mu, nu, hyper[0], hyper[1], hyper[2], hyper[3], to_counts(counts), false);
    auto &new_mu = std::get<0>(out);
    auto &new_nu = std::get<1>(out);
    auto &new_u = std::get<2>(out);
    ctx->set_materialize_grads(false);  // absent gradients stay undefined -> the kernel skips those streams
    ctx->saved_data["counts"] = counts;
    ctx->saved_data["hyper"] = hyper;
    variable_list saved, result;
    saved.reserve(3 * L), result.reserve(3 * L);
    for (size_t i = 0; i < L; ++i) saved.push_back(tensors[i]);   // g
    for (size_t i = 0; i < L; ++i) saved.push_back(*new_u[i]);    // u
    for (size_t i = 0; i < L; ++i) saved.push_back(new_mu[i]);    // mu'
    ctx->save_for_backward(saved);
    for (size_t i = 0; i < L; ++i) result.push_back(new_mu[i]);
    for (size_t i = 0; i < L; ++i) result.push_back(new_nu[i]);
    for (size_t i = 0; i < L; ++i) result.push_back(*new_u[i]);
    return result;
  }

  static variable_list backward(AutogradContext *ctx, variable_list douts) {
    const auto counts = ctx->saved_data["counts"].toIntVector();
    const auto hyper = ctx->saved_data["hyper"].toDoubleVector();
    const size_t L = counts.size();
    const variable_list saved = ctx->get_saved_variables();
    std::vector<Tensor> g(saved.begin(), saved.begin() + L), u(saved.begin() + L, saved.begin() + 2 * L),
        new_mu(saved.begin() + 2 * L, saved.end());
    std::vector<OptTensor> dme(L), dne(L), du(L);
    std::vector<bool> need_g(L), need_mu(L), need_nu(L);
    for (size_t i = 0; i < L; ++i) {
      dme[i] = opt(douts[i]), dne[i] = opt(douts[L + i]), du[i] = opt(douts[2 * L + i]);
      need_g[i] = ctx->needs_input_grad(i), need_mu[i] = ctx->needs_input_grad(L + i);
      need_nu[i] = ctx->needs_input_grad(2 * L + i);
    }
    auto out = fused_backward(du, dme, dne, u, new_mu, g, need_g, need_mu, need_nu, hyper[0], hyper[1], hyper[3],
                              to_counts(counts));
    variable_list grads;
    grads.reserve(3 * L + 2);
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<0>(out)[i]));
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<1>(out)[i]));
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<2>(out)[i]));
    grads.emplace_back(), grads.emplace_back();  // counts, hyper
    return grads;
  }
};

// tensors = p_0.., g_0.., mu_0.., nu_0.. ; hyper = b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize ;
// outputs p'_0.., mu'_0.., nu'_0..
struct StepNode : public torch::autograd::Function<StepNode> {
  static variable_list forward(AutogradContext *ctx, at::TensorList tensors, std::vector<int64_t> counts,
                               std::vector<double> hyper) {
    const size_t L = counts.size();
    std::vector<Tensor> p(L), mu(L), nu(L);
    std::vector<OptTensor> g(L);
    for (size_t i = 0; i < L; ++i)
      p[i] = tensors[i], g[i] = tensors[L + i], mu[i] = tensors[2 * L + i], nu[i] = tensors[3 * L + i];
    auto out = fused_step(p, g, mu, nu, hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], to_counts(counts), false, false,
                          hyper[5], hyper[6], hyper[7] != 0.0);
    auto &new_p = std::get<0>(out);
    auto &new_mu = std::get<1>(out);
    auto &new_nu = std::get<2>(out);
    ctx->set_materialize_grads(false);
    ctx->saved_data["counts"] = counts;
    ctx->saved_data["hyper"] = hyper;
    // u is recomputed in the backward kernel; the raw parameters are needed only to re-derive g + wd*p
    variable_list saved, result;
    for (size_t i = 0; i < L; ++i) saved.push_back(tensors[L + i]);
    for (size_t i = 0; i < L; ++i) saved.push_back(new_mu[i]);
    for (size_t i = 0; i < L; ++i) saved.push_back(new_nu[i]);
    if (hyper[5] != 0.0)
      for (size_t i = 0; i < L; ++i) saved.push_back(tensors[i]);
    ctx->save_for_backward(saved);
    result.reserve(3 * L);
    for (size_t i = 0; i < L; ++i) result.push_back(new_p[i]);
    for (size_t i = 0; i < L; ++i) result.push_back(new_mu[i]);
    for (size_t i = 0; i < L; ++i) result.push_back(new_nu[i]);
    return result;
  }

  static variable_list backward(AutogradContext *ctx, variable_list douts) {
    const auto counts = ctx->saved_data["counts"].toIntVector();
    const auto hyper = ctx->saved_data["hyper"].toDoubleVector();
    const size_t L = counts.size();
    const bool l2 = hyper[5] != 0.0;
    const variable_list saved = ctx->get_saved_variables();
    std::vector<Tensor> g(saved.begin(), saved.begin() + L), new_mu(saved.begin() + L, saved.begin() + 2 * L),
        new_nu(saved.begin() + 2 * L, saved.begin() + 3 * L);
    std::vector<OptTensor> params(L), dp_in(L), dus(L), dme(L), dne(L);
    std::vector<bool> need_p(L), need_g(L), need_mu(L), need_nu(L);
    for (size_t i = 0; i < L; ++i) {
      if (l2) params[i] = saved[3 * L + i];
      dp_in[i] = opt(douts[i]), dme[i] = opt(douts[L + i]), dne[i] = opt(douts[2 * L + i]);
      need_p[i] = ctx->needs_input_grad(i), need_g[i] = ctx->needs_input_grad(L + i);
      need_mu[i] = ctx->needs_input_grad(2 * L + i), need_nu[i] = ctx->needs_input_grad(3 * L + i);
    }
    auto out = fused_step_backward(dp_in, dus, dme, dne, new_mu, new_nu, g, need_g, need_mu, need_nu, hyper[0], hyper[1],
                                   hyper[2], hyper[3], hyper[4], to_counts(counts), params, need_p, hyper[5], hyper[6],
                                   hyper[7] != 0.0);
    variable_list grads;
    grads.reserve(4 * L + 2);
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<3>(out)[i]));  // dparams
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<0>(out)[i]));  // dgrads
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<1>(out)[i]));  // dmu
    for (size_t i = 0; i < L; ++i) grads.push_back(undef(std::get<2>(out)[i]));  // dnu
    grads.emplace_back(), grads.emplace_back();  // counts, hyper
    return grads;
  }
};

// ---- flat-state whole step (the default path of MetaAdam / FuncAdam / Adam) -------------------------------------------
// The per-leaf nodes above hand autograd 4L inputs and 3L outputs per step; most of the host time of a 62-leaf step
// is that bookkeeping.  Here the moments of ALL leaves of one (device, dtype) group live in ONE buffer each
// (mu_flat, nu_flat; live leaf k occupies [offs[k], offs[k] + numel_k), 128-byte aligned), so autograd sees
// 2K + 2 inputs and K + 2 outputs; the parameters and gradients stay per-leaf tensors (the caller's model needs them that
// way) and reach the kernel through the leaf table like before.  Leaves WITHOUT a gradient this step are not passed at
// all: their parameters are not touched and their moment ranges (`keep` = begin/end pairs) are copied through, which is
// the reference's skip semantics (accelerated_op/adam_op.py:148-149, optim/meta/base.py:58-96).
// Padding between leaves is never read by any kernel and holds unspecified values.
inline int flat_family(const Tensor &p, const Tensor &mu_flat) { return family(p, mu_flat, "adamStepCUDA", true); }

struct FlatPtrs {
  std::vector<const void *> p, g, mu, nu, dp, dme, dne;
  std::vector<void *> po, muo, nuo, dg, dmu, dnu, dpo;
  std::vector<int64_t> n;
  std::vector<uint64_t> cnt;
  explicit FlatPtrs(size_t K)
      : p(K), g(K), mu(K), nu(K), dp(K), dme(K), dne(K), po(K), muo(K), nuo(K), dg(K), dmu(K), dnu(K), dpo(K), n(K),
        cnt(K) {}
};

inline char *at_offset(const Tensor &flat, int64_t off) {
  return flat.defined() ? (char *)flat.data_ptr() + off * (int64_t)flat.element_size() : nullptr;
}

inline void copy_ranges(Tensor &dst, const Tensor &src, const std::vector<int64_t> &keep) {
  for (size_t r = 0; r + 1 < keep.size(); r += 2)
    if (keep[r + 1] > keep[r]) dst.narrow(0, keep[r], keep[r + 1] - keep[r]).copy_(src.narrow(0, keep[r], keep[r + 1] - keep[r]));
}

// forward without autograd: returns p'_0.., mu'_flat, nu'_flat ; `cp`, `cg` receive the row-major tensors actually read
// Capturable form: `tensors` carries five more entries -- the per-leaf device step counts of THIS step (an int64 vector,
// entry cidx[k] belongs to live leaf k; already incremented, snapshotted by the caller) and the four bias tables
// c1, c2, o1, c2e -- and `cidx` is non-empty; the host `counts` are then ignored.
struct CapTensors {
  const Tensor *cnt = nullptr, *c1 = nullptr, *c2 = nullptr, *o1 = nullptr, *c2e = nullptr;
  bool on() const { return cnt != nullptr; }
  void check(int dt, const Tensor &mu_flat) const {
    const auto table_type = dt == B200_ADAM_F64 ? at::kDouble : at::kFloat;
    for (const Tensor *t : {c1, c2, o1, c2e})
      if (!t->is_cuda() || t->scalar_type() != table_type || !t->is_contiguous() || t->numel() != c1->numel() ||
          t->numel() == 0 || t->device() != mu_flat.device())
        throw std::runtime_error("flat_step: bias tables must be contiguous CUDA tensors of the arithmetic type");
    if (!cnt->is_cuda() || cnt->scalar_type() != at::kLong || !cnt->is_contiguous() || cnt->device() != mu_flat.device())
      throw std::runtime_error("flat_step: the device step counts must be a contiguous int64 CUDA vector");
  }
};

inline CapTensors cap_tensors(at::TensorList tensors, size_t K, const std::vector<int64_t> &cidx) {
  CapTensors c;
  if (cidx.empty()) return c;
  if (tensors.size() != 2 * K + 7 || cidx.size() != K)
    throw std::runtime_error("flat_step: the capturable form expects count vector + 4 bias tables after the flat moments");
  c.cnt = &tensors[2 * K + 2], c.c1 = &tensors[2 * K + 3], c.c2 = &tensors[2 * K + 4];
  c.o1 = &tensors[2 * K + 5], c.c2e = &tensors[2 * K + 6];
  for (int64_t i : cidx)
    if (i < 0 || i >= c.cnt->numel()) throw std::runtime_error("flat_step: step-count index out of range");
  return c;
}

variable_list flat_step_forward(at::TensorList tensors, const std::vector<int64_t> &offs, const std::vector<int64_t> &keep,
                                const std::vector<int64_t> &counts, const std::vector<double> &hyper,
                                std::vector<Tensor> &cp, std::vector<Tensor> &cg, const std::vector<int64_t> &cidx) {
  const size_t K = counts.size();
  const CapTensors cap = cap_tensors(tensors, K, cidx);
  if ((!cap.on() && tensors.size() != 2 * K + 2) || offs.size() != K || hyper.size() != 8)
    throw std::runtime_error("flat_step: expected p_0.., g_0.., mu_flat, nu_flat with one offset and count per leaf");
  const Tensor &mu_flat = tensors[2 * K], &nu_flat = tensors[2 * K + 1];
  if (mu_flat.dim() != 1 || !mu_flat.is_contiguous() || !nu_flat.is_contiguous() || mu_flat.sizes() != nu_flat.sizes())
    throw std::runtime_error("flat_step: the flat moment buffers must be contiguous 1-D tensors of equal length");
  require_cuda(mu_flat);
  same_type(mu_flat, nu_flat);
  Tensor mu_out = at::empty_like(mu_flat), nu_out = at::empty_like(nu_flat);
  copy_ranges(mu_out, mu_flat, keep), copy_ranges(nu_out, nu_flat, keep);
  variable_list result;
  result.reserve(K + 2);
  if (K > 0) {
    const int dt = flat_family(tensors[0], mu_flat);
    Ctx c(mu_flat);
    FlatAlloc fa;
    cp.resize(K), cg.resize(K);
    for (size_t k = 0; k < K; ++k) {
      const Tensor &p = tensors[k], &g = tensors[K + k];
      if (flat_family(p, mu_flat) != dt || p.device() != mu_flat.device())
        throw std::runtime_error("flat_step: all leaves of a group must share one device and one dtype family");
      same_type(p, g);
      cp[k] = p.is_contiguous() ? p : p.contiguous();  // flat moments are row-major: so is every array of the leaf
      cg[k] = conform(g, cp[k]);
      if (offs[k] < 0 || offs[k] + cp[k].numel() > mu_flat.numel())
        throw std::runtime_error("flat_step: a leaf does not fit the flat moment buffers");
      fa.request(cp[k]);
    }
    fa.commit();
    FlatPtrs q(K);
    for (size_t k = 0; k < K; ++k) {
      result.push_back(fa[k]);
      q.p[k] = cp[k].data_ptr(), q.g[k] = cg[k].data_ptr();
      q.mu[k] = at_offset(mu_flat, offs[k]), q.nu[k] = at_offset(nu_flat, offs[k]);
      q.po[k] = fa[k].data_ptr(), q.muo[k] = at_offset(mu_out, offs[k]), q.nuo[k] = at_offset(nu_out, offs[k]);
      q.n[k] = cp[k].numel(), q.cnt[k] = (uint64_t)counts[k];
    }
    if (cap.on()) {
      cap.check(dt, mu_flat);
      std::vector<const int64_t *> pc(K);
      for (size_t k = 0; k < K; ++k) pc[k] = cap.cnt->data_ptr<int64_t>() + cidx[k];
      check(b200_adam_step_forward_capturable(dt, (int64_t)K, q.p.data(), q.g.data(), q.mu.data(), q.nu.data(),
                                              q.po.data(), q.muo.data(), q.nuo.data(), nullptr, q.n.data(), pc.data(),
                                              cap.c1->data_ptr(), cap.c2->data_ptr(), cap.c1->numel(), hyper[0], hyper[1],
                                              hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7] != 0.0 ? 1 : 0,
                                              c.stream),
            "flat_step (capturable)");
    } else {
      check(b200_adam_step_forward(dt, (int64_t)K, q.p.data(), q.g.data(), q.mu.data(), q.nu.data(), q.po.data(),
                                   q.muo.data(), q.nuo.data(), nullptr, q.n.data(), q.cnt.data(), hyper[0], hyper[1],
                                   hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7] != 0.0 ? 1 : 0, c.stream),
            "flat_step");
    }
  }
  result.push_back(mu_out), result.push_back(nu_out);
  return result;
}

struct FlatStepNode : public torch::autograd::Function<FlatStepNode> {
  static variable_list forward(AutogradContext *ctx, at::TensorList tensors, std::vector<int64_t> offs,
                               std::vector<int64_t> keep, std::vector<int64_t> counts, std::vector<double> hyper,
                               std::vector<int64_t> cidx) {
    const size_t K = counts.size();
    std::vector<Tensor> cp, cg;
    variable_list result = flat_step_forward(tensors, offs, keep, counts, hyper, cp, cg, cidx);
    ctx->set_materialize_grads(false);
    ctx->saved_data["offs"] = offs, ctx->saved_data["keep"] = keep;
    ctx->saved_data["counts"] = counts, ctx->saved_data["hyper"] = hyper, ctx->saved_data["cidx"] = cidx;
    variable_list saved;
    saved.reserve(2 * K + 2);
    saved.push_back(result[K]), saved.push_back(result[K + 1]);  // mu', nu' (u is recomputed in the backward kernel)
    // the tensor the kernel read is saved: the caller's own gradient when it was already row-major (then it is an INPUT,
    // saved without a reference cycle), a private copy otherwise
    for (size_t k = 0; k < K; ++k) saved.push_back(cg[k]);
    if (hyper[5] != 0.0)
      for (size_t k = 0; k < K; ++k) saved.push_back(cp[k]);  // L2 decay: g2 = g + wd * p is re-derived
    if (!cidx.empty())
      for (size_t a = 0; a < 5; ++a) saved.push_back(tensors[2 * K + 2 + a]);  // count snapshot + the four tables
    ctx->save_for_backward(saved);
    return result;
  }

  static variable_list backward(AutogradContext *ctx, variable_list douts) {
    const auto offs = ctx->saved_data["offs"].toIntVector();
    const auto keep = ctx->saved_data["keep"].toIntVector();
    const auto counts = ctx->saved_data["counts"].toIntVector();
    const auto hyper = ctx->saved_data["hyper"].toDoubleVector();
    const auto cidx = ctx->saved_data["cidx"].toIntVector();
    const size_t K = counts.size();
    const bool l2 = hyper[5] != 0.0, decay = l2 || hyper[6] != 0.0, capt = !cidx.empty();
    const variable_list saved = ctx->get_saved_variables();
    const Tensor &new_mu = saved[0], &new_nu = saved[1];
    const size_t n_in = 2 * K + 2 + (capt ? 5 : 0);   // tensor inputs of the forward
    Tensor dme = douts[K].defined() ? douts[K].contiguous() : Tensor();
    Tensor dne = douts[K + 1].defined() ? douts[K + 1].contiguous() : Tensor();
    const bool need_mu = ctx->needs_input_grad(2 * K), need_nu = ctx->needs_input_grad(2 * K + 1);
    variable_list grads(n_in + 5);  // dp_0.., dg_0.., dmu_flat, dnu_flat [, 5 capturable tensors], offs, keep, counts, hyper, cidx
    // a leaf that receives no gradient at all leaves its range of dmu / dnu unwritten by the kernel: start from zeros
    // whenever that can happen (some dp' missing, or skipped leaves), from uninitialised memory otherwise
    auto fully_written = [&](const Tensor &dext) {  // will the kernel write the range of EVERY leaf?
      if (!keep.empty()) return false;
      if (dext.defined()) return true;
      for (size_t k = 0; k < K; ++k)
        if (!douts[k].defined()) return false;
      return true;
    };
    Tensor dmu_flat, dnu_flat;
    if (need_mu) dmu_flat = fully_written(dme) ? at::empty_like(new_mu) : at::zeros_like(new_mu);
    if (need_nu) dnu_flat = fully_written(dne) ? at::empty_like(new_nu) : at::zeros_like(new_nu);
    if (dmu_flat.defined() && dme.defined()) copy_ranges(dmu_flat, dme, keep);
    if (dnu_flat.defined() && dne.defined()) copy_ranges(dnu_flat, dne, keep);
    if (K == 0) {
      grads[0] = dmu_flat, grads[1] = dnu_flat;
      return grads;
    }
    const int dt = flat_family(saved[2], new_mu);
    Ctx c(new_mu);
    FlatAlloc fa;
    std::vector<Tensor> cdp(K);
    std::vector<long> slot_dg(K, -1), slot_dp(K, -1);
    std::vector<size_t> live;
    live.reserve(K);
    for (size_t k = 0; k < K; ++k) {
      const Tensor &g = saved[2 + k];
      const bool hdp = douts[k].defined();
      const bool need_p = ctx->needs_input_grad(k), need_g = ctx->needs_input_grad(K + k);
      if (need_p && !decay) grads[k] = douts[k];  // p' = p + us: pass-through, no bytes move
      if (!hdp && !dme.defined() && !dne.defined()) continue;
      const bool wp = need_p && decay;
      if (!need_g && !wp && !dmu_flat.defined() && !dnu_flat.defined()) continue;
      if (hdp) cdp[k] = conform(douts[k], g), same_type(cdp[k], g);
      if (need_g) slot_dg[k] = (long)fa.request(g);
      if (wp) slot_dp[k] = (long)fa.request(g);
      live.push_back(k);
    }
    fa.commit();
    const size_t G = live.size();
    FlatPtrs q(G);
    for (size_t j = 0; j < G; ++j) {
      const size_t k = live[j];
      const Tensor &g = saved[2 + k];
      if (slot_dg[k] >= 0) grads[K + k] = fa[slot_dg[k]];
      if (slot_dp[k] >= 0) grads[k] = fa[slot_dp[k]];
      q.dp[j] = cdp[k].defined() ? cdp[k].data_ptr() : nullptr;
      q.mu[j] = at_offset(new_mu, offs[k]), q.nu[j] = at_offset(new_nu, offs[k]), q.g[j] = g.data_ptr();
      q.dme[j] = at_offset(dme, offs[k]), q.dne[j] = at_offset(dne, offs[k]);
      q.dg[j] = slot_dg[k] >= 0 ? grads[K + k].data_ptr() : nullptr;
      // a moment gradient exists for this leaf iff something flows into it (dp' or the incoming moment gradient)
      q.dmu[j] = (cdp[k].defined() || dme.defined()) ? (void *)at_offset(dmu_flat, offs[k]) : nullptr;
      q.dnu[j] = (cdp[k].defined() || dne.defined()) ? (void *)at_offset(dnu_flat, offs[k]) : nullptr;
      q.p[j] = l2 ? saved[2 + K + k].data_ptr() : nullptr;
      q.dpo[j] = slot_dp[k] >= 0 ? grads[k].data_ptr() : nullptr;
      q.n[j] = g.numel(), q.cnt[j] = (uint64_t)counts[k];
    }
    if (G > 0 && capt) {
      const size_t base = 2 + K + (l2 ? K : 0);
      CapTensors cap;
      cap.cnt = &saved[base], cap.c1 = &saved[base + 1], cap.c2 = &saved[base + 2], cap.o1 = &saved[base + 3];
      cap.c2e = &saved[base + 4];
      std::vector<const int64_t *> pc(G);
      for (size_t j = 0; j < G; ++j) pc[j] = cap.cnt->data_ptr<int64_t>() + cidx[live[j]];
      check(b200_adam_step_backward_capturable(dt, (int64_t)G, q.dp.data(), nullptr, q.mu.data(), q.nu.data(), q.g.data(),
                                               q.dme.data(), q.dne.data(), q.dg.data(), q.dmu.data(), q.dnu.data(),
                                               q.p.data(), q.dpo.data(), q.n.data(), pc.data(), cap.c1->data_ptr(),
                                               cap.c2->data_ptr(), cap.o1->data_ptr(), cap.c2e->data_ptr(),
                                               cap.c1->numel(), hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5],
                                               hyper[6], hyper[7] != 0.0 ? 1 : 0, c.stream),
            "flat_step backward (capturable)");
    } else if (G > 0) {
      check(b200_adam_step_backward(dt, (int64_t)G, q.dp.data(), nullptr, q.mu.data(), q.nu.data(), q.g.data(),
                                    q.dme.data(), q.dne.data(), q.dg.data(), q.dmu.data(), q.dnu.data(), q.p.data(),
                                    q.dpo.data(), q.n.data(), q.cnt.data(), hyper[0], hyper[1], hyper[2], hyper[3], hyper[4],
                                    hyper[5], hyper[6], hyper[7] != 0.0 ? 1 : 0, c.stream),
            "flat_step backward");
    }
    grads[2 * K] = dmu_flat, grads[2 * K + 1] = dnu_flat;
    return grads;
  }
};

// Python entry point of the flat-state step.  `params` / `grads`: one entry per leaf of the group, `grads[i]` may be
// None (leaf skipped this step: parameter returned as is, moments and count untouched).  `offsets[i]`: element offset
// of leaf i in the flat moment buffers.  `counts[i]`: the step count AFTER this step for leaves with a gradient.
// Returns (new_params, mu_flat', nu_flat').  `inplace`: parameters and flat moments are updated in place, no autograd
// (Optimizer.step); gradients are overwritten with the scaled update unless keep_grads (the reference's side effect).
std::tuple<std::vector<Tensor>, Tensor, Tensor> flat_step(const std::vector<Tensor> &params,
                                                          const std::vector<OptTensor> &grads, const Tensor &mu_flat,
                                                          const Tensor &nu_flat, const std::vector<int64_t> &offsets,
                                                          const std::vector<int64_t> &counts, double b1, double b2,
                                                          double eps, double eps_root, double step_size, double wd_l2,
                                                          double wd_decoupled, bool maximize, bool inplace,
                                                          bool keep_grads, const OptTensor &count_dev,
                                                          const std::vector<Tensor> &tables) {
  const size_t L = params.size();
  if (grads.size() != L || offsets.size() != L || counts.size() != L)
    throw std::runtime_error("flat_step: params, grads, offsets and counts must have the same number of leaves");
  const bool capt = count_dev.has_value();
  if (capt && (tables.size() != 4 || inplace || count_dev->numel() != (int64_t)L))
    throw std::runtime_error("flat_step: the capturable form takes one device count per leaf and the 4 bias tables "
                             "(c1, c2, o1, c2e) and is out of place");
  std::vector<size_t> live;
  std::vector<int64_t> keep;
  for (size_t i = 0; i < L; ++i) {
    if (grads[i].has_value()) {
      live.push_back(i);
    } else if (params[i].numel() > 0) {
      keep.push_back(offsets[i]), keep.push_back(offsets[i] + params[i].numel());
    }
  }
  const size_t K = live.size();
  std::vector<Tensor> new_p(params);
  const std::vector<double> hyper{b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize ? 1.0 : 0.0};
  if (K == 0) return {new_p, mu_flat, nu_flat};
  if (inplace) {
    require_cuda(mu_flat);
    const int dt = flat_family(params[live[0]], mu_flat);
    Ctx c(mu_flat);
    FlatPtrs q(K);
    std::vector<Tensor> cg(K);
    for (size_t k = 0; k < K; ++k) {
      const size_t i = live[k];
      const Tensor &p = params[i];
      if (!p.is_contiguous() || flat_family(p, mu_flat) != dt || p.device() != mu_flat.device())
        throw std::runtime_error("torchopt_b200: in-place step: parameters must be row-major tensors of one device and dtype");
      cg[k] = conform(*grads[i], p);
      q.po[k] = p.data_ptr(), q.dg[k] = cg[k].data_ptr();
      q.muo[k] = at_offset(mu_flat, offsets[i]), q.nuo[k] = at_offset(nu_flat, offsets[i]);
      q.n[k] = p.numel(), q.cnt[k] = (uint64_t)counts[i];
    }
    check(b200_adam_step_inplace_ex(dt, (int64_t)K, q.po.data(), q.dg.data(), q.muo.data(), q.nuo.data(), q.n.data(),
                                    q.cnt.data(), b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize ? 1 : 0,
                                    keep_grads ? 0 : 1, c.stream),
          "flat_step (inplace)");
    if (!keep_grads)
      for (size_t k = 0; k < K; ++k) {
        const Tensor &g = *grads[live[k]];
        if (!cg[k].is_same(g)) const_cast<Tensor &>(g).copy_(cg[k]);
      }
    return {new_p, mu_flat, nu_flat};
  }
  std::vector<Tensor> flat;
  std::vector<int64_t> offs(K), cnt(K), cidx;
  flat.reserve(2 * K + 7);
  for (size_t k = 0; k < K; ++k) flat.push_back(params[live[k]]), offs[k] = offsets[live[k]], cnt[k] = counts[live[k]];
  for (size_t k = 0; k < K; ++k) flat.push_back(*grads[live[k]]);
  flat.push_back(mu_flat), flat.push_back(nu_flat);
  if (capt) {
    flat.push_back(*count_dev);
    for (const Tensor &t : tables) flat.push_back(t);
    cidx.assign(live.begin(), live.end());
  }
  variable_list out = FlatStepNode::apply(at::TensorList(flat), offs, keep, cnt, hyper, cidx);
  for (size_t k = 0; k < K; ++k) new_p[live[k]] = out[k];
  return {new_p, out[K], out[K + 1]};
}

// Python entry points: all lists hold one tensor per leaf (no None: the callers filter leaves without a gradient)
std::vector<Tensor> fused_autograd(const std::vector<Tensor> &updates, const std::vector<Tensor> &mu,
                                   const std::vector<Tensor> &nu, double b1, double b2, double eps, double eps_root,
                                   const std::vector<int64_t> &count) {
  const size_t L = updates.size();
  if (mu.size() != L || nu.size() != L || count.size() != L)
    throw std::runtime_error("fused_autograd: updates, mu, nu and count must have the same number of leaves");
  std::vector<Tensor> flat;
  flat.reserve(3 * L);
  flat.insert(flat.end(), updates.begin(), updates.end());
  flat.insert(flat.end(), mu.begin(), mu.end());
  flat.insert(flat.end(), nu.begin(), nu.end());
  return FusedNode::apply(at::TensorList(flat), count, std::vector<double>{b1, b2, eps, eps_root});
}

std::vector<Tensor> step_autograd(const std::vector<Tensor> &params, const std::vector<Tensor> &grads,
                                  const std::vector<Tensor> &mu, const std::vector<Tensor> &nu, double b1, double b2,
                                  double eps, double eps_root, double step_size, const std::vector<int64_t> &count,
                                  double wd_l2, double wd_decoupled, bool maximize) {
  const size_t L = params.size();
  if (grads.size() != L || mu.size() != L || nu.size() != L || count.size() != L)
    throw std::runtime_error("step_autograd: params, grads, mu, nu and count must have the same number of leaves");
  std::vector<Tensor> flat;
  flat.reserve(4 * L);
  flat.insert(flat.end(), params.begin(), params.end());
  flat.insert(flat.end(), grads.begin(), grads.end());
  flat.insert(flat.end(), mu.begin(), mu.end());
  flat.insert(flat.end(), nu.begin(), nu.end());
  return StepNode::apply(at::TensorList(flat), count,
                         std::vector<double>{b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled, maximize ? 1.0 : 0.0});
}

// host-buffer step (CPU tensors in, CPU tensors out, work done on the current CUDA device)
std::array<Tensor, 6> host_step(const Tensor &g, const Tensor &mu, const Tensor &nu, const Tensor &du,
                                const Tensor &dmu_ext, const Tensor &dnu_ext, std::array<Tensor, 6> out, double b1,
                                double b2, double eps, double eps_root, size_t count, int64_t slice_elems,
                                bool allow_pageable) {
  const Tensor *ins[6] = {&g, &mu, &nu, &du, &dmu_ext, &dnu_ext};
  for (int a = 0; a < 6; ++a) {
    // cudaMemcpyAsync from / to pageable memory is staged through the driver's own bounce buffer and serialises the
    // three streams of the pipeline: refuse it unless the caller says that is acceptable
    if (!allow_pageable && (!ins[a]->is_pinned() || !out[a].is_pinned()))
      throw std::runtime_error("host_step: expects page-locked (pin_memory()) CPU tensors; pass allow_pageable=True to "
                               "accept pageable ones at a fraction of the link rate");
    if (!ins[a]->device().is_cpu() || !out[a].device().is_cpu() || ins[a]->scalar_type() != at::kFloat ||
        out[a].scalar_type() != at::kFloat || !ins[a]->is_contiguous() || !out[a].is_contiguous() ||
        ins[a]->numel() != g.numel() || out[a].numel() != g.numel())
      throw std::runtime_error("host_step: expects contiguous float32 CPU tensors of equal numel");
  }
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = b200_adam_host_step_f32(g.data_ptr<float>(), mu.data_ptr<float>(), nu.data_ptr<float>(), du.data_ptr<float>(),
                                 dmu_ext.data_ptr<float>(), dnu_ext.data_ptr<float>(), out[0].data_ptr<float>(),
                                 out[1].data_ptr<float>(), out[2].data_ptr<float>(), out[3].data_ptr<float>(),
                                 out[4].data_ptr<float>(), out[5].data_ptr<float>(), g.numel(), b1, b2, eps, eps_root,
                                 count, slice_elems);
  }
  check(rc, "host_step");
  return out;
}

}  // namespace

PYBIND11_MODULE(_C, mod) {
  mod.doc() = "torchopt_b200 native module: B200 (sm_100a) accelerated differentiable Adam";
  py::module m = mod.def_submodule("adam_op", "Adam Ops");
  // -- the reference surface: names, keyword names and order as in adam_op.cpp:157-214
  m.def("forward_", &forward_inplace, "Adam forward inplace", py::arg("updates"), py::arg("mu"), py::arg("nu"),
        py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("count"));
  m.def("forward_mu", &forward_mu, "Adam forward mu", py::arg("updates"), py::arg("mu"), py::arg("b1"));
  m.def("forward_nu", &forward_nu, "Adam forward nu", py::arg("updates"), py::arg("nu"), py::arg("b2"));
  m.def("forward_updates", &forward_updates, "Adam forward updates", py::arg("new_mu"), py::arg("new_nu"),
        py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("count"));
  m.def("backward_mu", &backward_mu, "Adam backward mu", py::arg("dmu"), py::arg("updates"), py::arg("mu"),
        py::arg("b1"));
  m.def("backward_nu", &backward_nu, "Adam backward nu", py::arg("dnu"), py::arg("updates"), py::arg("nu"),
        py::arg("b1"));
  m.def("backward_updates", &backward_updates, "Adam backward updates", py::arg("dupdates"), py::arg("updates"),
        py::arg("new_mu"), py::arg("new_nu"), py::arg("b1"), py::arg("b2"), py::arg("eps_root"), py::arg("count"));
  // -- new: fused, multi-tensor
  m.def("fused_forward", &fused_forward, "Fused differentiable Adam forward over all leaves (one launch)",
        py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("b1"), py::arg("b2"), py::arg("eps"),
        py::arg("eps_root"), py::arg("count"), py::arg("inplace") = false);
  m.def("fused_backward", &fused_backward, "Fused differentiable Adam backward over all leaves (one launch)",
        py::arg("dupdates"), py::arg("dmu_ext"), py::arg("dnu_ext"), py::arg("new_updates"), py::arg("new_mu"),
        py::arg("updates"), py::arg("need_dupdates"), py::arg("need_dmu"), py::arg("need_dnu"), py::arg("b1"),
        py::arg("b2"), py::arg("eps_root"), py::arg("count"));
  m.def("fused_step", &fused_step, "Adam update + scale(step_size) + apply_updates over all leaves (one launch)",
        py::arg("params"), py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("b1"), py::arg("b2"),
        py::arg("eps"), py::arg("eps_root"), py::arg("step_size"), py::arg("count"), py::arg("inplace") = false,
        py::arg("want_updates") = false, py::arg("wd_l2") = 0.0, py::arg("wd_decoupled") = 0.0,
        py::arg("maximize") = false, py::arg("keep_grads") = false);
  m.def("fused_step_backward", &fused_step_backward, "Backward of fused_step over all leaves (one launch)",
        py::arg("dparams"), py::arg("dscaled_updates"), py::arg("dmu_ext"), py::arg("dnu_ext"), py::arg("new_mu"),
        py::arg("new_nu"), py::arg("updates"), py::arg("need_dupdates"), py::arg("need_dmu"), py::arg("need_dnu"),
        py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("step_size"), py::arg("count"),
        py::arg("params"), py::arg("need_dparams"), py::arg("wd_l2") = 0.0, py::arg("wd_decoupled") = 0.0,
        py::arg("maximize") = false);
  m.def("bias_tables", &bias_tables, "host-computed bias-correction tables (c1, c2) for the capturable step",
        py::arg("dtype_family"), py::arg("b1"), py::arg("b2"));
  m.def("fused_step_capturable", &fused_step_capturable,
        "in-place Adam step with device-resident step counts (recordable in a CUDA graph)", py::arg("params"),
        py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("count"), py::arg("c1_table"), py::arg("c2_table"),
        py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("step_size"),
        py::arg("wd_l2") = 0.0, py::arg("wd_decoupled") = 0.0, py::arg("maximize") = false,
        py::arg("keep_grads") = false);
  m.def("fused_autograd", &fused_autograd,
        "fused_forward as ONE native autograd node (backward = fused_backward): returns mu'.., nu'.., u..",
        py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("b1"), py::arg("b2"), py::arg("eps"),
        py::arg("eps_root"), py::arg("count"));
  m.def("step_autograd", &step_autograd,
        "fused_step as ONE native autograd node (backward = fused_step_backward): returns p'.., mu'.., nu'..",
        py::arg("params"), py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("b1"), py::arg("b2"),
        py::arg("eps"), py::arg("eps_root"), py::arg("step_size"), py::arg("count"), py::arg("wd_l2") = 0.0,
        py::arg("wd_decoupled") = 0.0, py::arg("maximize") = false);
  m.def("flat_step", &flat_step,
        "whole Adam step with FLAT moment buffers (one node, 2K+2 autograd inputs): returns (new_params, mu', nu')",
        py::arg("params"), py::arg("updates"), py::arg("mu_flat"), py::arg("nu_flat"), py::arg("offsets"),
        py::arg("count"), py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("step_size"),
        py::arg("wd_l2") = 0.0, py::arg("wd_decoupled") = 0.0, py::arg("maximize") = false,
        py::arg("inplace") = false, py::arg("keep_grads") = false, py::arg("count_dev") = py::none(),
        py::arg("tables") = std::vector<Tensor>());
  m.def("bias_tables_ex",
        [](int dtype_family, double b1, double b2, double eps_root) {
          const int64_t len = b200_adam_bias_table_len_ex(dtype_family, b1, b2, eps_root);
          if (len <= 0)
            throw std::runtime_error("bias_tables_ex: betas this close to 1 need more than 2^24 table entries");
          auto opts = at::TensorOptions().dtype(dtype_family == B200_ADAM_F64 ? at::kDouble : at::kFloat).device(at::kCPU);
          std::vector<Tensor> t{at::empty({len}, opts), at::empty({len}, opts), at::empty({len}, opts), at::empty({len}, opts)};
          check(b200_adam_bias_table_ex(dtype_family, b1, b2, eps_root, len, t[0].data_ptr(), t[1].data_ptr(),
                                        t[2].data_ptr(), t[3].data_ptr()),
                "bias_tables_ex");
          return t;
        },
        "host-computed bias-correction tables [c1, c2, o1, c2e] for the capturable differentiable step",
        py::arg("dtype_family"), py::arg("b1"), py::arg("b2"), py::arg("eps_root"));
  m.def("host_step", &host_step, "Fused forward+backward over pinned HOST tensors (pipelined H2D/compute/D2H)",
        py::arg("updates"), py::arg("mu"), py::arg("nu"), py::arg("dupdates"), py::arg("dmu_ext"), py::arg("dnu_ext"),
        py::arg("out"), py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("eps_root"), py::arg("count"),
        py::arg("slice_elems") = 0, py::arg("allow_pageable") = false);
  // -- utilities
  mod.def("kernel_launches", []() { return b200_adam_kernel_launches(); });
  mod.def("abi_version", []() { return b200_adam_abi_version(); });
  mod.def("set_tuning", [](int ctas_per_sm, int tiles_per_chunk) {
    check(b200_adam_set_tuning(ctas_per_sm, tiles_per_chunk), "set_tuning");
  }, py::arg("ctas_per_sm") = 0, py::arg("tiles_per_chunk") = 0);
  mod.def("query_launch", [](int dtype, bool backward, int64_t n) {
    int32_t out[5];
    check(b200_adam_query_launch(dtype, backward ? 1 : 0, n, out), "query_launch");
    return py::dict(py::arg("grid") = out[0], py::arg("block") = out[1], py::arg("chunk_elems") = out[2],
                    py::arg("vector_elems") = out[3], py::arg("unroll") = out[4]);
  }, py::arg("dtype"), py::arg("backward"), py::arg("n"));
  mod.def("host_release", []() { b200_adam_host_release(); });
}
