// maintainer_adam_op.cpp -- TEST INFRASTRUCTURE: INTEGRATION.md section 2, compiled.
//
// What a TorchOpt maintainer's `src/adam_op/adam_op.cpp` looks like after the CUDA branch of its dispatcher has been
// re-pointed at the C ABI of libb200adam.so.  oracle/build_ref.py (`build_hybrid`) compiles THIS file together with the
// reference's own, unmodified `src/extension.cpp` and `src/adam_op/adam_op_impl_cpu.cpp` (where they lie under
// /root/reference) into `oracle/_ref/hybrid/_C.so`: TorchOpt's module with TorchOpt's CPU kernels behind `is_cpu()` and
// this repo's kernels behind `is_cuda()`.  tests/test_dropin_gpu.py runs the reference's own test files on it -- including
// `test_accelerated_op_is_available`, which asserts `accelerated_op_available('cpu')` and therefore cannot pass on the
// product module (`torchopt_b200/_C.so` has no CPU kernels by design).
//
// It defines exactly what the reference header declares (include/adam_op/adam_op.h:30-74): the seven dispatchers and
// `buildSubmodule`.  Nothing here is part of the product; the product binding is torchopt_b200/csrc/torch_shim.cpp.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <stdexcept>
#include <string>

#include "b200_adam.h"                          // this repo's C ABI
#include "include/adam_op/adam_op.h"            // reference declarations (resolved under /root/reference)
#include "include/adam_op/adam_op_impl_cpu.h"   // TorchOpt's own CPU launchers, kept as they are

namespace torchopt {
namespace adam_op {
namespace {

// everything one CUDA call needs besides its pointers
struct OnDevice {
  c10::cuda::CUDAGuard guard;
  void *stream;
  int dtype;
  OnDevice(const torch::Tensor &t, const char *op)
      : guard(t.device()), stream((void *)c10::cuda::getCurrentCUDAStream(t.device().index()).stream()) {
    if (t.scalar_type() == at::kFloat) {
      dtype = B200_ADAM_F32;
    } else if (t.scalar_type() == at::kDouble) {
      dtype = B200_ADAM_F64;
    } else {
      throw std::runtime_error(std::string("\"") + op + "\" not implemented for '" + c10::toString(t.scalar_type()) + "'");
    }
  }
  void done(int rc) const {
    if (rc != 0) throw std::runtime_error(b200_adam_error_string(rc));
  }
};

enum class Where { Cuda, Cpu };

Where placement(const torch::Tensor &t) {
  if (t.device().is_cuda()) return Where::Cuda;
  if (t.device().is_cpu()) return Where::Cpu;
  throw std::runtime_error("Not implemented");   // adam_op.cpp:48 -- relied on by tests/test_accelerated_op.py:37-40
}

int64_t count_of(const torch::Tensor &t) { return (int64_t)t.numel(); }

}  // namespace

TensorArray<3> adamForwardInplace(const torch::Tensor &updates, const torch::Tensor &mu, const torch::Tensor &nu,
                                  const pyfloat_t b1, const pyfloat_t b2, const pyfloat_t eps, const pyfloat_t eps_root,
                                  const pyuint_t count) {
  if (placement(updates) == Where::Cpu) return adamForwardInplaceCPU(updates, mu, nu, b1, b2, eps, eps_root, count);
  OnDevice dev(updates, "adamForwardInplaceCUDA");
  dev.done(b200_adam_forward_inplace(dev.dtype, updates.data_ptr(), mu.data_ptr(), nu.data_ptr(), count_of(updates), b1, b2,
                                     eps, eps_root, count, dev.stream));
  return TensorArray<3>{updates, mu, nu};
}

torch::Tensor adamForwardMu(const torch::Tensor &updates, const torch::Tensor &mu, const pyfloat_t b1) {
  if (placement(updates) == Where::Cpu) return adamForwardMuCPU(updates, mu, b1);
  OnDevice dev(updates, "adamForwardMuCUDA");
  auto out = torch::empty_like(mu);
  dev.done(b200_adam_forward_mu(dev.dtype, updates.data_ptr(), mu.data_ptr(), out.data_ptr(), count_of(updates), b1, dev.stream));
  return out;
}

torch::Tensor adamForwardNu(const torch::Tensor &updates, const torch::Tensor &nu, const pyfloat_t b2) {
  if (placement(updates) == Where::Cpu) return adamForwardNuCPU(updates, nu, b2);
  OnDevice dev(updates, "adamForwardNuCUDA");
  auto out = torch::empty_like(nu);
  dev.done(b200_adam_forward_nu(dev.dtype, updates.data_ptr(), nu.data_ptr(), out.data_ptr(), count_of(updates), b2, dev.stream));
  return out;
}

torch::Tensor adamForwardUpdates(const torch::Tensor &new_mu, const torch::Tensor &new_nu, const pyfloat_t b1,
                                 const pyfloat_t b2, const pyfloat_t eps, const pyfloat_t eps_root, const pyuint_t count) {
  if (placement(new_mu) == Where::Cpu) return adamForwardUpdatesCPU(new_mu, new_nu, b1, b2, eps, eps_root, count);
  OnDevice dev(new_mu, "adamForwardUpdatesCUDA");
  auto out = torch::empty_like(new_mu);
  dev.done(b200_adam_forward_updates(dev.dtype, new_mu.data_ptr(), new_nu.data_ptr(), out.data_ptr(), count_of(new_mu), b1,
                                     b2, eps, eps_root, count, dev.stream));
  return out;
}

TensorArray<2> adamBackwardMu(const torch::Tensor &dmu, const torch::Tensor &updates, const torch::Tensor &mu,
                              const pyfloat_t b1) {
  if (placement(dmu) == Where::Cpu) return adamBackwardMuCPU(dmu.contiguous(), updates, mu, b1);
  const auto grad = dmu.contiguous();   // adam_op.cpp:107: only the incoming gradient is made contiguous
  OnDevice dev(grad, "adamBackwardMuCUDA");
  auto dupdates = torch::empty_like(updates), dmu_prev = torch::empty_like(mu);
  dev.done(b200_adam_backward_mu(dev.dtype, grad.data_ptr(), dupdates.data_ptr(), dmu_prev.data_ptr(), count_of(grad), b1,
                                 dev.stream));
  return TensorArray<2>{dupdates, dmu_prev};
}

TensorArray<2> adamBackwardNu(const torch::Tensor &dnu, const torch::Tensor &updates, const torch::Tensor &nu,
                              const pyfloat_t b2) {
  if (placement(dnu) == Where::Cpu) return adamBackwardNuCPU(dnu.contiguous(), updates, nu, b2);
  const auto grad = dnu.contiguous();
  OnDevice dev(grad, "adamForwardNuCUDA");   // sic: the reference's dispatch name, adam_op_impl_cuda.cu:382
  auto dupdates = torch::empty_like(updates), dnu_prev = torch::empty_like(nu);
  dev.done(b200_adam_backward_nu(dev.dtype, grad.data_ptr(), updates.data_ptr(), dupdates.data_ptr(), dnu_prev.data_ptr(),
                                 count_of(grad), b2, dev.stream));
  return TensorArray<2>{dupdates, dnu_prev};
}

TensorArray<2> adamBackwardUpdates(const torch::Tensor &dupdates, const torch::Tensor &updates, const torch::Tensor &new_mu,
                                   const torch::Tensor &new_nu, const pyfloat_t b1, const pyfloat_t b2,
                                   const pyfloat_t eps_root, const pyuint_t count) {
  if (placement(dupdates) == Where::Cpu)
    return adamBackwardUpdatesCPU(dupdates.contiguous(), updates, new_mu, new_nu, b1, b2, eps_root, count);
  const auto grad = dupdates.contiguous();
  OnDevice dev(grad, "adamBackwardUpdatesCUDA");
  auto dmu = torch::empty_like(new_mu), dnu = torch::empty_like(new_nu);
  dev.done(b200_adam_backward_updates(dev.dtype, grad.data_ptr(), updates.data_ptr(), new_mu.data_ptr(), dmu.data_ptr(),
                                      dnu.data_ptr(), count_of(grad), b1, b2, eps_root, count, dev.stream));
  return TensorArray<2>{dmu, dnu};
}

// Python surface: names, keyword names (incl. backward_nu's `b1`) and order as the reference registers them
// (src/adam_op/adam_op.cpp:155-215) -- they are the contract torchopt/accelerated_op/adam_op.py calls through.
void buildSubmodule(py::module &mod) {  // NOLINT[runtime/references]
  using namespace pybind11::literals;
  py::module m = mod.def_submodule("adam_op", "Adam Ops");
  m.def("forward_", &adamForwardInplace, "Adam forward inplace", "updates"_a, "mu"_a, "nu"_a, "b1"_a, "b2"_a, "eps"_a,
        "eps_root"_a, "count"_a);
  m.def("forward_mu", &adamForwardMu, "Adam forward mu", "updates"_a, "mu"_a, "b1"_a);
  m.def("forward_nu", &adamForwardNu, "Adam forward nu", "updates"_a, "nu"_a, "b2"_a);
  m.def("forward_updates", &adamForwardUpdates, "Adam forward updates", "new_mu"_a, "new_nu"_a, "b1"_a, "b2"_a, "eps"_a,
        "eps_root"_a, "count"_a);
  m.def("backward_mu", &adamBackwardMu, "Adam backward mu", "dmu"_a, "updates"_a, "mu"_a, "b1"_a);
  m.def("backward_nu", &adamBackwardNu, "Adam backward nu", "dnu"_a, "updates"_a, "nu"_a, "b1"_a);
  m.def("backward_updates", &adamBackwardUpdates, "Adam backward updates", "dupdates"_a, "updates"_a, "new_mu"_a, "new_nu"_a,
        "b1"_a, "b2"_a, "eps_root"_a, "count"_a);
}

}  // namespace adam_op
}  // namespace torchopt
