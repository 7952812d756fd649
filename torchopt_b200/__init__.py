"""torchopt_b200 -- the accelerated differentiable Adam operator of metaopt/torchopt, rebuilt for NVIDIA B200.

One hot path only: ``torchopt._C.adam_op`` + ``torchopt.accelerated_op.AdamOp``, as
hand-written sm_100a CUDA kernels behind a C ABI (include/b200_adam.h).  Importing this package loads the native
extension and fails loudly if it has not been built -- there is no CPU or pure-Python fallback.

    torchopt_b200._C.adam_op            the reference's seven operators (+ fused multi-tensor entry points)
    torchopt_b200.accelerated_op        AdamOp, is_available
    torchopt_b200.transform             scale_by_accelerated_adam, ScaleByAdamState
    torchopt_b200.optim                 adam, apply_updates, MetaAdam, Adam   (thin callers)
    torchopt_b200.abi                   ctypes binding of the raw C ABI
    torchopt_b200.parallel              flat-shard partitioning, task-parallel meta-gradient all-reduce
    torchopt_b200.flat                  flat-buffer layout: FlatLayout, FlatMetaAdam, FlatAdam (single-range launches)
    torchopt_b200.utils                 extract_state_dict / recover_state_dict / stop_gradient (MAML loop helpers)
    torchopt_b200.graphs                CUDA-graph capture of launch-bound unrolls (inner loop + meta-backward)
"""
from torchopt_b200._native import _C, adam_op
from torchopt_b200 import abi, accelerated_op, flat, graphs, optim, parallel, transform, utils
from torchopt_b200.accelerated_op import AdamOp, AdamStep
from torchopt_b200.accelerated_op import is_available as accelerated_op_available
from torchopt_b200.flat import FlatAdam, FlatAdamW, FlatLayout, FlatMetaAdam, FlatMetaAdamW
from torchopt_b200.optim import (Adam, AdamW, FuncAdam, FuncAdamW, MetaAdam, MetaAdamW, adam, adamw,
                                 apply_updates)
from torchopt_b200.utils import ModuleState, extract_state_dict, recover_state_dict, stop_gradient
from torchopt_b200.transform import GradientTransformation, ScaleByAdamState, scale_by_accelerated_adam

__version__ = "0.1.0"

__all__ = [
    "_C", "adam_op", "abi", "accelerated_op", "flat", "graphs", "optim", "parallel", "transform",
    "AdamOp", "AdamStep", "FlatAdam", "FlatAdamW", "FlatLayout", "FlatMetaAdam", "FlatMetaAdamW", "accelerated_op_available", "Adam", "AdamW", "FuncAdam", "FuncAdamW", "MetaAdam", "MetaAdamW", "adamw", "adam", "apply_updates",
    "GradientTransformation", "ScaleByAdamState", "scale_by_accelerated_adam",
    "ModuleState", "extract_state_dict", "recover_state_dict", "stop_gradient", "utils",
]
