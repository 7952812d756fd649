"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200_adam.h
declares; the torch shim exposes the reference's operator surface; the product never touches oracle/."""
from __future__ import annotations

import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "b200_adam.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_adam_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_typed():
    from torchopt_b200 import abi
    lib = abi.load()
    syms = declared_symbols()
    assert len(syms) >= 17
    assert sorted(abi.SYMBOLS) == syms, "abi.SYMBOLS and include/b200_adam.h disagree"
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported by libb200adam.so"
    # nm agrees (dynamic symbol table, C linkage = unmangled names)
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (b200_adam_\w+)", out))
    assert set(syms) <= exported


def test_abi_version_and_error_strings_without_gpu():
    from torchopt_b200 import abi
    lib = abi.load()
    assert lib.b200_adam_abi_version() == 2
    assert lib.b200_adam_error_string(0) == b"ok"
    assert b"dtype" in lib.b200_adam_error_string(-1)
    assert b"argument" in lib.b200_adam_error_string(-2)
    # argument validation happens before any CUDA call: usable without a device
    assert lib.b200_adam_forward_mu(99, None, None, None, 4, 0.9, None) == -1
    assert lib.b200_adam_forward_mu(abi.F32, None, None, None, 0, 0.9, None) == 0
    assert lib.b200_adam_forward_mu(abi.F32, None, None, None, -3, 0.9, None) == -2
    assert lib.b200_adam_fused_forward(abi.F32, 0, None, None, None, None, None, None, None, None, 0.9, 0.999, 1e-8, 0.0,
                                       None) == 0
    assert lib.b200_adam_set_tuning(-1, 0) == -2
    assert lib.b200_adam_kernel_launches() == 0 or lib.b200_adam_kernel_launches() > 0


REFERENCE_SURFACE = {
    # name: keyword names in order -- /root/reference/src/adam_op/adam_op.cpp:157-214 (backward_nu's b1 is sic)
    "forward_": ["updates", "mu", "nu", "b1", "b2", "eps", "eps_root", "count"],
    "forward_mu": ["updates", "mu", "b1"],
    "forward_nu": ["updates", "nu", "b2"],
    "forward_updates": ["new_mu", "new_nu", "b1", "b2", "eps", "eps_root", "count"],
    "backward_mu": ["dmu", "updates", "mu", "b1"],
    "backward_nu": ["dnu", "updates", "nu", "b1"],
    "backward_updates": ["dupdates", "updates", "new_mu", "new_nu", "b1", "b2", "eps_root", "count"],
}


def test_shim_exposes_reference_operator_surface():
    import torch
    import torchopt_b200 as tb
    for name, kwargs in REFERENCE_SURFACE.items():
        fn = getattr(tb.adam_op, name)
        sig = fn.__doc__.splitlines()[0]
        got = re.findall(r"(\w+): ", sig)
        assert got == kwargs, f"{name}: keyword names {got} != reference {kwargs}"
    # multi-returns are Python lists (std::array<Tensor, N>), single returns a Tensor
    assert "list[torch.Tensor]" in tb.adam_op.forward_.__doc__
    assert "-> torch.Tensor" in tb.adam_op.forward_mu.__doc__
    # devices without a kernel: the dispatcher's RuntimeError('Not implemented') (adam_op.cpp:48) -- CPU included
    for dev in ("cpu", "meta"):
        t = torch.ones(3, device=dev)
        with pytest.raises(RuntimeError, match="Not implemented"):
            tb.adam_op.forward_mu(t, t, 0.9)
        with pytest.raises(RuntimeError, match="Not implemented"):
            tb.adam_op.forward_(t, t, t, 0.9, 0.999, 1e-8, 0.0, 1)
        assert tb.accelerated_op_available(dev) is False
    # reference import paths (tests/test_import.py:19-24)
    from torchopt_b200.accelerated_op.adam_op import AdamOp  # noqa: F401
    from torchopt_b200.accelerated_op import is_available  # noqa: F401
    assert AdamOp.MuOp and AdamOp.NuOp and AdamOp.UpdatesOp and AdamOp.FusedOp


def test_product_never_uses_the_oracle():
    """No file of the product package may import, load or mention oracle/ (a CPU fallback would void parity)."""
    pkg = os.path.join(ROOT, "torchopt_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"oracle|_src\.adam_op|liboracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_sass_is_sm100a_with_256bit_vector_accesses():
    """The shipped cubin is sm_100a only and the fused kernels use LDG.E.256 / STG.E.256 (no tensor-core or TMA
    instructions are expected on this path: it has no contraction and no reuse)."""
    from torchopt_b200 import abi
    elf = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf)
    sass = subprocess.run(["cuobjdump", "-sass", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "LDG.E.NA.ENL2.256.CONSTANT" in sass or re.search(r"LDG\.E\.[\w.]*256", sass)
    assert re.search(r"STG\.E\.[\w.]*256", sass)
    assert "LDL" not in sass and "STL" not in sass, "local-memory spills in the kernels"
