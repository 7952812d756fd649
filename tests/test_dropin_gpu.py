"""Drop-in acceptance: the reference's OWN Python package and its OWN tests on top of this repo's native module.

``oracle/_ref/dropin.tar.gz`` (packed by tests/dropin/stage.py from /root/reference; git-ignored, travels to the GPU box)
holds an unmodified copy of ``torchopt/`` and of the reference test files that touch the Adam path.  Here it is unpacked
into a scratch directory, the
freshly built ``torchopt_b200/_C.so`` + ``libb200adam.so`` are placed inside that package as ``torchopt/_C.so`` -- the
exact swap INTEGRATION.md describes -- and the reference's tests run in a subprocess:

* ``tests/test_alias.py::test_adam_accelerated_cuda`` and ``tests/test_optim.py::test_Adam_accelerated_cuda`` (the
  reference's two CUDA tests, fp64, vs ``torch.optim.Adam(W)``) unchanged;
* the bodies of ``test_accelerated_op``, ``test_maml_accelerated_op``, ``test_adam(w)``, ``test_maml_adam``,
  ``test_Adam(W)``, ``test_maml_meta_adam`` on ``device='cuda'``: the test files are byte-identical, a generated
  conftest re-targets ``helpers.get_models`` (every body hard-codes ``'cpu'``) -- SURVEY.md section 4;
* the same with INTEGRATION.md section 3's two Python edits applied (``fused`` variant: this repo's ``AdamOp`` with the
  single fused node + ``scale_by_accelerated_adam`` calling ``op.multi``), so the one-launch path is proven inside real
  TorchOpt callers (transform/scale_by_adam.py:325-332).

``optree`` / ``graphviz`` are not installed in this image; ``tests/dropin/shims`` holds pure-Python stand-ins (test only).
Not run by this harness: ``test_accelerated_op_is_available`` -- it asserts ``accelerated_op_available('cpu')``, which is
False by design for a product without CPU kernels (choice ii); the CUDA / meta halves of that
assertion are checked in ``test_is_available_through_reference_python`` below.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHIVE = os.path.join(ROOT, "oracle", "_ref", "dropin.tar.gz")
SHIMS = os.path.join(ROOT, "tests", "dropin", "shims")
STAGE = ""     # scratch directory the archive is unpacked into (once per session, see _need_stage)
PKG = os.path.join(ROOT, "torchopt_b200")

pytestmark = pytest.mark.gpu

# test ids per run; parametrisations with use_accelerated_op(False) never reach the native module
CUDA_TESTS = ["test_alias.py::test_adam_accelerated_cuda", "test_optim.py::test_Adam_accelerated_cuda"]
RETARGETED = ["test_accelerated_op.py::test_accelerated_op", "test_accelerated_op.py::test_maml_accelerated_op",
              "test_alias.py::test_adam", "test_alias.py::test_adamw", "test_alias.py::test_maml_adam",
              "test_alias.py::test_empty", "test_optim.py::test_Adam", "test_optim.py::test_AdamW",
              "test_meta_optim.py::test_maml_meta_adam", "test_import.py::test_accelerated_op_import"]
NOT_ACCELERATED = "use_accelerated_op(False)"


def _need_stage() -> None:
    global STAGE
    if STAGE:
        return
    if not os.path.isfile(ARCHIVE):
        pytest.skip("drop-in archive missing: run __graft_entry__.build() where /root/reference is mounted")
    import atexit
    import tarfile
    import tempfile
    STAGE = tempfile.mkdtemp(prefix="b200_dropin_")
    atexit.register(shutil.rmtree, STAGE, ignore_errors=True)
    with tarfile.open(ARCHIVE) as tar:
        tar.extractall(STAGE)


HYBRID = os.path.join(ROOT, "oracle", "_ref", "hybrid", "_C.so")


def _install_native(variant: str) -> str:
    """Put the built native module inside the staged reference package: torchopt/_C.so (+ the library it links).
    Variant ``maintainer`` is the unmodified package with the MAINTAINER's module (oracle/_ref/hybrid/_C.so: the reference's
    own module entry and CPU kernels + tests/dropin/maintainer_adam_op.cpp, linked against libb200adam.so)."""
    pkg = os.path.join(STAGE, variant, "torchopt")
    if variant == "maintainer" and not os.path.isdir(pkg):
        shutil.copytree(os.path.join(STAGE, "plain", "torchopt"), pkg, ignore=shutil.ignore_patterns("*.so", "__pycache__"))
    for name in ("_C.so", "libb200adam.so"):
        src = HYBRID if (variant == "maintainer" and name == "_C.so") else os.path.join(PKG, name)
        assert os.path.isfile(src), f"{src} missing: the CUDA extension is not built (no fallback)"
        shutil.copy2(src, os.path.join(pkg, name))
    return os.path.dirname(pkg)


def _env(variant: str, device: str | None, cap: int, deselect: str = "") -> dict:
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([_install_native(variant), SHIMS])
    env["B200_DROPIN_MAX_PER_FUNC"] = str(cap)
    env["B200_DROPIN_EXPECT_PKG"] = os.path.join(STAGE, variant, "torchopt")
    env["B200_DROPIN_EXPECT_NATIVE"] = "maintainer" if variant == "maintainer" else "product"
    env["B200_DROPIN_DESELECT"] = deselect
    env.pop("B200_DROPIN_DEVICE", None)
    if device:
        env["B200_DROPIN_DEVICE"] = device
    return env


def _run_reference_tests(variant: str, ids: list, device: str | None, cap: int, deselect: str = "") -> int:
    tests_dir = os.path.join(STAGE, "reftests")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "--rootdir", tests_dir, "--confcutdir", tests_dir,
           "-p", "no:cacheprovider", "-W", "ignore::FutureWarning", *ids]
    res = subprocess.run(cmd, cwd=tests_dir, env=_env(variant, device, cap, deselect), capture_output=True, text=True,
                         timeout=1500)
    log = res.stdout + "\n" + res.stderr
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        kind = "retargeted" if device else ("cuda_tests" if ids == CUDA_TESTS else ids[0].split(".")[0])
        tag = f"dropin_{variant}_{kind}.log"
        with open(os.path.join(out_dir, tag), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
    tail = "\n".join(log.strip().splitlines()[-25:])
    assert res.returncode == 0, f"reference tests failed on the {variant} variant:\n{tail}"
    m = re.search(r"(\d+) passed", log)
    assert m, tail
    assert not re.search(r"\d+ (failed|error)", log), tail
    return int(m.group(1))


def test_reference_cuda_tests_on_new_native_module():
    """tests/test_alias.py:568 + tests/test_optim.py:365, every parametrisation, nothing re-targeted."""
    _need_stage()
    passed = _run_reference_tests("plain", CUDA_TESTS, device=None, cap=0)
    assert passed == 192 + 48, passed


def test_reference_adam_tests_retargeted_to_cuda():
    """The reference's CPU-parametrised Adam tests (forward parity vs its own non-accelerated Adam / torch.optim, and the
    assertion-free MAML unrolls that drive the backward operators) with models and batches on the B200."""
    _need_stage()
    passed = _run_reference_tests("plain", RETARGETED, device="cuda", cap=256, deselect=NOT_ACCELERATED)
    assert passed >= 900, passed


def test_reference_tests_with_fused_python_edits():
    """INTEGRATION.md section 3 applied to the reference's Python: one fused node, one launch for all leaves."""
    _need_stage()
    passed = _run_reference_tests("fused", CUDA_TESTS, device=None, cap=0)
    assert passed == 192 + 48, passed
    passed = _run_reference_tests("fused", RETARGETED, device="cuda", cap=128, deselect=NOT_ACCELERATED)
    assert passed >= 600, passed


def test_maintainer_build_runs_the_reference_suite_on_cpu_and_cuda():
    """INTEGRATION.md section 2, compiled and tested: the module a TorchOpt maintainer ships after re-pointing the CUDA
    branch of src/adam_op/adam_op.cpp at this repo's C ABI (TorchOpt's own, unmodified module entry and OpenMP kernels for
    CPU tensors; libb200adam for CUDA tensors).  On it the reference's test_accelerated_op.py passes IN FULL and unchanged --
    including test_accelerated_op_is_available (cpu True, meta False, cuda True), which the CPU-less product module cannot
    pass by design -- next to the reference's two CUDA tests."""
    _need_stage()
    if not os.path.isfile(HYBRID):
        pytest.skip("oracle/_ref/hybrid/_C.so missing: run __graft_entry__.build() where /root/reference is mounted")
    passed = _run_reference_tests("maintainer", ["test_accelerated_op.py", "test_import.py::test_accelerated_op_import"],
                                  device=None, cap=0)
    assert passed == 1 + 7 + 55 + 1, passed
    passed = _run_reference_tests("maintainer", CUDA_TESTS, device=None, cap=0)
    assert passed == 192 + 48, passed


def test_is_available_through_reference_python():
    """accelerated_op/__init__.py:30-50 with the new module: CUDA True, 'meta' False (the dispatcher throws, as the
    reference's own compiled extension does, tests/test_accelerated_op.py:37-40), 'cpu' False (no CPU kernels here)."""
    _need_stage()
    code = (
        "import torch, torchopt\n"
        "from torchopt._C import adam_op\n"
        "assert torchopt.accelerated_op_available('cuda') and torchopt.accelerated_op_available('cuda:0')\n"
        "assert torchopt.accelerated_op_available(0) and torchopt.accelerated_op_available(torch.device('cuda'))\n"
        "assert not torchopt.accelerated_op_available('meta')\n"
        "assert not torchopt.accelerated_op_available(['cuda', 'meta'])\n"
        "assert not torchopt.accelerated_op_available('cpu')\n"
        "import torchopt._C.adam_op\n"
        "print('launches', torchopt._C.kernel_launches())\n"
    )
    res = subprocess.run([sys.executable, "-c", code], env=_env("plain", None, 0), capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert int(res.stdout.split("launches")[1]) >= 4, res.stdout
