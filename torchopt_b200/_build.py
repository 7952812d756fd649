"""In-tree build of the native pieces (run by ``__graft_entry__.build()``; no JIT cache, no pip install).

    libb200adam.so : nvcc, sm_100a only -- the CUDA kernels + the C ABI of include/b200_adam.h (no torch).
    _C.so          : g++ -- the pybind11/torch shim (module ``_C`` with submodule ``adam_op``), linked against
                     libb200adam.so through an $ORIGIN rpath.

Both land next to this file so they travel with the source tree (they are git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libb200adam.so")
SHIM = os.path.join(PKG, "_C.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # B200 only: no PTX for other archs, no multi-arch fatbin
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # belt and braces: the kernels already use explicit *_rn intrinsics
    "-Xcompiler", "-fPIC", "-shared",
]
LIB_SOURCES = ["b200_adam.cu", "host_pipeline.cu"]
LIB_DEPS = LIB_SOURCES + ["adam_kernels.cuh"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(cand):
        raise RuntimeError("nvcc not found: the B200 kernels cannot be built")
    return cand


def build_lib(force: bool = False, verbose: bool = True, extra_flags: list[str] | None = None) -> str:
    deps = [os.path.join(CSRC, f) for f in LIB_DEPS] + [os.path.join(INCLUDE, "b200_adam.h"), __file__]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *(extra_flags or []), f"-I{INCLUDE}", f"-I{CSRC}",
           *[os.path.join(CSRC, f) for f in LIB_SOURCES], "-o", LIB]
    if verbose:
        print("[torchopt_b200] " + " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


def build_shim(force: bool = False, verbose: bool = True) -> str:
    src = os.path.join(CSRC, "torch_shim.cpp")
    deps = [src, os.path.join(INCLUDE, "b200_adam.h"), LIB, __file__]
    if not force and _newer(SHIM, deps):
        return SHIM
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce

    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = [INCLUDE, pybind11.get_include(), sysconfig.get_paths()["include"], *ce.include_paths(),
           os.path.join(cuda_home, "include")]
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=_C",
           *[f"-I{p}" for p in inc], src, "-o", SHIM,
           f"-L{PKG}", "-lb200adam", "-Wl,-rpath,$ORIGIN",
           f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}",
           "-ltorch_python", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda"]
    if verbose:
        print("[torchopt_b200] " + " ".join(cmd[:8]) + " ... torch_shim.cpp", flush=True)
    subprocess.check_call(cmd)
    return SHIM


def build_tools(force: bool = False, verbose: bool = True) -> list[str]:
    """Stand-alone measurement programs under tools/*.cu (HBM ceiling, TMA staging experiment) -> build/<name>.
    Not part of the product; built here so that every CUDA source in the tree is compiled for sm_100a by one entry."""
    tools = os.path.join(ROOT, "tools")
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    built = []
    for name in sorted(f for f in os.listdir(tools) if f.endswith(".cu")):
        src, exe = os.path.join(tools, name), os.path.join(out_dir, name[:-3])
        if force or not _newer(exe, [src, os.path.join(CSRC, "adam_kernels.cuh")]):
            # linked against the product library (some tools time their variant against the product kernel)
            cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "--fmad=false", f"-I{INCLUDE}", f"-I{CSRC}", src, f"-L{PKG}", "-lb200adam",
                   # driver API (tools/guard_pages.cu: virtual memory management); the stub only satisfies the linker
                   f"-L{os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'lib64', 'stubs')}", "-lcuda",
                   "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../torchopt_b200", "-o", exe]
            if verbose:
                print("[torchopt_b200] " + " ".join(cmd), flush=True)
            try:
                subprocess.check_call(cmd)
            except subprocess.CalledProcessError as exc:   # a measurement tool must never fail the product build
                print(f"[torchopt_b200] WARNING: tool {name} did not build ({exc}); the product is unaffected", flush=True)
                continue
        built.append(exe)
    return built


def build(force: bool = False, verbose: bool = True) -> tuple[str, str]:
    out = build_lib(force, verbose), build_shim(force, verbose)
    build_tools(force, verbose)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
