"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CPU adam_op into oracle/_ref/.

Compiles, where they lie under /root/reference (nothing is copied into this repo):
    src/extension.cpp, src/adam_op/adam_op.cpp, src/adam_op/adam_op_impl_cpu.cpp
with the reference's own flags (-O3 + OpenMP, no -march: CMakeLists.txt:37,48) into the CPython
extension ``oracle/_ref/_C.so`` (``PYBIND11_MODULE(_C, mod)``, src/extension.cpp:20).  The result
is (i) the primary parity oracle "O1", (ii) the generator of tests/golden/*.npz and (iii) the timed
OpenMP CPU baseline of ``bench.py --impl reference``.  It is never imported by the product package.

The reference's CMake build is NOT run (it hard-codes _GLIBCXX_USE_CXX11_ABI=0, which is wrong for the
torch in this image); this is a 3-file g++ recipe.  ``/root/reference`` only exists in the build
container; on the GPU box the prebuilt ``oracle/_ref/_C.so`` is used as is.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TORCHOPT_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "_C.so")
SOURCES = ["src/extension.cpp", "src/adam_op/adam_op.cpp", "src/adam_op/adam_op_impl_cpu.cpp"]


def reference_present() -> bool:
    return all(os.path.isfile(os.path.join(REF, s)) for s in SOURCES)


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(REF, s)) <= t for s in SOURCES)


def build(force: bool = False, verbose: bool = True) -> str | None:
    """Build oracle/_ref/_C.so; returns its path, or None when /root/reference is absent."""
    if not reference_present():
        return OUT if os.path.isfile(OUT) else None
    if not force and up_to_date():
        return OUT
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [REF, pybind11.get_include(), sysconfig.get_paths()["include"], *ce.include_paths()]
    objs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace("/", "_") + ".o")
        cmd = ["g++", "-std=c++17", "-O3", "-fopenmp", "-fPIC", "-w",
               f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
               "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=_C",
               *[f"-I{p}" for p in inc], "-c", os.path.join(REF, src), "-o", obj]
        if verbose:
            print("[oracle/_ref]", " ".join(cmd[:6]), "...", src, flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    link = ["g++", "-shared", *objs, "-o", OUT, f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}",
            "-ltorch_python", "-ltorch", "-ltorch_cpu", "-lc10",
            "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return OUT


CUDA_OUT_DIR = os.path.join(OUT_DIR, "cuda")
CUDA_OUT = os.path.join(CUDA_OUT_DIR, "_C.so")
CUDA_SOURCE = "src/adam_op/adam_op_impl_cuda.cu"


def build_cuda(force: bool = False, verbose: bool = True) -> str | None:
    """Optional comparison point "O4": the reference's UNMODIFIED CUDA adam_op (src/adam_op/adam_op_impl_cuda.cu,
    -D__USE_CUDA__) compiled as is for sm_100 -> oracle/_ref/cuda/_C.so.  Used only by tests/perf/leaf_sweep.py to time
    "the reference's generic kernels recompiled for B200" beside this repo's kernels.  ~4 minutes of nvcc."""
    srcs = SOURCES + [CUDA_SOURCE]
    if not all(os.path.isfile(os.path.join(REF, s)) for s in srcs):
        return CUDA_OUT if os.path.isfile(CUDA_OUT) else None
    if not force and os.path.isfile(CUDA_OUT):
        return CUDA_OUT
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(CUDA_OUT_DIR, exist_ok=True)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [REF, pybind11.get_include(), sysconfig.get_paths()["include"], *ce.include_paths(),
           "/usr/local/cuda/include"]
    defs = [f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-DTORCH_EXTENSION_NAME=_C", "-D__USE_CUDA__"]
    objs = []
    for src in srcs:
        obj = os.path.join(CUDA_OUT_DIR, src.replace("/", "_") + ".o")
        if src.endswith(".cu"):
            cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC",
                   "-Xcompiler", "-fopenmp", "-w", "--expt-relaxed-constexpr", *defs, *[f"-I{p}" for p in inc],
                   "-c", os.path.join(REF, src), "-o", obj]
        else:
            cmd = ["g++", "-std=c++17", "-O3", "-fopenmp", "-fPIC", "-w", *defs, *[f"-I{p}" for p in inc], "-c",
                   os.path.join(REF, src), "-o", obj]
        if verbose:
            print("[oracle/_ref/cuda]", cmd[0], "...", src, flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    link = ["g++", "-shared", *objs, "-o", CUDA_OUT, f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}",
            "-ltorch_python", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
            "-L/usr/local/cuda/lib64", "-lcudart", "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return CUDA_OUT


HYBRID_OUT_DIR = os.path.join(OUT_DIR, "hybrid")
HYBRID_OUT = os.path.join(HYBRID_OUT_DIR, "_C.so")
HYBRID_DISPATCHER = os.path.join(os.path.dirname(HERE), "tests", "dropin", "maintainer_adam_op.cpp")
HYBRID_REF_SOURCES = ["src/extension.cpp", "src/adam_op/adam_op_impl_cpu.cpp"]


def build_hybrid(force: bool = False, verbose: bool = True) -> str | None:
    """INTEGRATION.md section 2, compiled: the reference's UNMODIFIED module entry (src/extension.cpp) and CPU kernels
    (src/adam_op/adam_op_impl_cpu.cpp) + tests/dropin/maintainer_adam_op.cpp (our re-pointed dispatcher, standing in for the
    reference's src/adam_op/adam_op.cpp) linked against torchopt_b200/libb200adam.so -> oracle/_ref/hybrid/_C.so.
    This is the module a TorchOpt maintainer ships after the integration: TorchOpt's own kernels for CPU tensors, this repo's
    for CUDA tensors.  Test infrastructure (tests/test_dropin_gpu.py); the product never loads it."""
    pkg = os.path.join(os.path.dirname(HERE), "torchopt_b200")
    lib = os.path.join(pkg, "libb200adam.so")
    if not all(os.path.isfile(os.path.join(REF, s)) for s in HYBRID_REF_SOURCES) or not os.path.isfile(lib):
        return HYBRID_OUT if os.path.isfile(HYBRID_OUT) else None
    deps = [os.path.join(REF, s) for s in HYBRID_REF_SOURCES] + [HYBRID_DISPATCHER, lib,
                                                                os.path.join(os.path.dirname(HERE), "include", "b200_adam.h")]
    if not force and os.path.isfile(HYBRID_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HYBRID_OUT) for d in deps):
        return HYBRID_OUT
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(HYBRID_OUT_DIR, exist_ok=True)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [REF, os.path.join(os.path.dirname(HERE), "include"), pybind11.get_include(), sysconfig.get_paths()["include"],
           *ce.include_paths(), "/usr/local/cuda/include"]
    defs = [f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-DTORCH_EXTENSION_NAME=_C"]
    objs = []
    for src in [os.path.join(REF, s) for s in HYBRID_REF_SOURCES] + [HYBRID_DISPATCHER]:
        obj = os.path.join(HYBRID_OUT_DIR, os.path.basename(src) + ".o")
        cmd = ["g++", "-std=c++17", "-O3", "-fopenmp", "-fPIC", "-w", *defs, *[f"-I{p}" for p in inc], "-c", src, "-o", obj]
        if verbose:
            print("[oracle/_ref/hybrid] g++ ...", src, flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    link = ["g++", "-shared", *objs, "-o", HYBRID_OUT, f"-L{pkg}", "-lb200adam", "-Wl,-rpath,$ORIGIN:$ORIGIN/../../../torchopt_b200",
            f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}", "-ltorch_python", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10",
            "-lc10_cuda", "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return HYBRID_OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print(path if path else "reference sources not present and no prebuilt oracle/_ref/_C.so")
    if "--cuda" in sys.argv:
        print(build_cuda(force="--force" in sys.argv))
    if "--hybrid" in sys.argv:
        print(build_hybrid(force="--force" in sys.argv))
