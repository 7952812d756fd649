"""ctypes binding of the C ABI in ``include/b200_adam.h`` (libb200adam.so) -- raw device pointers in, int out.

This is the boundary the parity tests call through (``tests/ -m gpu``) and what ``bench.py`` times for the
kernel-only number.  It needs no torch: any owner of device memory can call it with integers that are device
addresses.  ``SYMBOLS`` lists every function the header declares (checked by tests/test_abi.py).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200adam.so")

F32, F64, BF16_F32 = 0, 1, 2
DTYPE_NAMES = {F32: "f32", F64: "f64", BF16_F32: "bf16_f32"}

_vpp = POINTER(c_void_p)
_i64p = POINTER(c_int64)
_u64p = POINTER(c_uint64)

# name -> (restype, argtypes); mirrors include/b200_adam.h declaration by declaration
SYMBOLS = {
    "b200_adam_forward_inplace": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double,
                                          c_double, c_uint64, c_void_p]),
    "b200_adam_forward_mu": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "b200_adam_forward_nu": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "b200_adam_forward_updates": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double,
                                          c_double, c_uint64, c_void_p]),
    "b200_adam_backward_mu": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "b200_adam_backward_nu": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "b200_adam_backward_updates": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double,
                                           c_double, c_double, c_uint64, c_void_p]),
    "b200_adam_fused_forward": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _i64p, _u64p, c_double,
                                        c_double, c_double, c_double, c_void_p]),
    "b200_adam_forward_inplace_mt": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _i64p, _u64p, c_double, c_double,
                                             c_double, c_double, c_void_p]),
    "b200_adam_fused_backward": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _i64p,
                                         _u64p, c_double, c_double, c_double, c_void_p]),
    "b200_adam_step_forward": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _i64p, _u64p,
                                       c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_int,
                                       c_void_p]),
    "b200_adam_step_inplace": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _i64p, _u64p, c_double, c_double,
                                       c_double, c_double, c_double, c_double, c_double, c_int, c_void_p]),
    "b200_adam_step_inplace_ex": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _i64p, _u64p, c_double, c_double,
                                          c_double, c_double, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    "b200_adam_bias_table_len": (c_int64, [c_int, c_double, c_double]),
    "b200_adam_bias_table": (c_int, [c_int, c_double, c_double, c_int64, c_void_p, c_void_p]),
    "b200_adam_step_inplace_capturable": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _i64p, _vpp, c_void_p,
                                                  c_void_p, c_int64, c_double, c_double, c_double, c_double, c_double,
                                                  c_double, c_double, c_int, c_int, c_void_p]),
    "b200_adam_bias_table_len_ex": (c_int64, [c_int, c_double, c_double, c_double]),
    "b200_adam_bias_table_ex": (c_int, [c_int, c_double, c_double, c_double, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "b200_adam_step_forward_capturable": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _i64p,
                                                  _vpp, c_void_p, c_void_p, c_int64, c_double, c_double, c_double,
                                                  c_double, c_double, c_double, c_double, c_int, c_void_p]),
    "b200_adam_step_backward_capturable": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp,
                                                   _vpp, _vpp, _vpp, _i64p, _vpp, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_int64, c_double, c_double, c_double, c_double, c_double, c_double,
                                                   c_double, c_int, c_void_p]),
    "b200_adam_prepare_fused": (c_void_p, [c_int, c_int64] + [_vpp] * 12 + [_i64p, _u64p, c_double, c_double, c_double,
                                                                       c_double]),
    "b200_adam_launch_prepared": (c_int, [c_void_p, c_int, c_void_p]),
    "b200_adam_release_prepared": (c_int, [c_void_p]),
    "b200_adam_step_backward": (c_int, [c_int, c_int64, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp, _vpp,
                                        _vpp, _vpp, _i64p, _u64p, c_double, c_double, c_double, c_double, c_double,
                                        c_double, c_double, c_int, c_void_p]),
    "b200_adam_host_step_f32": (c_int, [c_void_p] * 12 + [c_int64, c_double, c_double, c_double, c_double, c_uint64,
                                                          c_int64]),
    "b200_adam_host_release": (c_int, []),
    "b200_adam_abi_version": (c_int, []),
    "b200_adam_error_string": (c_char_p, [c_int]),
    "b200_adam_kernel_launches": (c_uint64, []),
    "b200_adam_set_tuning": (c_int, [c_int, c_int]),
    "b200_adam_set_pdl": (c_int, [c_int]),
    "b200_adam_query_launch": (c_int, [c_int, c_int, c_int64, POINTER(c_int32)]),
}


class B200AdamError(RuntimeError):
    pass


_lib = None


def load(path: str | None = None) -> ctypes.CDLL:
    """dlopen libb200adam.so and type every symbol; raises (never falls back) when the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise B200AdamError(f"{p} not found: build it with __graft_entry__.build(); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = "b200_adam") -> None:
    if rc != 0:
        raise B200AdamError(f"{what}: {load().b200_adam_error_string(rc).decode()} (code {rc})")


_ARR_CACHE: dict = {}
_ARR_CACHE_MAX = 512


def _arr(ctype, values):
    """ctypes array of ``values``; cached by content (the same leaf set is marshalled once, not once per call -- the
    arrays are only read by the library)."""
    key = (ctype, tuple(values))
    hit = _ARR_CACHE.get(key)
    if hit is None:
        if len(_ARR_CACHE) >= _ARR_CACHE_MAX:
            _ARR_CACHE.clear()
        hit = _ARR_CACHE[key] = (ctype * len(key[1]))(*key[1])
    return hit


def _ptrs(ptrs):
    return _arr(c_void_p, [None if (p is None or p == 0) else int(p) for p in ptrs])


# ---- thin wrappers taking integer device addresses -------------------------------------------------
def forward_inplace(dtype, updates, mu, nu, n, b1, b2, eps, eps_root, count, stream=0):
    check(load().b200_adam_forward_inplace(dtype, updates, mu, nu, n, b1, b2, eps, eps_root, count, stream), "forward_")


def forward_mu(dtype, updates, mu, mu_out, n, b1, stream=0):
    check(load().b200_adam_forward_mu(dtype, updates, mu, mu_out, n, b1, stream), "forward_mu")


def forward_nu(dtype, updates, nu, nu_out, n, b2, stream=0):
    check(load().b200_adam_forward_nu(dtype, updates, nu, nu_out, n, b2, stream), "forward_nu")


def forward_updates(dtype, new_mu, new_nu, updates_out, n, b1, b2, eps, eps_root, count, stream=0):
    check(load().b200_adam_forward_updates(dtype, new_mu, new_nu, updates_out, n, b1, b2, eps, eps_root, count, stream),
          "forward_updates")


def backward_mu(dtype, dmu, dupdates_out, dmu_out, n, b1, stream=0):
    check(load().b200_adam_backward_mu(dtype, dmu, dupdates_out, dmu_out, n, b1, stream), "backward_mu")


def backward_nu(dtype, dnu, updates, dupdates_out, dnu_out, n, b2, stream=0):
    check(load().b200_adam_backward_nu(dtype, dnu, updates, dupdates_out, dnu_out, n, b2, stream), "backward_nu")


def backward_updates(dtype, dupdates, updates, new_mu, dmu_out, dnu_out, n, b1, b2, eps_root, count, stream=0):
    check(load().b200_adam_backward_updates(dtype, dupdates, updates, new_mu, dmu_out, dnu_out, n, b1, b2, eps_root,
                                            count, stream), "backward_updates")


def fused_forward(dtype, g, mu, nu, mu_out, nu_out, u_out, n, count, b1, b2, eps, eps_root, stream=0):
    """All array arguments are sequences (one entry per leaf) of device addresses; n / count sequences of ints."""
    L = len(n)
    check(load().b200_adam_fused_forward(dtype, L, _ptrs(g), _ptrs(mu), _ptrs(nu), _ptrs(mu_out), _ptrs(nu_out),
                                         _ptrs(u_out), _arr(c_int64, n), _arr(c_uint64, count), b1, b2, eps, eps_root,
                                         stream), "fused_forward")


def forward_inplace_mt(dtype, updates, mu, nu, n, count, b1, b2, eps, eps_root, stream=0):
    L = len(n)
    check(load().b200_adam_forward_inplace_mt(dtype, L, _ptrs(updates), _ptrs(mu), _ptrs(nu), _arr(c_int64, n),
                                              _arr(c_uint64, count), b1, b2, eps, eps_root, stream), "forward_ (mt)")


def fused_backward(dtype, du, u, new_mu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n, count, b1, b2, eps_root, stream=0):
    L = len(n)
    check(load().b200_adam_fused_backward(dtype, L, _ptrs(du), _ptrs(u), _ptrs(new_mu), _ptrs(g), _ptrs(dmu_ext),
                                          _ptrs(dnu_ext), _ptrs(dg), _ptrs(dmu), _ptrs(dnu), _arr(c_int64, n),
                                          _arr(c_uint64, count), b1, b2, eps_root, stream), "fused_backward")


def step_forward(dtype, p, g, mu, nu, p_out, mu_out, nu_out, us_out, n, count, b1, b2, eps, eps_root, step_size,
                 stream=0, wd_l2=0.0, wd_decoupled=0.0, maximize=False):
    """Recipe + Adam + scale(step_size) + apply_updates for all leaves in one launch; ``us_out`` may be None."""
    L = len(n)
    check(load().b200_adam_step_forward(dtype, L, _ptrs(p), _ptrs(g), _ptrs(mu), _ptrs(nu), _ptrs(p_out),
                                        _ptrs(mu_out), _ptrs(nu_out), None if us_out is None else _ptrs(us_out),
                                        _arr(c_int64, n), _arr(c_uint64, count), b1, b2, eps, eps_root, step_size,
                                        wd_l2, wd_decoupled, int(maximize), stream), "step_forward")


def step_inplace(dtype, p, g, mu, nu, n, count, b1, b2, eps, eps_root, step_size, stream=0, wd_l2=0.0,
                 wd_decoupled=0.0, maximize=False, overwrite_grads=None):
    """``overwrite_grads=None`` calls b200_adam_step_inplace (g <- scaled update, the reference's side effect);
    True / False call b200_adam_step_inplace_ex with that choice."""
    L = len(n)
    if overwrite_grads is None:
        check(load().b200_adam_step_inplace(dtype, L, _ptrs(p), _ptrs(g), _ptrs(mu), _ptrs(nu), _arr(c_int64, n),
                                            _arr(c_uint64, count), b1, b2, eps, eps_root, step_size, wd_l2,
                                            wd_decoupled, int(maximize), stream), "step_inplace")
    else:
        check(load().b200_adam_step_inplace_ex(dtype, L, _ptrs(p), _ptrs(g), _ptrs(mu), _ptrs(nu), _arr(c_int64, n),
                                               _arr(c_uint64, count), b1, b2, eps, eps_root, step_size, wd_l2,
                                               wd_decoupled, int(maximize), int(bool(overwrite_grads)), stream),
              "step_inplace_ex")


def step_backward(dtype, dp, dus, new_mu, new_nu, g, dmu_ext, dnu_ext, dg, dmu, dnu, n, count, b1, b2, eps, eps_root,
                  step_size, stream=0, p=None, dp_out=None, wd_l2=0.0, wd_decoupled=0.0, maximize=False):
    L = len(n)
    check(load().b200_adam_step_backward(dtype, L, _ptrs(dp), None if dus is None else _ptrs(dus), _ptrs(new_mu),
                                         _ptrs(new_nu), _ptrs(g), _ptrs(dmu_ext), _ptrs(dnu_ext), _ptrs(dg),
                                         _ptrs(dmu), _ptrs(dnu), None if p is None else _ptrs(p),
                                         None if dp_out is None else _ptrs(dp_out), _arr(c_int64, n),
                                         _arr(c_uint64, count), b1, b2, eps, eps_root, step_size, wd_l2, wd_decoupled,
                                         int(maximize), stream), "step_backward")


class PreparedFused:
    """Argument block of a repeated fused forward + backward pair.  When the backward reads exactly what the forward
    writes (u = u_out, new_mu = mu_out -- the usual case) the block lives on the LIBRARY side
    (b200_adam_prepare_fused) and every launch is a two-argument call; otherwise the pre-marshalled ctypes arrays are
    passed on every call.  Same kernels and results either way."""

    handle = None     # instances assembled by hand (tools/tune_sweep.py) take the per-call path

    def __init__(self, dtype, leaves_fwd, leaves_bwd, n, count, b1, b2, eps, eps_root):
        self.lib = load()
        self.dtype, self.L = dtype, len(n)
        self.n, self.count = _arr(c_int64, n), _arr(c_uint64, count)
        self.fwd = [_ptrs(col) for col in leaves_fwd]   # g mu nu mu_out nu_out u_out
        self.bwd = [_ptrs(col) for col in leaves_bwd]   # du u new_mu g dmu_ext dnu_ext dg dmu dnu
        self.h = (b1, b2, eps, eps_root)
        self.handle = None
        g, mu, nu, mu_out, nu_out, u_out = (list(c) for c in leaves_fwd)
        du, u, new_mu, g_b, dme, dne, dg, dmu, dnu = (list(c) for c in leaves_bwd)
        if u == u_out and new_mu == mu_out and g_b == g:
            f, b = self.fwd, self.bwd
            self.handle = self.lib.b200_adam_prepare_fused(dtype, self.L, f[0], f[1], f[2], f[3], f[4], f[5], b[0], b[4],
                                                           b[5], b[6], b[7], b[8], self.n, self.count, b1, b2, eps, eps_root)
        self._launch = self.lib.b200_adam_launch_prepared

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.b200_adam_release_prepared(self.handle)
            self.handle = None

    def forward(self, stream=0):
        if self.handle:
            return self._launch(self.handle, 0, stream)
        b1, b2, eps, er = self.h
        return self.lib.b200_adam_fused_forward(self.dtype, self.L, *self.fwd, self.n, self.count, b1, b2, eps, er,
                                                stream)

    def backward(self, stream=0):
        if self.handle:
            return self._launch(self.handle, 1, stream)
        b1, b2, _, er = self.h
        return self.lib.b200_adam_fused_backward(self.dtype, self.L, *self.bwd, self.n, self.count, b1, b2, er, stream)


def host_step_f32(g, mu, nu, du, dmu_ext, dnu_ext, mu_out, nu_out, u_out, dg, dmu, dnu, n, b1, b2, eps, eps_root,
                  count, slice_elems=0):
    """All twelve arguments are HOST addresses (ideally page-locked)."""
    check(load().b200_adam_host_step_f32(g, mu, nu, du, dmu_ext, dnu_ext, mu_out, nu_out, u_out, dg, dmu, dnu, n, b1,
                                         b2, eps, eps_root, count, slice_elems), "host_step_f32")


def kernel_launches() -> int:
    return int(load().b200_adam_kernel_launches())


def set_tuning(ctas_per_sm: int = 0, tiles_per_chunk: int = 0) -> None:
    check(load().b200_adam_set_tuning(ctas_per_sm, tiles_per_chunk), "set_tuning")


def set_pdl(enabled: bool = True) -> None:
    check(load().b200_adam_set_pdl(int(bool(enabled))), "set_pdl")


def query_launch(dtype: int, backward: bool, n: int) -> dict:
    out = (c_int32 * 5)()
    check(load().b200_adam_query_launch(dtype, int(backward), n, out), "query_launch")
    return dict(grid=out[0], block=out[1], chunk_elems=out[2], vector_elems=out[3], unroll=out[4])
