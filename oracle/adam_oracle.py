"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracles for the accelerated Adam operator.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product package ``torchopt_b200`` never does (a test greps
for it).

Three oracles live here:

* ``COracle``      -- ctypes binding of ``oracle/adam_oracle.c`` (``liboracle_adam.so``): the C restatement.
* ``NumpyOracle``  -- an independent NumPy restatement of the same arithmetic (float32/float64 IEEE ops in
                      source order; ``np.sqrt`` is correctly rounded, ``torch.sqrt`` on CPU is not).
* ``load_ref()``   -- the reference's own compiled ``adam_op`` (``oracle/_ref/_C.so``, built from the
                      unmodified sources by ``oracle/build_ref.py``).  Needs torch; not available if the
                      .so was never built.

Parity status: PINNED -- ``tests/test_oracle.py`` checks COracle and NumpyOracle bit-for-bit against the
golden vectors in ``tests/golden`` (generated from ``oracle/_ref`` by ``tests/golden/make_golden.py``) and
against ``oracle/_ref`` live when present.

All functions take/return flat NumPy arrays and follow the reference signatures
(/root/reference/torchopt/_C/adam_op.pyi:20-62; kernels /root/reference/src/adam_op/adam_op_impl_cpu.cpp).
"""
from __future__ import annotations

import ctypes
import importlib.util
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
C_SRC = os.path.join(HERE, "adam_oracle.c")
C_LIB = os.path.join(HERE, "liboracle_adam.so")
REF_SO = os.path.join(HERE, "_ref", "_C.so")

_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def build_c_oracle(force: bool = False) -> str:
    """gcc-compile adam_oracle.c (no FMA contraction, no -march) into liboracle_adam.so."""
    if not force and os.path.isfile(C_LIB) and os.path.getmtime(C_LIB) >= os.path.getmtime(C_SRC):
        return C_LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", C_SRC, "-o", C_LIB,
           "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp", "-lm"]
    subprocess.check_call(cmd)
    return C_LIB


# ----------------------------------------------------------------------------------------------
# host scalars (adam_op_impl_cpu.cpp:72-74, 190-192, 327-329) -- math.pow is libm pow on doubles
# ----------------------------------------------------------------------------------------------
def host_scalars(dtype, b1: float, b2: float, eps: float, eps_root: float, count: int) -> dict:
    T = np.dtype(dtype).type
    pb1 = math.pow(b1, float(count))
    pb2 = math.pow(b2, float(count))
    B1, B2 = T(b1), T(b2)

    def inv(x):  # 1/x in double with IEEE semantics for x == 0 (count == 0 -> inf, like the C++)
        return math.inf if x == 0 else 1.0 / x

    return {
        "B1": B1, "B2": B2, "EPS": T(eps), "ER": T(eps_root),
        "OMB1": T(1) - B1, "OMB2": T(1) - B2, "TWO_OMB2": T(2) * (T(1) - B2),
        "C1": T(inv(1.0 - pb1)), "C2": T(inv(1.0 - pb2)),
        "O1": T(1.0 - pb1), "C2E": T(inv(1.0 - pb2 + eps_root)),
    }


class NumpyOracle:
    """NumPy restatement; every line is one rounded IEEE op in the array dtype (Appendix A of SURVEY.md)."""

    @staticmethod
    def forward_mu(updates, mu, b1):  # adam_op_impl_cpu.cpp:102-107
        s = host_scalars(mu.dtype, b1, 0.0, 0.0, 0.0, 1)
        with np.errstate(all="ignore"):
            return (s["B1"] * mu) + (s["OMB1"] * updates)

    @staticmethod
    def forward_nu(updates, nu, b2):  # adam_op_impl_cpu.cpp:136-142
        s = host_scalars(nu.dtype, 0.0, b2, 0.0, 0.0, 1)
        with np.errstate(all="ignore"):
            return (s["B2"] * nu) + ((s["OMB2"] * updates) * updates)

    @staticmethod
    def forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count):  # adam_op_impl_cpu.cpp:174-180
        s = host_scalars(new_mu.dtype, b1, b2, eps, eps_root, count)
        with np.errstate(all="ignore"):
            mu_hat = new_mu * s["C1"]
            nu_hat = new_nu * s["C2"]
            return mu_hat / (np.sqrt(nu_hat + s["ER"]) + s["EPS"])

    @staticmethod
    def forward_(updates, mu, nu, b1, b2, eps, eps_root, count):  # adam_op_impl_cpu.cpp:46-60
        """Returns (updates_out, mu_out, nu_out) as NEW arrays (the reference writes them in place)."""
        new_mu = NumpyOracle.forward_mu(updates, mu, b1)
        new_nu = NumpyOracle.forward_nu(updates, nu, b2)
        new_u = NumpyOracle.forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count)
        return new_u, new_mu, new_nu

    @staticmethod
    def backward_mu(dmu, b1):  # adam_op_impl_cpu.cpp:220-225
        s = host_scalars(dmu.dtype, b1, 0.0, 0.0, 0.0, 1)
        with np.errstate(all="ignore"):
            return s["OMB1"] * dmu, s["B1"] * dmu

    @staticmethod
    def backward_nu(dnu, updates, b2):  # adam_op_impl_cpu.cpp:257-263
        s = host_scalars(dnu.dtype, 0.0, b2, 0.0, 0.0, 1)
        with np.errstate(all="ignore"):
            return (s["TWO_OMB2"] * updates) * dnu, s["B2"] * dnu

    @staticmethod
    def backward_updates(dupdates, updates, new_mu, b1, b2, eps_root, count):  # adam_op_impl_cpu.cpp:298-316
        s = host_scalars(dupdates.dtype, b1, b2, 0.0, eps_root, count)
        T = dupdates.dtype
        with np.errstate(all="ignore"):
            r = updates / new_mu
            den = r * s["O1"]
            dmu = dupdates * r
            t = ((-dupdates) * updates) * den
            tail = ((t.astype(np.float64) * 0.5) * np.float64(s["C2E"])) * den.astype(np.float64)
            dnu = tail.astype(T)
        zero = new_mu == 0
        dmu = np.where(zero, T.type(0), dmu)
        dnu = np.where(zero, T.type(0), dnu)
        return dmu, dnu


class COracle:
    """ctypes binding of liboracle_adam.so (adam_oracle.c)."""

    def __init__(self, path: str | None = None):
        self.lib = ctypes.CDLL(path or build_c_oracle())
        self.lib.oracle_adam_max_threads.restype = ctypes.c_int

    @staticmethod
    def _p(a):
        return ctypes.c_void_p(a.ctypes.data)

    def _fn(self, name, dtype):
        return getattr(self.lib, f"oracle_adam_{name}_{_SUF[np.dtype(dtype)]}")

    def max_threads(self) -> int:
        return int(self.lib.oracle_adam_max_threads())

    def forward_(self, updates, mu, nu, b1, b2, eps, eps_root, count):
        """In place over the three arrays, like the reference; returns (updates, mu, nu)."""
        self._fn("forward_inplace", updates.dtype)(
            self._p(updates), self._p(mu), self._p(nu), ctypes.c_size_t(updates.size),
            ctypes.c_double(b1), ctypes.c_double(b2), ctypes.c_double(eps), ctypes.c_double(eps_root),
            ctypes.c_uint64(count))
        return updates, mu, nu

    def forward_mu(self, updates, mu, b1):
        out = np.empty_like(mu)
        self._fn("forward_mu", mu.dtype)(self._p(updates), self._p(mu), self._p(out),
                                         ctypes.c_size_t(mu.size), ctypes.c_double(b1))
        return out

    def forward_nu(self, updates, nu, b2):
        out = np.empty_like(nu)
        self._fn("forward_nu", nu.dtype)(self._p(updates), self._p(nu), self._p(out),
                                         ctypes.c_size_t(nu.size), ctypes.c_double(b2))
        return out

    def forward_updates(self, new_mu, new_nu, b1, b2, eps, eps_root, count):
        out = np.empty_like(new_mu)
        self._fn("forward_updates", new_mu.dtype)(
            self._p(new_mu), self._p(new_nu), self._p(out), ctypes.c_size_t(new_mu.size),
            ctypes.c_double(b1), ctypes.c_double(b2), ctypes.c_double(eps), ctypes.c_double(eps_root),
            ctypes.c_uint64(count))
        return out

    def backward_mu(self, dmu, b1):
        dg, dm = np.empty_like(dmu), np.empty_like(dmu)
        self._fn("backward_mu", dmu.dtype)(self._p(dmu), self._p(dg), self._p(dm),
                                           ctypes.c_size_t(dmu.size), ctypes.c_double(b1))
        return dg, dm

    def backward_nu(self, dnu, updates, b2):
        dg, dn = np.empty_like(dnu), np.empty_like(dnu)
        self._fn("backward_nu", dnu.dtype)(self._p(dnu), self._p(updates), self._p(dg), self._p(dn),
                                           ctypes.c_size_t(dnu.size), ctypes.c_double(b2))
        return dg, dn

    def backward_updates(self, dupdates, updates, new_mu, b1, b2, eps_root, count):
        dm, dn = np.empty_like(dupdates), np.empty_like(dupdates)
        self._fn("backward_updates", dupdates.dtype)(
            self._p(dupdates), self._p(updates), self._p(new_mu), self._p(dm), self._p(dn),
            ctypes.c_size_t(dupdates.size), ctypes.c_double(b1), ctypes.c_double(b2),
            ctypes.c_double(eps_root), ctypes.c_uint64(count))
        return dm, dn


# ----------------------------------------------------------------------------------------------
# Compositions used by the parity tests of the FUSED kernels (last paragraph):
# what the reference's three autograd Functions + the engine's accumulation adds compute per leaf.
# ----------------------------------------------------------------------------------------------
def fused_forward(o, g, mu, nu, b1, b2, eps, eps_root, count):
    """(new_mu, new_nu, new_updates) = MuOp, NuOp, UpdatesOp forward (accelerated_op/adam_op.py:168-180)."""
    new_mu = o.forward_mu(g, mu, b1)
    new_nu = o.forward_nu(g, nu, b2)
    new_u = o.forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count)
    return new_mu, new_nu, new_u


def fused_backward(o, du, u, new_mu, g, dmu_ext, dnu_ext, b1, b2, eps_root, count):
    """(dg, dmu, dnu): UpdatesOp.backward -> (+ external moment grads) -> MuOp.backward, NuOp.backward -> add.

    accelerated_op/adam_op.py:105-121, 53-60, 79-86; the adds are the autograd engine's accumulation.
    ``du`` None means no gradient reached the update (dmu' = dnu' = nothing to add).
    """
    zeros = np.zeros_like(g)
    if du is not None:
        dmu_new, dnu_new = o.backward_updates(du, u, new_mu, b1, b2, eps_root, count)
    else:
        dmu_new, dnu_new = None, None

    def acc(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return a + b

    dmu_tot = acc(dmu_ext, dmu_new)
    dnu_tot = acc(dnu_ext, dnu_new)
    dg_mu, dmu = o.backward_mu(dmu_tot, b1) if dmu_tot is not None else (None, None)
    dg_nu, dnu = o.backward_nu(dnu_tot, g, b2) if dnu_tot is not None else (None, None)
    dg = acc(dg_mu, dg_nu)
    return (dg if dg is not None else zeros), dmu, dnu


def fused_step_forward(o, p, g, mu, nu, b1, b2, eps, eps_root, count, step_size):
    """(new_p, new_mu, new_nu, us): the chain scale_by_accelerated_adam -> scale(step_size) -> apply_updates.
    ``g.mul(step_size)`` (transform/scale.py:103) and ``p.add(u)`` (update.py:79) are single rounded ops in the
    tensor dtype with the Python scalar rounded to that dtype once."""
    T = p.dtype.type
    new_mu, new_nu, u = fused_forward(o, g, mu, nu, b1, b2, eps, eps_root, count)
    with np.errstate(all="ignore"):
        us = u * T(step_size)
        new_p = p + us
    return new_p, new_mu, new_nu, us


def fused_step_backward(o, dp, new_mu, new_nu, g, dmu_ext, dnu_ext, b1, b2, eps, eps_root, count, step_size,
                        dus=None):
    """(dg, dmu, dnu) for the chain above: AddBackward hands dp to the scaled update (plus ``dus`` if the scaled
    update is used elsewhere), MulBackward multiplies by the rounded scalar, then the Adam backward."""
    T = g.dtype.type
    u = o.forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count)
    du = None
    if dp is not None or dus is not None:
        tot = dp if dus is None else (dus if dp is None else dp + dus)
        with np.errstate(all="ignore"):
            du = tot * T(step_size)
    return fused_backward(o, du, u, new_mu, g, dmu_ext, dnu_ext, b1, b2, eps_root, count)


def load_ref():
    """Import the reference's compiled adam_op (oracle/_ref/_C.so) -> module with .adam_op; None if absent."""
    if not os.path.isfile(REF_SO):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    spec = importlib.util.spec_from_file_location("_C", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
