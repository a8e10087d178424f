"""TEST INFRASTRUCTURE -- CPU oracle dynamics.  Not part of the product path.

ctypes binding of ``oracle/_build/libhostmodels.so`` (host build of the analytic model
templates).  ``HostDynamics`` is what the oracle solver and the pydrake shim call where the
reference calls Drake (/root/reference/ilqr.py:208-272).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhostmodels.so")
_SRC = os.path.join(_HERE, "hostmodels.cpp")
_HDRS = [os.path.join(_HERE, "..", "drake_ddp_b200", "csrc", h)
         for h in ("models.h", "dual.h")]
_lib = None


def build(force: bool = False) -> str:
    """Compile the host model library with g++ (a few seconds)."""
    deps = [_SRC] + _HDRS
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(d) for d in deps if os.path.exists(d))):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    tmp = _LIB_PATH + f".{os.getpid()}.tmp"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", _SRC, "-o", tmp])
    os.replace(tmp, _LIB_PATH)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.hostmodel_dims.argtypes = [ctypes.c_int, ip, ip, ip]
        L.hostmodel_step.argtypes = [ctypes.c_int, dp, dp, dp, dp]
        L.hostmodel_jac.argtypes = [ctypes.c_int, dp, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class HostDynamics:
    """x+ = f(x,u) and its exact Jacobian on the host, for one AnalyticSystem."""

    def __init__(self, system):
        self.system = system
        self.n, self.m = system.n, system.m
        self.model_id = int(system.model_id)
        self.params = np.ascontiguousarray(system.params, dtype=np.float64)
        n, m, npar = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = lib().hostmodel_dims(self.model_id, ctypes.byref(n), ctypes.byref(m), ctypes.byref(npar))
        assert rc == 0, f"unknown model id {self.model_id}"
        assert (n.value, m.value) == (self.n, self.m)
        assert npar.value == self.params.size, (npar.value, self.params.size)
        self._L = lib()

    def step(self, x, u):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        xn = np.empty(self.n)
        self._L.hostmodel_step(self.model_id, _p(x), _p(u), _p(self.params), _p(xn))
        return xn

    def jac(self, x, u):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        xn = np.empty(self.n)
        fx = np.empty((self.n, self.n))
        fu = np.empty((self.n, self.m))
        self._L.hostmodel_jac(self.model_id, _p(x), _p(u), _p(self.params), _p(xn), _p(fx), _p(fu))
        return fx, fu
