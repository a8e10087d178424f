"""TEST INFRASTRUCTURE -- runs the *unmodified* reference solver without Drake.

``load_reference_ilqr()`` imports /root/reference/ilqr.py by path after seeding
``sys.modules['pydrake.all']`` with the two free functions the solver uses
(``InitializeAutoDiff``, ``ExtractGradient``; /root/reference/ilqr.py:254,268).
``ShimSystem`` duck-types the nine Drake methods the solver calls on its ``system``
(/root/reference/ilqr.py:37-58,223-229,259-265,725) on top of ``HostDynamics``.

Only usable where /root/reference exists (this container): used by the CPU tests that pin
``oracle/ilqr_port.py`` to the real reference and by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from .dynamics import HostDynamics

REFERENCE_DIR = os.environ.get("DDP_REFERENCE_DIR", "/root/reference")


class _GradVector(np.ndarray):
    """float vector that carries the Jacobian of the update that produced it."""
    grad = None


def InitializeAutoDiff(v):
    return np.asarray(v, dtype=np.float64)


def ExtractGradient(x_next):
    return x_next.grad


class _Vector:
    def __init__(self, ctx):
        self._ctx = ctx

    def size(self):
        return self._ctx.x.size

    def value(self):
        return self._ctx.x.reshape(-1, 1)

    def CopyToVector(self):
        out = self._ctx.x.view(_GradVector)
        out.grad = self._ctx.grad
        return out


class _State:
    def __init__(self, ctx):
        self._ctx = ctx

    def get_vector(self):
        return _Vector(self._ctx)


class _Context:
    def __init__(self, n, m):
        self.x = np.zeros(n)
        self.u = np.zeros(m)
        self.grad = None

    def get_discrete_state_vector(self):
        return _Vector(self)

    def SetDiscreteState(self, x):
        self.x = np.array(x, dtype=np.float64).reshape(-1)

    def get_discrete_state(self):
        return _State(self)


class _Port:
    def __init__(self, m):
        self._m = m

    def size(self):
        return self._m

    def FixValue(self, ctx, u):
        ctx.u = np.array(u, dtype=np.float64).reshape(-1)


class ShimSystem:
    def __init__(self, system, autodiff=False, dyn=None):
        self._system = system
        self._dyn = dyn or HostDynamics(system)
        self._ad = autodiff

    def IsDifferenceEquationSystem(self):
        return (True, self._system.dt)

    def CreateDefaultContext(self):
        return _Context(self._system.n, self._system.m)

    def get_input_port(self, idx):
        return _Port(self._system.m)

    def ToAutoDiffXd(self):
        return ShimSystem(self._system, autodiff=True, dyn=self._dyn)

    def CalcForcedDiscreteVariableUpdate(self, ctx, state):
        if self._ad:
            fx, fu = self._dyn.jac(ctx.x, ctx.u)
            ctx.grad = np.hstack([fx, fu])
        xn = self._dyn.step(ctx.x, ctx.u)
        if not np.all(np.isfinite(xn)):
            # Drake throws on a failed discrete update; the solver turns it into L = inf
            # (/root/reference/ilqr.py:317-323).
            raise RuntimeError("non-finite discrete update")
        ctx.x = xn

    def GetSubsystemByName(self, name):
        return self._system

    def time_step(self):
        return self._system.dt


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "ilqr.py"))


def _load(name, alias):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_DIR, name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cached = None


def load_reference_ilqr():
    """Return (ilqr_module, utils_module) of the reference, executed as-is."""
    global _cached
    if _cached is not None:
        return _cached
    assert reference_available(), f"{REFERENCE_DIR}/ilqr.py not found"
    shim_all = types.ModuleType("pydrake.all")
    shim_all.InitializeAutoDiff = InitializeAutoDiff
    shim_all.ExtractGradient = ExtractGradient
    shim_all.__all__ = ["InitializeAutoDiff", "ExtractGradient"]
    pkg = types.ModuleType("pydrake")
    pkg.all = shim_all
    saved = {k: sys.modules.get(k) for k in ("pydrake", "pydrake.all", "utils_derivs_interpolation")}
    try:
        sys.modules["pydrake"] = pkg
        sys.modules["pydrake.all"] = shim_all
        utils = _load("utils_derivs_interpolation.py", "reference_utils_derivs_interpolation")
        sys.modules["utils_derivs_interpolation"] = utils
        ilqr = _load("ilqr.py", "reference_ilqr")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = (ilqr, utils)
    return _cached
