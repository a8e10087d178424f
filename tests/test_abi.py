"""The C-ABI library loads and exports every symbol include/ddp_b200.h declares; host-only
entry points behave (no compute calls: there is no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

from drake_ddp_b200 import _lib, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if _lib.is_stale():
        _lib.build()
    return _lib.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "ddp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_loader_agree():
    assert header_functions() == sorted(_lib.SYMBOLS)


def test_every_declared_symbol_is_exported(L):
    for name in header_functions():
        assert hasattr(L, name), f"{name} declared in include/ddp_b200.h but not exported"


def test_model_dims(L):
    for prob in (problems.pendulum(), problems.acrobot(), problems.cart_pole(), problems.cart_pole_with_wall(),
                 problems.quadruped(), problems.quadruped_quat(), problems.arm_ball(),
                 problems.affine_sin(37, 12, 10)):
        n, m, npar = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert L.ddp_model_dims(prob.system.model_id, ctypes.byref(n), ctypes.byref(m), ctypes.byref(npar)) == 0
        assert (n.value, m.value, npar.value) == (prob.system.n, prob.system.m, prob.system.params.size)
    assert L.ddp_model_dims(999, ctypes.byref(n), ctypes.byref(m), ctypes.byref(npar)) < 0
    assert b"unknown model" in L.ddp_last_error()


def test_workspace_bytes(L):
    # C4: the seven trajectory arrays alone are 1024*(7200+2388+5184+2388+199+257904+85968)*8 B
    prob = problems.quadruped(200)
    nbytes = L.ddp_workspace_bytes(prob.system.model_id, 200, 1024, 1)
    arrays = 1024 * (200 * 36 + 199 * 12 + 199 * 12 * 36 + 199 * 12 + 199 + 199 * 36 * 36 + 199 * 36 * 12) * 8
    assert arrays < nbytes < 1.2 * arrays + (1 << 24)
    assert L.ddp_workspace_bytes(prob.system.model_id, 2, 1, 1) == 0      # N < 3
    assert L.ddp_workspace_bytes(999, 10, 1, 1) == 0


def test_constructor_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from drake_ddp_b200.ilqr import IterativeLinearQuadraticRegulator
    p = problems.pendulum()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        IterativeLinearQuadraticRegulator(p.system, p.N)


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (tier rule 3)."""
    pkg = os.path.join(ROOT, "drake_ddp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("CPU oracle", ""), fn
