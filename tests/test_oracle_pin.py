"""Pins the CPU oracle (oracle/ilqr_port.py) to the reference:
 (a) against the UNMODIFIED /root/reference/ilqr.py run through the pydrake shim (only where
     the reference tree exists: this container), and
 (b) against the committed fixtures in tests/golden/ that were produced by that same
     reference (oracle/make_golden.py) -- this part also runs on the GPU box."""
import contextlib
import io
import os

import numpy as np
import pytest

from drake_ddp_b200 import problems
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation
from oracle import ilqr_port
from oracle.make_golden import CASES, SOLVE_CASES, run_reference, run_reference_solve
from oracle.pydrake_shim import reference_available
from tests.helpers import make_oracle, relerr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def oracle_iters(prob, kp, iters):
    o = make_oracle(prob, kp=kp)
    L, costs, eps, ls = np.inf, [], [], []
    for _ in range(iters):
        rec = o.iterate(L)
        L = rec.L
        costs.append(rec.L), eps.append(rec.eps), ls.append(rec.ls_iters)
    return o, np.array(costs), np.array(eps), np.array(ls)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_golden(name):
    factory, kp, iters = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    o, costs, eps, ls = oracle_iters(factory(), kp, iters)
    # tolerance: fp64 solver arithmetic, same algorithm, BLAS summation order may differ
    np.testing.assert_allclose(costs, g["costs"], rtol=1e-9)
    np.testing.assert_array_equal(eps, g["eps"])
    np.testing.assert_array_equal(ls, g["ls_iters"])
    assert relerr(o.x_bar.T, g["x_bar"]) < 1e-8
    assert relerr(o.u_bar.T, g["u_bar"]) < 1e-8
    assert relerr(o.K.transpose(1, 2, 0), g["K"]) < 1e-7
    assert relerr(o.kappa.T, g["kappa"]) < 1e-6
    assert relerr(o.fx.transpose(1, 2, 0), g["fx"]) < 1e-8
    assert relerr(o.fu.transpose(1, 2, 0), g["fu"]) < 1e-8
    assert abs(o.percentage_derivs - g["percentage_derivs"][-1]) < 1e-12


@needs_ref
@pytest.mark.parametrize("name", ["pendulum_N100", "acrobot_N40", "affine_4_1_adaptiveJerk",
                                  "affine_4_1_iterativeError", "quadruped_N30"])
def test_port_matches_live_reference(name):
    factory, kp, iters = CASES[name]
    ref = run_reference(factory(), kp, iters)
    o, costs, eps, ls = oracle_iters(factory(), kp, iters)
    np.testing.assert_allclose(costs, ref["costs"], rtol=1e-9)
    np.testing.assert_array_equal(ls, ref["ls_iters"])
    assert relerr(o.K.transpose(1, 2, 0), ref["K"]) < 1e-7


@needs_ref
def test_full_solve_matches_reference_pendulum():
    """Whole Solve() including the absolute stop rule (ilqr.py:692)."""
    from oracle.pydrake_shim import ShimSystem, load_reference_ilqr
    ref, _ = load_reference_ilqr()
    prob = problems.pendulum(100)
    r = ref.IterativeLinearQuadraticRegulator(ShimSystem(prob.system), prob.N, delta=prob.delta,
                                              beta=prob.beta, gamma=prob.gamma)
    r.SetInitialState(prob.x0.copy()); r.SetTargetState(prob.x_nom)
    r.SetRunningCost(prob.Q, prob.R); r.SetTerminalCost(prob.Qf); r.SetInitialGuess(prob.u_guess.copy())
    with contextlib.redirect_stdout(io.StringIO()) as f:
        x, u, _, L = r.Solve()
    n_iters = len(f.getvalue().strip().split("\n")) - 3
    o = make_oracle(prob)
    xo, uo, Lo = o.solve()
    assert len(o.trace) == n_iters
    assert abs(Lo - L) <= 1e-10 * abs(L)
    assert relerr(xo.T, x) < 1e-9


def solve_case_inputs(name):
    factory, kp, rows = SOLVE_CASES[name]
    prob = factory()
    kpc = prob.keypoints if kp == "problem" else kp
    nb = max(rows) + 1
    x0s = prob.batch_x0(nb, seed=0) if prob.sigma > 0 else np.repeat(prob.x0[None], nb, 0)
    return prob, kpc, rows, x0s


@pytest.mark.parametrize("name", sorted(SOLVE_CASES))
def test_port_full_solve_matches_golden(name):
    """Whole Solve() to convergence (ilqr.py:669-710, absolute stop rule :692) of the oracle port
    against the fixtures the UNMODIFIED reference produced for the same (x0, u_guess): same
    number of iterations, same line-search decisions, final cost within 1e-9 (north star: 1e-5),
    K, kappa within 1e-6 (north star: 1e-4); line-search failures (the reference's RuntimeError)
    happen at the same iteration."""
    prob, kp, rows, x0s = solve_case_inputs(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    for b in rows:
        o = make_oracle(prob, kp=kp, x0=x0s[b])
        failed = 0
        try:
            o.solve()
        except RuntimeError:
            failed = 1
        costs = np.array([r.L for r in o.trace])
        assert failed == int(g[f"failed_{b}"])
        assert len(costs) == len(g[f"costs_{b}"]), (b, len(costs), len(g[f"costs_{b}"]))
        np.testing.assert_array_equal([r.ls_iters for r in o.trace], g[f"ls_iters_{b}"])
        np.testing.assert_allclose(costs, g[f"costs_{b}"], rtol=1e-9)
        assert abs(costs[-1] - float(g[f"final_cost_{b}"])) <= 1e-9 * abs(float(g[f"final_cost_{b}"]))
        assert relerr(o.u_bar.T, g[f"u_bar_{b}"]) < 1e-6
        if f"K_{b}" in g.files:
            assert relerr(o.K.transpose(1, 2, 0), g[f"K_{b}"]) < 1e-6
            assert relerr(o.kappa.T, g[f"kappa_{b}"]) < 1e-6


@needs_ref
def test_full_solve_matches_live_reference_quadruped_n200():
    """The headline problem (C4: n=36, m=12, N=200) solved to convergence by the unmodified
    reference, live, against the port: identical iteration count, final cost 1e-9."""
    prob, kp, rows, x0s = solve_case_inputs("solve_quadruped_N200")
    b = 3
    ref = run_reference_solve(prob, kp, x0s[b])
    o = make_oracle(prob, kp=kp, x0=x0s[b])
    o.solve()
    assert len(o.trace) == len(ref["costs"]) and not ref["failed"]
    assert abs(o.trace[-1].L - ref["final_cost"]) <= 1e-9 * abs(ref["final_cost"])
    assert relerr(o.K.transpose(1, 2, 0), ref["K"]) < 1e-6


@needs_ref
def test_keypoint_known_answers():
    """Known-answer vectors captured from the reference code itself."""
    from oracle.pydrake_shim import ShimSystem, load_reference_ilqr
    ref, ref_utils = load_reference_ilqr()
    prob = problems.affine_sin(4, 1, 40)
    r = ref.IterativeLinearQuadraticRegulator(
        ShimSystem(prob.system), 40, derivs_keypoint_method=ref_utils.derivs_interpolation("setInterval", 5, 0, 0, 0))
    assert list(r.get_keypoints_set_interval()) == [0, 5, 10, 15, 20, 25, 30, 38]
    assert ilqr_port.keypoints_set_interval(40, 5) == [0, 5, 10, 15, 20, 25, 30, 38]
    # N=400, minN=5: 395 is dropped, 80 keypoints
    kp = ilqr_port.keypoints_set_interval(400, 5)
    assert len(kp) == 80 and kp[-3:] == [385, 390, 398]
    # jerk scan on a random trajectory: port == reference
    rng = np.random.default_rng(3)
    x = np.cumsum(rng.standard_normal((40, 4)) * 1e-2, axis=0)
    r.derivs_interpolation = ref_utils.derivs_interpolation("adaptiveJerk", 2, 10, 1e-4, 0)
    assert r.get_keypoints_adaptive_jerk(x.T, None) == ilqr_port.keypoints_adaptive_jerk(x, 2, 10, 1e-4)


def test_eps_table_counts():
    """Line-search candidate counts 27 / 175 / 360 for beta .5 / .9 / .95."""
    assert len(ilqr_port.eps_table(0.5)) == 27
    assert len(ilqr_port.eps_table(0.9)) == 175
    assert len(ilqr_port.eps_table(0.95)) == 360
    t = ilqr_port.eps_table(0.9)
    assert t[2] == 1.0 * 0.9 * 0.9          # repeated multiply, not 0.9**2


def test_lqr_special_case():
    """Linear dynamics: one iteration from any u_guess gives the finite-horizon Riccati gain and
    the second iteration improves by ~0."""
    from drake_ddp_b200 import systems

    class Lin:
        n, m = 3, 2
        A = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.0, -0.2, 0.9]])
        Bm = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.1]])

        def step(self, x, u):
            return self.A @ x + self.Bm @ u

        def jac(self, x, u):
            return self.A.copy(), self.Bm.copy()

    N = 25
    o = ilqr_port.IlqrOracle(Lin(), N, delta=1e-9, beta=0.5)
    Q, R, Qf = np.diag([1.0, 2.0, 0.5]), np.diag([0.1, 0.3]), np.diag([5.0, 5.0, 5.0])
    o.set_initial_state(np.array([1.0, -0.5, 0.3])); o.set_target_state(np.zeros(3))
    o.set_running_cost(Q, R); o.set_terminal_cost(Qf)
    o.set_initial_guess(np.random.default_rng(0).standard_normal((2, N - 1)))
    r1 = o.iterate(np.inf)
    P = Qf.copy()
    for t in range(N - 2, -1, -1):
        Kt = np.linalg.solve(R + Lin.Bm.T @ P @ Lin.Bm, Lin.Bm.T @ P @ Lin.A)
        np.testing.assert_allclose(o.K[t], Kt, rtol=1e-9, atol=1e-12)
        P = Q + Lin.A.T @ P @ (Lin.A - Lin.Bm @ Kt)
    r2 = o.iterate(r1.L)
    x0 = np.array([1.0, -0.5, 0.3])
    assert r2.eps == 1.0 and abs(r2.L - x0 @ P @ x0) < 1e-9 * r2.L     # optimal cost = x0' P0 x0
    try:   # at the optimum nothing improves: tiny improvement, or the reference's RuntimeError
        r3 = o.iterate(r2.L)
        assert abs(r3.improvement) < 1e-9 * r2.L
    except RuntimeError as e:
        assert "linesearch failed" in str(e)
