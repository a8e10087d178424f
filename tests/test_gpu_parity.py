"""Parity of the CUDA path (called through the C ABI) against the CPU oracle and against the
reference-generated fixtures in tests/golden/.  Tolerances: the north star asks for 1e-5
relative on cost and 1e-4 on K, kappa; fp64 end to end gives far better, so the per-iteration
checks below use 1e-9 on cost and 1e-6 on gains, and discrete decisions (accepted eps,
ls_iters, keypoints) must be identical."""
import contextlib
import io
import os

import numpy as np
import pytest

from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation
from oracle import ilqr_port
from oracle.make_golden import CASES
from tests.helpers import gpu_state, make_gpu, make_oracle, relerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
COST_RTOL, GAIN_RTOL = 1e-9, 1e-6


def check_state(s, o, b=0, tol=GAIN_RTOL):
    x, u, K, kappa, dV, fx, fu = gpu_state(s, b)
    uscale = max(1.0, np.abs(o.u_bar).max())
    assert relerr(x, o.x_bar) < tol
    assert np.abs(u - o.u_bar).max() < tol * uscale
    assert relerr(K, o.K) < tol
    assert np.abs(kappa - o.kappa).max() < tol * max(uscale, np.abs(o.kappa).max())
    assert np.abs(dV - o.dV).max() < tol * max(1.0, np.abs(o.dV).max())
    assert relerr(fx, o.fx) < tol and relerr(fu, o.fu) < tol


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_reference_golden(name):
    factory, kp, iters = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = make_gpu(factory(), kp=kp)
    s.begin_solve()
    for i in range(iters):
        s.iterate()
        assert abs(s.cost[0] - g["costs"][i]) <= 1e-8 * abs(g["costs"][i])
        assert s.get(_lib.EPS)[0] == g["eps"][i]
        assert s.get_int(_lib.I_LS_ITERS)[0] == g["ls_iters"][i]
    x, u, K, kappa, dV, fx, fu = gpu_state(s)
    assert relerr(x.T, g["x_bar"]) < 1e-7
    assert relerr(K.transpose(1, 2, 0), g["K"]) < 1e-6         # north star: 1e-4
    assert np.abs(kappa.T - g["kappa"]).max() < 1e-6 * max(1.0, np.abs(g["u_bar"]).max())
    assert relerr(fx.transpose(1, 2, 0), g["fx"]) < 1e-7
    assert abs(100.0 * s.get_int(_lib.I_NUM_KEYPOINTS)[0] / (s.N - 1) - g["percentage_derivs"][-1]) < 1e-9


PER_ITER = [
    ("pendulum", lambda: problems.pendulum(100), None, 5),
    ("acrobot", lambda: problems.acrobot(40), None, 5),
    ("cart_pole", lambda: problems.cart_pole(100), None, 4),
    ("wall", lambda: problems.cart_pole_with_wall(100), None, 3),
    ("affine_6_2", lambda: problems.affine_sin(6, 2, 30), None, 3),
    ("affine_37_12", lambda: problems.affine_sin(37, 12, 30), None, 3),
    ("affine_36_12_jerk", lambda: problems.affine_sin(36, 12, 40), derivs_interpolation("adaptiveJerk", 2, 10, 1e-4, 0), 3),
    ("affine_27_7_ie", lambda: problems.affine_sin(27, 7, 40), derivs_interpolation("iterativeError", 2, 0, 0, 1e-9), 3),
    ("quadruped", lambda: problems.quadruped(50), None, 4),
    ("quadruped_quat_n37", lambda: problems.quadruped_quat(50), None, 4),
    ("quadruped_interval", lambda: problems.quadruped(50), derivs_interpolation("setInterval", 4, 0, 0, 0), 3),
    ("quadruped_ie", lambda: problems.quadruped(40), derivs_interpolation("iterativeError", 2, 0, 0, 1e-7), 3),
]


@pytest.mark.parametrize("name,factory,kp,iters", PER_ITER, ids=[p[0] for p in PER_ITER])
def test_gpu_matches_oracle_per_iteration(name, factory, kp, iters):
    prob = factory()
    s, o = make_gpu(prob, kp=kp), make_oracle(prob, kp=kp)
    s.begin_solve()
    L = np.inf
    for i in range(iters):
        s.iterate()
        rec = o.iterate(L)
        L = rec.L
        assert abs(s.cost[0] - rec.L) <= COST_RTOL * abs(rec.L), (i, s.cost[0], rec.L)
        assert s.get(_lib.EPS)[0] == rec.eps
        assert s.get_int(_lib.I_LS_ITERS)[0] == rec.ls_iters
        assert s.keypoints()[0] == rec.keypoints
        check_state(s, o)


@pytest.mark.parametrize("n,m,N", [(4, 1, 20), (6, 2, 25), (27, 7, 12), (36, 12, 10), (37, 12, 10)])
def test_backward_pass_teacher_forced(n, m, N):
    """Feed identical fx, fu, x_bar, u_bar and a DENSE (non-diagonal) cost to ilqr._backward_pass's
    restatement and to the kernel; compare K, kappa, dV."""
    rng = np.random.default_rng(n * 100 + m)
    prob = problems.affine_sin(n, m, N)
    T = N - 1

    def spd(k, scale):
        A = rng.standard_normal((k, k))
        return scale * (A @ A.T / k + np.eye(k))

    prob.Q, prob.R, prob.Qf = spd(n, 0.1), spd(m, 0.05), spd(n, 2.0)
    B = 3
    s = make_gpu(prob, B=B)
    fx = np.eye(n)[None, None] + 0.1 * rng.standard_normal((B, T, n, n))
    fu = 0.3 * rng.standard_normal((B, T, n, m))
    xb, ub = rng.standard_normal((B, N, n)), rng.standard_normal((B, T, m))
    for which, arr in ((_lib.FX, fx), (_lib.FU, fu), (_lib.X_BAR, xb), (_lib.U_BAR, ub)):
        s.put(which, arr)
    s.run_phase(_lib.PHASE_BACKWARD)
    for b in range(B):
        o = make_oracle(prob)
        o.fx, o.fu, o.x_bar, o.u_bar = fx[b].copy(), fu[b].copy(), xb[b].copy(), ub[b].copy()
        o.backward_pass()
        assert relerr(s.get(_lib.K)[b], o.K) < 1e-9
        assert relerr(s.get(_lib.KAPPA)[b], o.kappa) < 1e-9
        assert relerr(s.get(_lib.DV)[b], o.dV) < 1e-9


@pytest.mark.parametrize("name,A", [("pendulum", 8), ("quadruped", 4), ("affine", 27)])
def test_linesearch_teacher_forced(name, A):
    """Identical x_bar, u_bar, K, kappa, dV, L_last on both sides: candidate costs, expected
    improvements, the first-satisfying pick and the committed trajectory must agree."""
    rng = np.random.default_rng(7)
    prob = {"pendulum": lambda: problems.pendulum(40), "quadruped": lambda: problems.quadruped(30),
            "affine": lambda: problems.affine_sin(6, 2, 30)}[name]()
    n, m, N, T = prob.system.n, prob.system.m, prob.N, prob.N - 1
    o = make_oracle(prob)
    r0 = o.iterate(np.inf)                      # a sensible x_bar/u_bar/K/kappa/dV to start from
    o.kappa = o.kappa * 6.0                     # overshoot so eps=1 is rejected sometimes
    s = make_gpu(prob, A=A)
    for which, arr in ((_lib.X_BAR, o.x_bar), (_lib.U_BAR, o.u_bar), (_lib.K, o.K), (_lib.KAPPA, o.kappa),
                       (_lib.DV, o.dV)):
        s.put(which, arr[None])
    s.put(_lib.COST, np.array([r0.L]))
    s.run_phase(_lib.PHASE_LINESEARCH)
    eps_o, x_o, u_o, L_o, ls_o = o.linesearch(r0.L)
    assert s.get(_lib.EPS)[0] == eps_o and s.get_int(_lib.I_LS_ITERS)[0] == ls_o
    assert relerr(s.get(_lib.X_BAR)[0], x_o) < 1e-10 and relerr(s.get(_lib.U_BAR)[0], u_o) < 1e-10
    # every candidate of the last round against the oracle's rollout at the same eps
    table = ilqr_port.eps_table(prob.beta)
    base = ((ls_o - 1) // s.A) * s.A
    Lc, Ec = s.get(_lib.CAND_COST)[0], s.get(_lib.CAND_EXPECTED)[0]
    for ai in range(min(s.A, len(table) - base)):
        _, _, L_ref, E_ref = o.rollout(table[base + ai])
        assert abs(Lc[ai] - L_ref) <= 1e-10 * abs(L_ref)
        assert abs(Ec[ai] - E_ref) <= 1e-10 * max(1e-300, abs(E_ref))


@pytest.mark.parametrize("kp", [derivs_interpolation("setInterval", 1, 0, 0, 0),
                                derivs_interpolation("setInterval", 7, 0, 0, 0),
                                derivs_interpolation("adaptiveJerk", 3, 9, 2e-3, 0),
                                derivs_interpolation("iterativeError", 3, 0, 0, 1e-6)],
                         ids=["every", "interval7", "jerk", "iterError"])
def test_derivatives_teacher_forced(kp):
    """_get_derivatives on a given (x, u): keypoints, exact Jacobians at keypoints, lerp between."""
    rng = np.random.default_rng(11)
    prob = problems.cart_pole_with_wall(60)
    N, T = prob.N, prob.N - 1
    x = prob.x0[None] + np.cumsum(0.05 * rng.standard_normal((N, 4)), axis=0)
    u = 5.0 * rng.standard_normal((T, 1))
    s, o = make_gpu(prob, kp=kp), make_oracle(prob, kp=kp)
    s.put(_lib.X_BAR, x[None]); s.put(_lib.U_BAR, u[None])
    s.run_phase(_lib.PHASE_DERIVATIVES)
    kps = o.get_derivatives(x, u)
    assert s.keypoints()[0] == kps
    assert relerr(s.get(_lib.FX)[0], o.fx) < 1e-10 and relerr(s.get(_lib.FU)[0], o.fu) < 1e-10


def test_batch_equals_independent_solves():
    """Trajectories are independent: a batch of 6 seeds equals 6 single solves bit for bit, for
    any number of line-search candidates evaluated per round."""
    prob = problems.acrobot(40)
    x0 = prob.batch_x0(6, seed=0)
    big = make_gpu(prob, B=6, A=2, x0=x0)
    big.solve(max_iters=6)
    for b in range(6):
        one = make_gpu(prob, B=1, A=5, x0=x0[b])
        one.solve(max_iters=6)
        assert one.cost[0] == big.cost[b]
        np.testing.assert_array_equal(one.get(_lib.K)[0], big.get(_lib.K)[b])
        assert one.get_int(_lib.I_ITERS)[0] == big.get_int(_lib.I_ITERS)[b]
        o = make_oracle(prob, x0=x0[b])
        o.solve(max_iters=6)
        assert abs(o.trace[-1].L - big.cost[b]) <= 1e-8 * abs(big.cost[b])
        assert len(o.trace) == big.get_int(_lib.I_ITERS)[b]


def test_dropin_class_solve_and_save(tmp_path):
    """The reference call sequence of pendulum.py:85-100 on the drop-in class."""
    from ilqr import IterativeLinearQuadraticRegulator
    prob = problems.pendulum(100)
    ilqr = IterativeLinearQuadraticRegulator(prob.system, prob.N)
    ilqr.SetInitialState(prob.x0)
    ilqr.SetTargetState(prob.x_nom)
    ilqr.SetRunningCost(prob.Q, prob.R)
    ilqr.SetTerminalCost(prob.Qf)
    ilqr.SetInitialGuess(prob.u_guess)
    with contextlib.redirect_stdout(io.StringIO()) as f:
        states, inputs, solve_time, optimal_cost = ilqr.Solve()
    o = make_oracle(prob)
    xo, uo, Lo = o.solve()
    assert states.shape == (2, 100) and inputs.shape == (1, 99)
    assert len(f.getvalue().strip().split("\n")) - 3 == len(o.trace)        # same iteration count
    assert abs(optimal_cost - Lo) <= 1e-5 * abs(Lo)                          # north-star tolerance
    assert relerr(states, xo.T) < 1e-7
    assert relerr(ilqr.K, o.K.transpose(1, 2, 0)) < 1e-4 and ilqr.K.shape == (1, 2, 99)
    assert ilqr.fx.shape == (2, 2, 99) and ilqr.fu.shape == (2, 1, 99) and ilqr.kappa.shape == (1, 99)
    fname = str(tmp_path / "sol.npz")
    ilqr.SaveSolution(fname)                                                 # ilqr.py:712-733
    d = np.load(fname)
    assert set(d.files) == {"t", "x_bar", "u_bar", "K"}
    assert d["x_bar"].shape == (2, 99) and d["t"].shape == (99,) and d["K"].shape == (1, 2, 99)


def test_mpc_resolve_keeps_stale_state():
    """Receding-horizon resolves on the SAME object (acrobot.py:142-155): K, kappa, x_bar of the
    previous solve feed the first rollout of the next one."""
    from ilqr import IterativeLinearQuadraticRegulator
    prob = problems.acrobot(40)
    ilqr = IterativeLinearQuadraticRegulator(prob.system, prob.N, beta=prob.beta)
    o = make_oracle(prob)
    ilqr.SetTargetState(prob.x_nom); ilqr.SetRunningCost(prob.Q, prob.R); ilqr.SetTerminalCost(prob.Qf)
    x0, u_guess, replan = prob.x0, prob.u_guess, 2
    for resolve in range(3):
        ilqr.SetInitialState(x0); ilqr.SetInitialGuess(u_guess)
        with contextlib.redirect_stdout(io.StringIO()):
            x, u, _, L = ilqr.Solve()
        o.set_initial_state(x0); o.set_initial_guess(u_guess)
        xo, uo, Lo = o.solve()
        assert abs(L - Lo) <= 1e-7 * abs(Lo), resolve
        assert relerr(x, xo.T) < 1e-6
        u_guess = np.block([u[:, replan:], np.repeat(u[:, -1][np.newaxis].T, replan, axis=1)])
        x0 = x[:, replan]


def test_linesearch_failure_raises_like_reference():
    """At an optimum no eps improves the cost: RuntimeError('linesearch failed after 27 iterations')."""
    from ilqr import IterativeLinearQuadraticRegulator
    prob = problems.affine_sin(4, 1, 40)
    prob.delta = -1.0          # never stop on improvement -> must end in the line-search failure
    ilqr = IterativeLinearQuadraticRegulator(prob.system, prob.N, beta=0.5, delta=prob.delta)
    ilqr.SetInitialState(prob.x0); ilqr.SetTargetState(prob.x_nom)
    ilqr.SetRunningCost(prob.Q, prob.R); ilqr.SetTerminalCost(prob.Qf); ilqr.SetInitialGuess(prob.u_guess)
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(RuntimeError, match="linesearch failed after 27 iterations"):
            ilqr.Solve()
    o = make_oracle(prob)
    with pytest.raises(RuntimeError, match="linesearch failed after 27 iterations"):
        o.solve()


def test_infeasible_rollout_is_infinite_cost():
    """Non-finite dynamics => L = inf for that candidate (ilqr.py:317-323); on the first iteration
    inf - inf = nan is rejected for every eps, so the solve fails exactly like the reference."""
    prob = problems.pendulum(20)
    prob.u_guess = np.full((1, 19), 1e308)
    s = make_gpu(prob, A=8)
    s.begin_solve()
    s.iterate()
    assert s.status[0] == _lib.TRAJ_LINESEARCH_FAILED
    assert np.all(np.isinf(s.get(_lib.CAND_COST)[0][:1]))
    o = make_oracle(prob)
    with pytest.raises(RuntimeError, match="linesearch failed after 360 iterations"):
        o.solve()


def test_full_size_c4_properties():
    """BASELINE config C4 at full size (n=36, m=12, N=200, B=1024): size-independent properties
    (duplicate seeds bit-identical, accepted steps decrease the cost) plus spot checks of three
    trajectories against the oracle for three iterations, both end to end and teacher-forced
    (the oracle's backward pass run on the very fx, fu, x_bar, u_bar the GPU produced)."""
    prob = problems.quadruped(200)
    B = 1024
    x0 = prob.batch_x0(B, seed=0)
    x0[-4:] = x0[:4]
    s = make_gpu(prob, B=B, A=2, x0=x0)
    s.begin_solve()
    prev = np.full(B, np.inf)
    prev_iters = np.zeros(B, dtype=np.int32)
    spot = [0, 1, 517]
    oracles = [make_oracle(prob, x0=x0[b]) for b in spot]
    Ls = [np.inf] * len(spot)
    for it in range(3):
        s.iterate()
        cost, status, iters = s.cost, s.status, s.get_int(_lib.I_ITERS)
        ok = status != _lib.TRAJ_LINESEARCH_FAILED
        ran = iters > prev_iters                                  # still iterating this round
        assert ok.mean() > 0.99
        assert np.all(cost[ok & ran] < prev[ok & ran])            # accepted steps decrease the cost
        np.testing.assert_array_equal(cost[~ran], prev[~ran])     # converged ones are frozen
        prev_iters = iters
        np.testing.assert_array_equal(cost[-4:], cost[:4])       # duplicate seeds: bit-identical
        K = s.get(_lib.K)
        np.testing.assert_array_equal(K[-4:], K[:4])
        fx, fu, xb, ub = s.get(_lib.FX), s.get(_lib.FU), s.get(_lib.X_BAR), s.get(_lib.U_BAR)
        kappa, dV = s.get(_lib.KAPPA), s.get(_lib.DV)
        for k, b in enumerate(spot):
            if not ran[b]:
                continue
            rec = oracles[k].iterate(Ls[k])
            Ls[k] = rec.L
            assert abs(cost[b] - rec.L) <= 1e-7 * abs(rec.L)      # north star: 1e-5
            assert s.get_int(_lib.I_LS_ITERS)[b] == rec.ls_iters
            assert relerr(K[b], oracles[k].K) < 1e-4              # north star: 1e-4 on K, k
            # teacher-forced backward pass at full size
            o = make_oracle(prob, x0=x0[b])
            o.fx, o.fu, o.x_bar, o.u_bar = fx[b].copy(), fu[b].copy(), xb[b].copy(), ub[b].copy()
            o.backward_pass()
            assert relerr(K[b], o.K) < 1e-6
            assert np.abs(kappa[b] - o.kappa).max() < 1e-6 * max(1.0, np.abs(o.kappa).max())
            assert np.abs(dV[b] - o.dV).max() < 1e-6 * max(1.0, np.abs(o.dV).max())
            for t in (0, 97, 198):   # Jacobians of the accepted rollout against the host AD
                fxo, fuo = o.dyn.jac(xb[b, t], ub[b, t])
                assert np.abs(fx[b, t] - fxo).max() < 1e-9 * max(1.0, np.abs(fxo).max())
                assert np.abs(fu[b, t] - fuo).max() < 1e-9 * max(1.0, np.abs(fuo).max())
        prev = cost


def test_quadruped_fused_linearization_matches_ad_kernel(monkeypatch):
    """The fused quadruped linearization (dual leg evaluation along the 16 local directions + DMMA
    chain, csrc/quadruped_fused.cuh, the default) against the generic forward-mode-AD kernel
    (DDP_QUAD_LINEARIZE=ad) on the same points, in contact and in flight, and against the host AD."""
    prob = problems.quadruped(60)
    B = 16
    rng = np.random.default_rng(5)
    x = prob.x0[None, None] + 0.05 * rng.standard_normal((B, prob.N, 36))
    x[B // 2:, :, 2] += 0.05                     # second half airborne
    u = prob.u_guess.T[None] + 2.0 * rng.standard_normal((B, prob.N - 1, 12))
    out = {}
    for mode in ("ad", "fused"):
        monkeypatch.setenv("DDP_QUAD_LINEARIZE", mode)
        s = make_gpu(prob, B=B)
        s.put(_lib.X_BAR, x)
        s.put(_lib.U_BAR, u)
        s.run_phase(_lib.PHASE_DERIVATIVES)
        out[mode] = (s.get(_lib.FX), s.get(_lib.FU))
    o = make_oracle(prob)
    assert relerr(out["fused"][0], out["ad"][0]) < 1e-12
    assert relerr(out["fused"][1], out["ad"][1]) < 1e-12
    for b in (0, B - 1):
        for t in (0, 17, 58):
            fxo, fuo = o.dyn.jac(x[b, t], u[b, t])
            assert np.abs(out["fused"][0][b, t] - fxo).max() < 1e-10 * max(1.0, np.abs(fxo).max())
            assert np.abs(out["fused"][1][b, t] - fuo).max() < 1e-10 * max(1.0, np.abs(fuo).max())


def test_quadruped_quat_fused_linearization_matches_ad_kernel(monkeypatch):
    """The fused linearization of the reference's n = 37 quaternion layout (leg derivatives along 16
    intermediate directions, mapped to state coordinates through the quaternion / body-frame
    relations, DMMA chain; csrc/quadruped_quat_fused.cuh, the default) against the generic
    forward-mode-AD kernel (DDP_QUAD_LINEARIZE=ad) and the host AD, on points in contact and in
    flight with tilted, NON-unit quaternions and nonzero angular velocity."""
    prob = problems.quadruped_quat(60)
    B = 16
    rng = np.random.default_rng(7)
    x = prob.x0[None, None] + 0.05 * rng.standard_normal((B, prob.N, 37))
    x[:, :, 0:4] += 0.1 * rng.standard_normal((B, prob.N, 4))      # tilted, |q| != 1
    x[:, :, 19:22] += 0.5 * rng.standard_normal((B, prob.N, 3))     # spinning
    x[B // 2:, :, 6] += 0.05                                          # second half airborne
    u = prob.u_guess.T[None] + 2.0 * rng.standard_normal((B, prob.N - 1, 12))
    out = {}
    for mode in ("ad", "fused"):
        monkeypatch.setenv("DDP_QUAD_LINEARIZE", mode)
        s = make_gpu(prob, B=B)
        s.put(_lib.X_BAR, x)
        s.put(_lib.U_BAR, u)
        s.run_phase(_lib.PHASE_DERIVATIVES)
        out[mode] = (s.get(_lib.FX), s.get(_lib.FU))
    o = make_oracle(prob)
    assert relerr(out["fused"][0], out["ad"][0]) < 1e-11
    assert relerr(out["fused"][1], out["ad"][1]) < 1e-11
    for b in (0, B - 1):
        for t in (0, 17, 58):
            fxo, fuo = o.dyn.jac(x[b, t], u[b, t])
            assert np.abs(out["fused"][0][b, t] - fxo).max() < 1e-10 * max(1.0, np.abs(fxo).max())
            assert np.abs(out["fused"][1][b, t] - fuo).max() < 1e-10 * max(1.0, np.abs(fuo).max())


def test_backward_newton_schulz_inverse_matches_gauss_jordan(monkeypatch):
    """backward_mma_kernel inverts Quu by Newton-Schulz on the tensor pipe, seeded with the inverse
    of the previous step, and falls back to Gauss-Jordan with partial pivoting when the seed does
    not contract (csrc/backward_mma.cuh).  Both must give the gains of ilqr.py:655-660: compare the
    default against DDP_BWD_INVERSE=gauss-jordan on the same linearization, smooth steps and
    contact switches (half of the batch lands during the horizon) included."""
    prob = problems.quadruped(80)
    B = 8
    rng = np.random.default_rng(11)
    x0 = prob.batch_x0(B, seed=3)
    x0[B // 2:, 2] += 0.03                        # dropped from 3 cm: touches down inside the horizon
    out = {}
    for mode in ("newton", "gauss-jordan"):
        monkeypatch.delenv("DDP_BWD_INVERSE", raising=False)
        if mode == "gauss-jordan":
            monkeypatch.setenv("DDP_BWD_INVERSE", "gauss-jordan")
        s = make_gpu(prob, B=B, x0=x0)
        s.begin_solve()
        s.iterate()
        out[mode] = (s.get(_lib.K), s.get(_lib.KAPPA), s.get(_lib.DV), s.cost.copy())
    for a, b_ in zip(out["newton"], out["gauss-jordan"]):
        assert relerr(a, b_) < 1e-10


@pytest.mark.parametrize("name", ["quadruped", "quadruped_quat"])
def test_quadruped_rollout8_matches_generic_rollout(monkeypatch, name):
    """The 8-lane quadruped rollout (csrc/quadruped_rollout.cuh; QUAT = the reference's n = 37
    layout) against the generic rollout kernel: same candidates, same accepted step, states equal
    up to the association of the feedback sum."""
    prob = getattr(problems, name)(60)
    B = 6
    x0 = prob.batch_x0(B, seed=2)
    out = {}
    for mode in ("quad8", "generic"):
        monkeypatch.delenv("DDP_QUAD_ROLLOUT", raising=False)
        if mode == "generic":
            monkeypatch.setenv("DDP_QUAD_ROLLOUT", "generic")
        for A in (8, 4):      # 8: one CTA per trajectory with staged operands; 4: the unstaged variant
            s = make_gpu(prob, B=B, x0=x0, A=A)
            s.begin_solve()
            for _ in range(3):
                s.iterate()
            out[mode, A] = (s.get(_lib.X_BAR), s.get(_lib.U_BAR), s.cost.copy(), s.get_int(_lib.I_LS_ITERS).copy(),
                            s.get(_lib.CAND_COST).copy())
    for A in (8, 4):
        q, g = out["quad8", A], out["generic", A]
        assert np.array_equal(q[3], g[3])
        assert relerr(q[0], g[0]) < 1e-10
        assert relerr(q[1], g[1]) < 1e-9
        assert relerr(q[2], g[2]) < 1e-11
        fin = np.isfinite(g[4])
        assert np.array_equal(fin, np.isfinite(q[4]))
        assert relerr(q[4][fin], g[4][fin]) < 1e-10


def test_arm_rollout8_matches_generic_rollout(monkeypatch):
    """The 8-lane arm + ball rollout (csrc/arm_rollout.cuh, config C5) against the generic rollout
    kernel: same candidates, same accepted step, states equal up to the association of the tool-tip
    velocity / feedback / cost sums.  A non-diagonal cost takes the generic kernel either way."""
    prob = problems.arm_ball(120)
    B = 5
    x0 = prob.batch_x0(B, seed=3)
    out = {}
    for mode in ("arm8", "generic"):
        monkeypatch.delenv("DDP_ARM_ROLLOUT", raising=False)
        if mode == "generic":
            monkeypatch.setenv("DDP_ARM_ROLLOUT", "generic")
        s = make_gpu(prob, B=B, x0=x0, A=6)
        s.begin_solve()
        for _ in range(3):
            s.iterate()
        out[mode] = (s.get(_lib.X_BAR), s.get(_lib.U_BAR), s.cost.copy(), s.get_int(_lib.I_LS_ITERS).copy(),
                     s.get(_lib.CAND_COST).copy(), s.get(_lib.CAND_X).copy())
    q, g = out["arm8"], out["generic"]
    assert np.array_equal(q[3], g[3])
    assert relerr(q[0], g[0]) < 1e-10
    assert relerr(q[1], g[1]) < 1e-9
    assert relerr(q[2], g[2]) < 1e-11
    fin = np.isfinite(g[4])
    assert np.array_equal(fin, np.isfinite(q[4]))
    assert relerr(q[4][fin], g[4][fin]) < 1e-10
    assert relerr(q[5], g[5]) < 1e-9


@pytest.mark.parametrize("name", ["pendulum", "acrobot", "cart_pole_with_wall"])
def test_small_backward_register_kernel_equals_cta_kernel(monkeypatch, name):
    """backward_small_kernel (n <= 4, m = 1: a thread per trajectory, matrices in registers) against
    the CTA kernel (DDP_SMALL_BACKWARD=cta): same formulas in the same order, so K, kappa, dV and the
    costs of the following iterations are bit-identical."""
    prob = getattr(problems, name)(60)
    B = 5                                    # the register kernel is used up to B = 8
    x0 = prob.batch_x0(B, seed=4)
    out = {}
    for mode in ("reg", "cta"):
        monkeypatch.delenv("DDP_SMALL_BACKWARD", raising=False)
        if mode == "cta":
            monkeypatch.setenv("DDP_SMALL_BACKWARD", "cta")
        s = make_gpu(prob, B=B, x0=x0)
        s.begin_solve()
        for _ in range(4):
            s.iterate()
        out[mode] = (s.get(_lib.K), s.get(_lib.KAPPA), s.get(_lib.DV), s.cost.copy())
    for a, b in zip(out["reg"], out["cta"]):
        assert np.array_equal(a, b)


def test_split_iterate_and_host_exchange_equal_iterate():
    """ddp_iterate_linesearch / _finish_async / _wait with the overlapped host exchange
    (HostExchange: staged upload, control read-back under the backward pass) give bit-identical
    results to ddp_iterate with blocking set_initial_* / get."""
    import torch
    prob = problems.quadruped(40)
    B = 5
    x0 = prob.batch_x0(B, seed=4)
    u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, prob.N - 1, 12)))
    ref = make_gpu(prob, B=B, x0=x0)
    ref.begin_solve()
    for _ in range(3):
        ref.iterate()
    s = make_gpu(prob, B=B, x0=x0)
    x0_pin = torch.from_numpy(x0.copy()).pin_memory()
    u_pin = torch.from_numpy(u0.copy()).pin_memory()
    s.begin_solve()
    ex = s.host_exchange()
    ex.stage_inputs(x0_pin, u_pin)
    for _ in range(3):
        ex.apply_inputs()
        s.iterate_linesearch()
        ex.read_controls(u_pin)
        s.iterate_finish_async()
        ex.wait_controls()
        ex.stage_inputs(x0_pin, u_pin)
        n_active = s.iterate_wait()
    assert n_active == int((ref.get_int(_lib.I_STATUS) == 0).sum())
    for which in (_lib.X_BAR, _lib.U_BAR, _lib.K, _lib.KAPPA, _lib.COST):
        assert np.array_equal(s.get(which), ref.get(which))
    assert np.array_equal(u_pin.numpy(), ref.get(_lib.U_BAR))


def test_host_exchange_with_device_rearm_equals_device_resident_loop():
    """The end-to-end loop of bench.py (host buffers uploaded and read back every iteration) with the
    device-side MPC re-arm on: a trajectory the device re-arms keeps its shifted tape over the stale
    uploaded row (ddp_apply_staged_inputs), so the loop is bit-identical to the device-resident one."""
    import torch
    prob = problems.acrobot(40)
    B, iters = 6, 24
    x0 = prob.batch_x0(B, seed=5)
    u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, prob.N - 1, 1)))
    ref = make_gpu(prob, B=B, x0=x0, A=4)
    ref.set_mpc_rearm(2)
    ref.begin_solve()
    for _ in range(iters):
        ref.iterate()
    assert ref.get_int(_lib.I_RESOLVES).sum() >= B
    s = make_gpu(prob, B=B, x0=x0, A=4)
    s.set_mpc_rearm(2)
    x0_pin = torch.from_numpy(x0.copy()).pin_memory()
    u_pin = torch.from_numpy(u0.copy()).pin_memory()
    cost_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    s.begin_solve()
    ex = s.host_exchange()
    ex.stage_inputs(x0_pin, u_pin)
    for _ in range(iters):
        ex.apply_inputs()
        s.iterate_linesearch()
        ex.read_controls(u_pin)
        s.iterate_finish_async()
        ex.wait_controls()
        ex.stage_inputs(x0_pin, u_pin)
        s.iterate_wait()
        ex.read_state(x0_pin, cost_pin)
    np.testing.assert_array_equal(s.get_int(_lib.I_RESOLVES), ref.get_int(_lib.I_RESOLVES))
    for which in (_lib.X_BAR, _lib.U_BAR, _lib.K, _lib.COST, _lib.X0, _lib.CONVERGED_COST):
        assert np.array_equal(s.get(which), ref.get(which)), which
    assert np.array_equal(cost_pin.numpy(), ref.cost) and np.array_equal(x0_pin.numpy(), ref.get(_lib.X0))


def test_control_limits_extension_matches_oracle(monkeypatch):
    """ddp_set_control_limits (SetControlLimits is `pass` in the reference): clamped
    rollouts against the oracle port with the same box, through the 8-lane quadruped rollout and the
    generic rollout kernel; switched off again it is bit-identical to a solver that never had limits.
    (Plain clamping can stall the line search -- then both sides report the reference's
    "linesearch failed" at the same iteration.)"""
    lim = 3.2                                     # |u_stand| reaches 3.06: the box binds as soon as the gait starts
    for mode in ("quad8", "generic"):
        monkeypatch.delenv("DDP_QUAD_ROLLOUT", raising=False)
        if mode == "generic":
            monkeypatch.setenv("DDP_QUAD_ROLLOUT", "generic")
        prob = problems.quadruped(40)
        s, o = make_gpu(prob), make_oracle(prob)
        s.set_control_limits(-lim, lim)
        o.u_lim = (-lim * np.ones(prob.system.m), lim * np.ones(prob.system.m))
        s.begin_solve()
        L = np.inf
        for _ in range(4):
            s.iterate()
            rec = o.iterate(L)
            L = rec.L
            assert abs(s.cost[0] - rec.L) <= COST_RTOL * abs(rec.L)
            assert int(s.get_int(_lib.I_LS_ITERS)[0]) == rec.ls_iters
            u = s.get(_lib.U_BAR)[0]
            assert u.max() <= lim and u.min() >= -lim
            assert np.abs(u - o.u_bar).max() < 1e-8 * max(1.0, np.abs(o.u_bar).max())
        assert (np.abs(u) == lim).any()                               # the box was active
    monkeypatch.delenv("DDP_QUAD_ROLLOUT", raising=False)
    # a clamped pendulum swing-up stalls: same "linesearch failed" on both sides, same iteration
    prob = problems.pendulum(60)
    s, o = make_gpu(prob), make_oracle(prob)
    s.set_control_limits(-2.0, 2.0)
    o.u_lim = (-2.0 * np.ones(1), 2.0 * np.ones(1))
    with pytest.raises(RuntimeError, match="linesearch failed"):
        o.solve(max_iters=10)
    s.solve(max_iters=10)
    assert s.status[0] == _lib.TRAJ_LINESEARCH_FAILED and s.get_int(_lib.I_ITERS)[0] in (len(o.trace), len(o.trace) + 1)
    assert abs(s.cost[0] - o.trace[-1].L) <= COST_RTOL * abs(o.trace[-1].L)
    prob = problems.quadruped(40)
    s2, ref = make_gpu(prob), make_gpu(prob)
    s2.set_control_limits(-lim, lim)
    s2.set_control_limits(None, None)
    for z in (s2, ref):
        z.begin_solve()
        for _ in range(3):
            z.iterate()
    assert np.array_equal(s2.get(_lib.U_BAR), ref.get(_lib.U_BAR)) and np.array_equal(s2.cost, ref.cost)


def test_compute_sanitizer_clean_on_a_whole_iteration():
    """compute-sanitizer memcheck / synccheck / racecheck on two full iLQR iterations (rollout,
    fused linearization, symmetric backward sweep) of small quadruped, quadruped_quat and pendulum
    problems: 0 errors.  memcheck runs on the product library; synccheck and racecheck on the
    test-only build -DDDP_SANITIZER_BUILD of the same sources (csrc/backward_sym.cuh explains:
    synccheck rejects an aligned named barrier reached from two role functions, racecheck does not
    model mbarrier ordering; scratch/ub/sync_named.cu and race_mbar.cu reproduce both)."""
    import shutil
    import subprocess
    import sys
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tool in ("memcheck", "synccheck", "racecheck"):
        env = dict(os.environ)
        if tool != "memcheck":
            env["DDP_B200_LIB"] = _lib.build_racecheck()      # built by build(); compiled here if it did not travel
        r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "7", sys.executable,
                            os.path.join(root, "tests", "sanitize_small.py")],
                           capture_output=True, text=True, timeout=900, cwd=root, env=env)
        tail = (r.stdout + r.stderr)[-1500:]
        assert r.returncode == 0, (tool, tail)
        assert "ERROR SUMMARY: 0 errors" in r.stdout + r.stderr or "RACECHECK SUMMARY: 0 hazards" in r.stdout + r.stderr, (tool, tail)


def test_quu_regularization_extension_matches_oracle():
    """ddp_set_regularization (SURVEY 8f-4 extension, default 0 = reference): Quu + mu*I in the
    backward pass, against the oracle port with the same mu; mu = 0 stays bit-identical to the
    default path."""
    prob = problems.quadruped(40)
    for mu in (0.0, 1e-2):
        s = make_gpu(prob)
        s.set_regularization(mu)
        o = make_oracle(prob)
        o.quu_reg = mu
        s.begin_solve()
        L = np.inf
        for _ in range(3):
            s.iterate()
            rec = o.iterate(L)
            L = rec.L
            assert abs(s.cost[0] - rec.L) <= COST_RTOL * abs(rec.L)
            assert int(s.get_int(_lib.I_LS_ITERS)[0]) == rec.ls_iters
            assert relerr(s.get(_lib.K)[0], o.K) < GAIN_RTOL
        if mu == 0.0:
            ref = make_gpu(prob)
            ref.begin_solve()
            for _ in range(3):
                ref.iterate()
            assert np.array_equal(ref.get(_lib.K), s.get(_lib.K))
    small = problems.acrobot(40)           # scalar backward kernel (n < 16)
    s = make_gpu(small)
    s.set_regularization(0.5)
    o = make_oracle(small)
    o.quu_reg = 0.5
    s.begin_solve()
    s.iterate()
    o.iterate(np.inf)
    assert relerr(s.get(_lib.K)[0], o.K) < GAIN_RTOL
