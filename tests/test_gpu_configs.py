"""BASELINE.json's configs C2, C3, C5 as parity cases (C1 and C4 are in test_gpu_parity.py), and
the device-side receding-horizon driver."""
import contextlib
import io

import numpy as np
import pytest

from drake_ddp_b200 import _lib, problems
from oracle import ilqr_port
from tests.helpers import make_gpu, make_oracle, relerr

pytestmark = pytest.mark.gpu


def test_c2_acrobot_mpc_batch_of_50_resolves():
    """C2: n=4, m=1, N=40, batch = the 50 (x0, u_guess) pairs of a receding-horizon run
    (replan_steps=2, acrobot.py:131-160) recorded from the CPU oracle, then solved as ONE batch of
    independent problems, each compared with a fresh oracle solve of the same pair."""
    prob = problems.acrobot(40)
    replan, n_pairs = 2, 50
    o = make_oracle(prob)
    x0, u_guess = prob.x0, prob.u_guess
    pairs = []
    for _ in range(n_pairs):
        pairs.append((np.array(x0), np.array(u_guess)))
        o.set_initial_state(x0)
        o.set_initial_guess(u_guess)
        try:
            x, u, _ = o.solve(max_iters=30)
        except RuntimeError:       # the reference would raise too; keep the last accepted iterate
            x, u = o.x_bar, o.u_bar
        u_guess = np.block([u.T[:, replan:], np.repeat(u.T[:, -1][np.newaxis].T, replan, axis=1)])
        x0 = x[replan]
    X0 = np.stack([p[0] for p in pairs])
    U0 = np.stack([p[1].T for p in pairs])
    s = make_gpu(prob, B=n_pairs, A=4, x0=X0, u_guess=U0)
    iters = 6
    s.solve(max_iters=iters)
    cost, ran, status = s.cost, s.get_int(_lib.I_ITERS), s.status
    for b in range(n_pairs):
        fresh = make_oracle(prob, x0=pairs[b][0], u_guess=pairs[b][1])
        try:
            fresh.solve(max_iters=iters)
            failed = False
        except RuntimeError:
            failed = True
        assert (status[b] == _lib.TRAJ_LINESEARCH_FAILED) == failed
        assert ran[b] == len(fresh.trace) + (1 if failed else 0) or ran[b] == len(fresh.trace)
        if fresh.trace:
            assert abs(cost[b] - fresh.trace[-1].L) <= 1e-6 * abs(fresh.trace[-1].L)   # north star 1e-5
            assert relerr(s.get(_lib.K)[b], fresh.K) < 1e-4


@pytest.mark.parametrize("beta,n_cand", [(0.95, 256), (0.5, 27)])
def test_c3_wall_256_linesearch_alphas_in_parallel(beta, n_cand):
    """C3: cart-pole with wall, n=4, m=1, N=200, the line-search step sizes beta^k evaluated in
    parallel for one trajectory (256 of the 360 candidates at beta=.95; all 27 at beta=.5).  Every
    candidate's cost and expected improvement is compared with the oracle's rollout at that eps,
    and the first-satisfying pick must agree."""
    prob = problems.cart_pole_with_wall(200, beta=beta)
    o = make_oracle(prob)
    L = np.inf
    for _ in range(2):                       # get non-trivial K, kappa, dV
        L = o.iterate(L).L
    o.kappa = o.kappa * 3.0                  # make eps = 1 overshoot so the pick is not trivial
    s = make_gpu(prob, A=n_cand)
    assert s.A == n_cand
    for which, arr in ((_lib.X_BAR, o.x_bar), (_lib.U_BAR, o.u_bar), (_lib.K, o.K), (_lib.KAPPA, o.kappa),
                       (_lib.DV, o.dV)):
        s.put(which, arr[None])
    s.put(_lib.COST, np.array([L]))
    s.run_phase(_lib.PHASE_LINESEARCH)
    table = ilqr_port.eps_table(beta)
    Lc, Ec = s.get(_lib.CAND_COST)[0], s.get(_lib.CAND_EXPECTED)[0]
    eps_o, x_o, u_o, L_o, ls_o = o.linesearch(L)
    if ls_o <= n_cand:                       # resolved in the first (parallel) round
        for k in range(n_cand):
            _, _, L_ref, E_ref = o.rollout(table[k])
            if np.isfinite(L_ref):
                # candidates that overshoot bounce off the wall chaotically (cost 1e5): their
                # rollouts amplify rounding, so only moderate-cost candidates get the tight bound
                tol = 1e-8 if L_ref <= 10 * L else 1e-3
                assert abs(Lc[k] - L_ref) <= tol * abs(L_ref), k
            else:
                assert not np.isfinite(Lc[k])
            assert abs(Ec[k] - E_ref) <= 1e-10 * max(1e-300, abs(E_ref))
    assert s.get(_lib.EPS)[0] == eps_o and s.get_int(_lib.I_LS_ITERS)[0] == ls_o
    assert relerr(s.get(_lib.X_BAR)[0], x_o) < 1e-9


def test_c5_arm_ball_interpolation_batch():
    """C5: n=27, m=7, N=400, derivative interpolation on (setInterval 5 -> 80 keypoints with the
    last one replaced by N-2; adaptiveJerk 5/40/1e-4), a batch of IC seeds; spot trajectories are
    compared with the oracle per iteration."""
    for kp_name in ("setInterval5", "adaptiveJerk"):
        prob = problems.arm_ball(400, keypoints=kp_name)
        B = 64
        x0 = prob.batch_x0(B, seed=0)
        s = make_gpu(prob, B=B, A=4, x0=x0)
        s.begin_solve()
        spot = [0, 37]
        oracles = [make_oracle(prob, x0=x0[b]) for b in spot]
        Ls = [np.inf, np.inf]
        for it in range(2):
            s.iterate()
            kps = s.keypoints()
            for k, b in enumerate(spot):
                rec = oracles[k].iterate(Ls[k])
                Ls[k] = rec.L
                assert kps[b] == rec.keypoints
                assert abs(s.cost[b] - rec.L) <= 1e-7 * abs(rec.L)
                assert s.get_int(_lib.I_LS_ITERS)[b] == rec.ls_iters
                assert relerr(s.get(_lib.FX)[b], oracles[k].fx) < 1e-7
                assert relerr(s.get(_lib.K)[b], oracles[k].K) < 1e-4
        if kp_name == "setInterval5":
            assert len(kps[0]) == 80 and kps[0][-3:] == [385, 390, 398]


def test_device_mpc_shift_matches_host_shift():
    """receding-horizon resolves with the control-tape shift done on the device
    (ddp_mpc_shift) equal the reference's host-side np.block shift (acrobot.py:145-153) followed
    by Solve() on the same object; batch of seeds, stale K/kappa/x_bar carried over."""
    prob = problems.acrobot(40)
    B, replan = 4, 2
    x0 = prob.batch_x0(B, seed=3)
    s = make_gpu(prob, B=B, A=4, x0=x0)
    oracles = [make_oracle(prob, x0=x0[b]) for b in range(B)]
    s.solve(max_iters=8)
    for o in oracles:
        o.solve(max_iters=8)
    for resolve in range(2):
        s.mpc_shift(replan)
        s.solve(max_iters=8)
        for b, o in enumerate(oracles):
            x, u = o.x_bar, o.u_bar
            u_guess = np.block([u.T[:, replan:], np.repeat(u.T[:, -1][np.newaxis].T, replan, axis=1)])
            o.set_initial_state(x[replan])
            o.set_initial_guess(u_guess)
            o.solve(max_iters=8)
            assert abs(s.cost[b] - o.trace[-1].L) <= 1e-6 * abs(o.trace[-1].L), (resolve, b)
            assert s.get_int(_lib.I_ITERS)[b] == len(o.trace)
            assert relerr(s.get(_lib.X_BAR)[b], o.x_bar) < 1e-5
