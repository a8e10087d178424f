"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
solver (/root/reference/ilqr.py, imported through oracle/pydrake_shim.py) on this repo's
analytic models.  Only runs where /root/reference exists; the fixtures are committed so the
GPU box (which has no reference tree) can check both the oracle port and the CUDA path
against the reference's own outputs.

    python -m oracle.make_golden
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from drake_ddp_b200 import problems  # noqa: E402
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation  # noqa: E402
from oracle.pydrake_shim import ShimSystem, load_reference_ilqr  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (problem factory, keypoint cfg or None, iterations)
CASES = {
    "pendulum_N100": (lambda: problems.pendulum(100), None, 5),
    "acrobot_N40": (lambda: problems.acrobot(40), None, 6),
    "cart_pole_N60": (lambda: problems.cart_pole(60), None, 6),
    "wall_N60": (lambda: problems.cart_pole_with_wall(60), None, 5),
    "affine_4_1_setInterval5": (lambda: problems.affine_sin(4, 1, 40),
                                derivs_interpolation("setInterval", 5, 0, 0, 0), 3),
    "affine_4_1_adaptiveJerk": (lambda: problems.affine_sin(4, 1, 40),
                                derivs_interpolation("adaptiveJerk", 2, 10, 1e-4, 0), 3),
    "affine_4_1_iterativeError": (lambda: problems.affine_sin(4, 1, 40),
                                  derivs_interpolation("iterativeError", 2, 0, 0, 1e-9), 3),
    "affine_27_7_N30": (lambda: problems.affine_sin(27, 7, 30), None, 3),
    "arm_ball_N40": (lambda: problems.arm_ball(40, keypoints=None), None, 3),
    "arm_ball_N40_setInterval5": (lambda: problems.arm_ball(40),
                                  derivs_interpolation("setInterval", 5, 40, 1e-4, 1e-2), 2),
    "quadruped_N30": (lambda: problems.quadruped(30), None, 3),
    "quadruped_quat_N30": (lambda: problems.quadruped_quat(30), None, 3),
    "quadruped_N30_adaptiveJerk": (lambda: problems.quadruped(30),
                                   derivs_interpolation("adaptiveJerk", 2, 20, 0.3, 10), 3),
}


def run_reference(prob, kp, iters):
    ref, ref_utils = load_reference_ilqr()
    method = None if kp is None else ref_utils.derivs_interpolation(
        kp.keypoint_method, kp.minN, kp.maxN, kp.jerk_threshold, kp.iterative_error_threshold)
    r = ref.IterativeLinearQuadraticRegulator(ShimSystem(prob.system), prob.N, delta=prob.delta,
                                              beta=prob.beta, gamma=prob.gamma,
                                              derivs_keypoint_method=method)
    r.SetInitialState(prob.x0.copy())
    r.SetTargetState(prob.x_nom)
    r.SetRunningCost(prob.Q, prob.R)
    r.SetTerminalCost(prob.Qf)
    r.SetInitialGuess(prob.u_guess.copy())
    L, costs, epss, lss, pct = np.inf, [], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(iters):     # the body of Solve()'s loop, ilqr.py:695-697
            L, eps, ls = r._forward_pass(L)
            r._backward_pass()
            costs.append(L), epss.append(eps), lss.append(ls), pct.append(r.percentage_derivs)
    return dict(costs=np.array(costs), eps=np.array(epss), ls_iters=np.array(lss),
                percentage_derivs=np.array(pct), x_bar=r.x_bar, u_bar=r.u_bar, K=r.K,
                kappa=r.kappa, dV_coeff=r.dV_coeff, fx=r.fx, fu=r.fu)


# Full Solve() runs to convergence through the reference's own loop (ilqr.py:669-710), the
# north-star quantity "final cost of a converged Solve()": name -> (factory, keypoints, seeds of
# the batch_x0(., seed=0) rows solved).  Stored per trajectory: per-iteration costs / eps /
# ls_iters (recorded by wrapping the bound _forward_pass, the reference source is untouched),
# the returned final cost, x_bar, u_bar; K, kappa for the first trajectory only (size).
SOLVE_CASES = {
    "solve_quadruped_N200": (lambda: problems.quadruped(200), None, [0, 1, 2, 3, 4, 5, 6, 7]),
    "solve_quadruped_quat_N200": (lambda: problems.quadruped_quat(200), None, [0, 1, 2, 3]),
    "solve_arm_ball_N400_setInterval5": (lambda: problems.arm_ball(400), "problem", [0, 1, 2, 3]),
    "solve_pendulum_N100": (lambda: problems.pendulum(100), None, [0]),
}


def run_reference_solve(prob, kp, x0):
    ref, ref_utils = load_reference_ilqr()
    method = None if kp is None else ref_utils.derivs_interpolation(
        kp.keypoint_method, kp.minN, kp.maxN, kp.jerk_threshold, kp.iterative_error_threshold)
    r = ref.IterativeLinearQuadraticRegulator(ShimSystem(prob.system), prob.N, delta=prob.delta,
                                              beta=prob.beta, gamma=prob.gamma,
                                              derivs_keypoint_method=method)
    r.SetInitialState(x0.copy())
    r.SetTargetState(prob.x_nom)
    r.SetRunningCost(prob.Q, prob.R)
    r.SetTerminalCost(prob.Qf)
    r.SetInitialGuess(prob.u_guess.copy())
    rec = []
    inner = r._forward_pass

    def recording_forward_pass(L_last):
        out = inner(L_last)
        rec.append(out)
        return out

    r._forward_pass = recording_forward_pass
    failed = 0
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            x, u, _, L = r.Solve()
        except RuntimeError:            # "linesearch failed after %s iterations" (ilqr.py:337)
            failed, L = 1, rec[-1][0] if rec else np.inf
    return dict(final_cost=L, failed=failed, costs=np.array([o[0] for o in rec]),
                eps=np.array([o[1] for o in rec]), ls_iters=np.array([o[2] for o in rec]),
                x_bar=r.x_bar.copy(), u_bar=r.u_bar.copy(), K=r.K.copy(), kappa=r.kappa.copy())


def main_solve():
    for name, (factory, kp, rows) in SOLVE_CASES.items():
        prob = factory()
        kpc = prob.keypoints if kp == "problem" else kp
        x0s = prob.batch_x0(max(rows) + 1, seed=0) if prob.sigma > 0 else np.repeat(prob.x0[None], max(rows) + 1, 0)
        out = {"rows": np.array(rows)}
        for i, b in enumerate(rows):
            g = run_reference_solve(prob, kpc, x0s[b])
            for k, v in g.items():
                if k in ("K", "kappa", "x_bar") and i > 0:
                    continue
                out[f"{k}_{b}"] = v
            print(name, b, "final", g["final_cost"], "iters", len(g["costs"]), "failed", g["failed"])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    main_solve()
    for name, (factory, kp, iters) in CASES.items():
        prob = factory()
        out = run_reference(prob, kp, iters)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), iters=iters, **out)
        print(name, "costs", out["costs"])


if __name__ == "__main__":
    main()
