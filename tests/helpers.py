"""Shared builders for the parity tests: the same Problem goes to the CUDA solver (through
the C ABI) and to the CPU oracle."""
import numpy as np

from drake_ddp_b200 import _lib
from oracle.dynamics import HostDynamics
from oracle.ilqr_port import IlqrOracle


def make_gpu(prob, B=1, kp="problem", A=None, x0=None, u_guess=None):
    from drake_ddp_b200.ilqr import BatchedILQR
    if kp == "problem":
        kp = prob.keypoints
    s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta,
                    gamma=prob.gamma, derivs_keypoint_method=kp, ls_parallel=A)
    s.set_cost(prob.Q, prob.R, prob.Qf)
    s.set_target(prob.x_nom)
    s.set_initial_state(prob.x0 if x0 is None else x0)
    s.set_initial_guess(prob.u_guess if u_guess is None else u_guess)
    return s


def make_oracle(prob, kp="problem", x0=None, u_guess=None):
    if kp == "problem":
        kp = prob.keypoints
    o = IlqrOracle(HostDynamics(prob.system), prob.N, delta=prob.delta, beta=prob.beta,
                   gamma=prob.gamma, keypoints=kp)
    o.set_initial_state(prob.x0 if x0 is None else x0)
    o.set_target_state(prob.x_nom)
    o.set_running_cost(prob.Q, prob.R)
    o.set_terminal_cost(prob.Qf)
    o.set_initial_guess(prob.u_guess if u_guess is None else u_guess)
    return o


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def gpu_state(s, b=0):
    """(x_bar, u_bar, K, kappa, dV, fx, fu) of trajectory b in the oracle's time-major layout."""
    g = s.get
    return (g(_lib.X_BAR)[b], g(_lib.U_BAR)[b], g(_lib.K)[b], g(_lib.KAPPA)[b], g(_lib.DV)[b],
            g(_lib.FX)[b], g(_lib.FU)[b])


# ---- many oracle solves in parallel (converged-Solve parity tests) ------------------------------
def _oracle_solve_worker(args):
    """Solve one problem to convergence with the oracle; x0 scaled by each of ``scales``.
    Returns per scale: (costs per iteration, failed flag, K, kappa)."""
    import os
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    expr, x0, scales, want_gains = args
    from drake_ddp_b200 import problems  # noqa: F401  (used by eval)
    prob = eval(expr)
    out = []
    for sc in scales:
        o = make_oracle(prob, x0=x0 * sc)
        failed = False
        try:
            o.solve(max_iters=200)
        except RuntimeError:
            failed = True
        out.append(([r.L for r in o.trace], failed, o.K.copy() if want_gains else None,
                    o.kappa.copy() if want_gains else None))
    return out


def oracle_solve_many(expr, x0s, scales=(1.0,), want_gains=True, procs=None):
    """``expr`` is a problem factory expression such as "problems.quadruped(200)" (evaluated in
    the worker processes, which only run the CPU oracle)."""
    import multiprocessing as mp
    import os
    procs = procs or max(1, min(len(x0s), (os.cpu_count() or 2)))
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        return pool.map(_oracle_solve_worker, [(expr, x0s[b], tuple(scales), want_gains) for b in range(len(x0s))],
                        chunksize=1)
