"""TEST INFRASTRUCTURE -- CPU oracle.  Not part of the product path.

A NumPy fp64 restatement of the reference solver's algorithm (one trajectory, one thread,
one timestep at a time -- the reference's execution model), used only as the checker in
``tests/``, in ``__graft_entry__.smoke()`` and as ``bench.py``'s CPU baseline.  Each function
cites the lines of /root/reference/ilqr.py it follows.  It is pinned against the unmodified
reference (imported through ``oracle/pydrake_shim.py``) by ``tests/test_oracle_pin.py`` and
against the committed fixtures in ``tests/golden/`` (made by ``oracle/make_golden.py``).

Parity status: the *solver* arithmetic is pinned to the reference's own code.  The dynamics
(Drake MultibodyPlant, unpinned version, not installable offline) are replaced by this
repo's analytic models on both sides, so parity with Drake's arithmetic is UNPINNED.

Arrays are time-major here: x (N,n), u (T,m), fx (T,n,n), fu (T,n,m), K (T,m,n),
kappa (T,m), dV (T,) with T = N-1.  The reference stores time last (ilqr.py:70-83).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class KeypointCfg:
    """utils_derivs_interpolation.py:3-9"""
    keypoint_method: str = "setInterval"
    minN: int = 1
    maxN: int = 0
    jerk_threshold: float = 0.0
    iterative_error_threshold: float = 0.0


def eps_table(beta: float, floor: float = 1e-8):
    """Candidate step sizes in the order the reference tries them: eps=1, eps*=beta while
    eps >= 1e-8 (ilqr.py:300-302,335).  Repeated multiply, not beta**k."""
    out, eps = [], 1.0
    while eps >= floor:
        out.append(eps)
        eps *= beta
    return np.array(out)


def keypoints_set_interval(N: int, minN: int):
    """ilqr.py:417-432 -- every minN steps, last one *replaced* by N-2."""
    kp = np.arange(0, N - 1, minN).astype(int)
    if kp[-1] != N - 2:
        kp[-1] = N - 2
    return [int(k) for k in kp]


def jerk_profile(x: np.ndarray):
    """ilqr.py:470-486 -- second difference of rows dof..2*dof-1, dof = int(n/2)."""
    N, n = x.shape
    dof = int(n / 2)
    vel = x[:, dof:2 * dof]
    return (vel[2:N - 1] - vel[1:N - 2]) - (vel[1:N - 2] - vel[0:N - 3])


def keypoints_adaptive_jerk(x: np.ndarray, minN: int, maxN: int, thr: float):
    """ilqr.py:434-468 -- signed jerk test gated by minN, forced keypoint at maxN."""
    N = x.shape[0]
    jerk = jerk_profile(x)
    kp, counter = [0], 0
    for t in range(jerk.shape[0]):
        counter += 1
        if counter >= minN and bool(np.any(jerk[t] > thr)):
            kp.append(t)
            counter = 0
        if counter >= maxN:
            kp.append(t)
            counter = 0
    if kp[-1] != N - 2:
        kp[-1] = N - 2
    return kp


@dataclass
class IterRecord:
    L: float
    eps: float
    ls_iters: int
    improvement: float
    keypoints: list = field(default_factory=list)


class IlqrOracle:
    def __init__(self, dyn, num_timesteps, delta=1e-2, beta=0.95, gamma=0.0, keypoints=None, quu_reg=0.0):
        """ilqr.py:21-100.  ``dyn`` has n, m, step(x,u), jac(x,u)."""
        self.dyn = dyn
        self.N, self.n, self.m = int(num_timesteps), dyn.n, dyn.m
        self.delta, self.beta, self.gamma = delta, beta, gamma
        self.quu_reg = float(quu_reg)   # extension (not in the reference): Quu + quu_reg*I; 0 = ilqr.py:654
        self.u_lim = None               # extension: (u_min, u_max) box the rollout clamps to; None = reference
        T, n, m = self.N - 1, self.n, self.m
        self.x0 = np.zeros(n)
        self.x_nom = np.zeros(n)
        self.Q, self.R, self.Qf = np.eye(n), np.eye(m), np.eye(n)
        self.x_bar, self.u_bar = np.zeros((self.N, n)), np.zeros((T, m))
        self.fx, self.fu = np.zeros((T, n, n)), np.zeros((T, n, m))
        self.kappa, self.K, self.dV = np.zeros((T, m)), np.zeros((T, m, n)), np.zeros(T)
        self.kp = keypoints if keypoints is not None else KeypointCfg("setInterval", 1, 0, 0, 0)
        self.percentage_derivs = 0.0
        self.trace = []

    # ---- setters (ilqr.py:102-156) ------------------------------------------------
    def set_initial_state(self, x0):
        self.x0 = np.asarray(x0, dtype=np.float64).reshape(self.n).copy()

    def set_target_state(self, x_nom):
        self.x_nom = np.asarray(x_nom, dtype=np.float64).reshape(self.n).copy()

    def set_running_cost(self, Q, R):
        assert Q.shape == (self.n, self.n) and R.shape == (self.m, self.m)
        self.Q, self.R = np.array(Q, dtype=np.float64), np.array(R, dtype=np.float64)

    def set_terminal_cost(self, Qf):
        assert Qf.shape == (self.n, self.n)
        self.Qf = np.array(Qf, dtype=np.float64)

    def set_initial_guess(self, u_guess_mT):
        """Takes the reference's (m, N-1) layout."""
        assert u_guess_mT.shape == (self.m, self.N - 1)
        self.u_bar = np.ascontiguousarray(np.asarray(u_guess_mT, dtype=np.float64).T)

    # ---- forward line search (ilqr.py:274-337) ---------------------------------------
    def rollout(self, eps):
        """One closed-loop rollout and its cost, ilqr.py:306-327.  Returns x, u, L, expected."""
        N, n, m = self.N, self.n, self.m
        x, u = np.zeros((N, n)), np.zeros((N - 1, m))
        L, expected = 0.0, 0.0
        x[0] = self.x0
        for t in range(N - 1):
            u[t] = self.u_bar[t] - eps * self.kappa[t] - self.K[t] @ (x[t] - self.x_bar[t])
            if self.u_lim is not None:
                u[t] = np.minimum(np.maximum(u[t], self.u_lim[0]), self.u_lim[1])
            xn = self.dyn.step(x[t], u[t])
            if not np.all(np.isfinite(xn)):      # Drake would throw: ilqr.py:317-323
                L = np.inf
                break
            x[t + 1] = xn
            e = x[t] - self.x_nom
            L += e.T @ self.Q @ e + u[t].T @ self.R @ u[t]
            expected += -eps * (1 - eps / 2) * self.dV[t]
        e = x[-1] - self.x_nom
        L += e.T @ self.Qf @ e
        return x, u, L, expected

    def linesearch(self, L_last):
        eps, n_ls = 1.0, 0
        while eps >= 1e-8:
            n_ls += 1
            x, u, L, expected = self.rollout(eps)
            if L_last - L > self.gamma * expected:      # first satisfying, ilqr.py:330-332
                return eps, x, u, L, n_ls
            eps *= self.beta
        raise RuntimeError("linesearch failed after %s iterations" % n_ls)

    # ---- derivatives (ilqr.py:380-621) ---------------------------------------------
    def _jac_at(self, x, u, t):
        self.fx[t], self.fu[t] = self.dyn.jac(x[t], u[t])

    def keypoints_iterative_error(self, x, u):
        """ilqr.py:488-593 -- breadth-first bisection; Jacobians memoised per index."""
        cfg, n = self.kp, self.n
        done = [False] * self.N
        pending = [(0, self.N - 2)]
        while pending:
            nxt = []
            for (s, e) in pending:
                if e - s <= cfg.minN:
                    continue
                mid = int((s + e) / 2)
                for t in (s, mid, e):
                    if not done[t]:
                        self._jac_at(x, u, t)
                        done[t] = True
                lin = (self.fx[e] + self.fx[s]) / 2
                err = float(np.sum((lin - self.fx[mid]) ** 2)) / (2 * n)
                if err > cfg.iterative_error_threshold:
                    nxt += [(s, mid), (mid, e)]
            pending = nxt
        return [t for t in range(self.N - 1) if done[t]]

    def interpolate(self, kp):
        """ilqr.py:596-621 -- lerp whole matrices between consecutive keypoints, j in [s, e)."""
        for s, e in zip(kp[:-1], kp[1:]):
            fxs, fxe = self.fx[s].copy(), self.fx[e]
            fus, fue = self.fu[s].copy(), self.fu[e]
            for j in range(s, e):
                self.fx[j] = fxs + (fxe - fxs) * (j - s) / (e - s)
                self.fu[j] = fus + (fue - fus) * (j - s) / (e - s)

    def get_derivatives(self, x, u):
        cfg = self.kp
        if cfg.keypoint_method == "setInterval":
            kp = keypoints_set_interval(self.N, cfg.minN)
        elif cfg.keypoint_method == "adaptiveJerk":
            kp = keypoints_adaptive_jerk(x, cfg.minN, cfg.maxN, cfg.jerk_threshold)
        elif cfg.keypoint_method == "iterativeError":
            kp = self.keypoints_iterative_error(x, u)
        else:
            raise Exception("unknown interpolation method")
        self.percentage_derivs = (len(kp) / (self.N - 1)) * 100
        if cfg.keypoint_method != "iterativeError":
            for t in kp:
                self._jac_at(x, u, t)
        if not (cfg.keypoint_method == "setInterval" and cfg.minN == 1):
            self.interpolate(kp)
        return kp

    # ---- backward Riccati sweep (ilqr.py:623-667, cost partials :161-206) -----------
    def backward_pass(self):
        Q, R, Qf, xn = self.Q, self.R, self.Qf, self.x_nom
        Vx = 2 * Qf @ self.x_bar[-1] - 2 * xn.T @ Qf
        Vxx = 2 * Qf
        for t in range(self.N - 2, -1, -1):
            x, u = self.x_bar[t], self.u_bar[t]
            lx = 2 * Q @ x - 2 * xn.T @ Q
            lu = 2 * R @ u
            fx, fu = self.fx[t], self.fu[t]
            Qx = lx + fx.T @ Vx
            Qu = lu + fu.T @ Vx
            Qxx = 2 * Q + fx.T @ Vxx @ fx
            Quu = 2 * R + fu.T @ Vxx @ fu
            if self.quu_reg:
                Quu = Quu + self.quu_reg * np.eye(self.m)
            Quu_inv = np.linalg.inv(Quu)
            Qux = fu.T @ Vxx @ fx
            self.kappa[t] = Quu_inv @ Qu
            self.K[t] = Quu_inv @ Qux
            self.dV[t] = Qu.T @ Quu_inv @ Qu
            Vx = Qx - Qu.T @ Quu_inv @ Qux
            Vxx = Qxx - Qux.T @ Quu_inv @ Qux

    # ---- outer loop (ilqr.py:669-710) ------------------------------------------------
    def iterate(self, L_last):
        eps, x, u, L, n_ls = self.linesearch(L_last)
        kp = self.get_derivatives(x, u)
        self.u_bar, self.x_bar = u, x
        self.backward_pass()
        rec = IterRecord(L=float(L), eps=float(eps), ls_iters=n_ls,
                         improvement=float(L_last - L), keypoints=list(kp))
        self.trace.append(rec)
        return rec

    def solve(self, max_iters=None):
        """Returns x_bar (N,n), u_bar (T,m), L.  ``max_iters`` caps the loop for fixed-count runs."""
        L, improvement, it = math.inf, math.inf, 0
        self.trace = []
        while improvement > self.delta and (max_iters is None or it < max_iters):
            rec = self.iterate(L)
            improvement, L = L - rec.L, rec.L
            it += 1
        return self.x_bar, self.u_bar, L
