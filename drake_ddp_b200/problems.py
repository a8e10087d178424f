"""Problem data of the reference's example scripts and of BASELINE.json's five configs.

Every number is a module-level constant of one of the reference's scripts (cited per
builder) re-expressed for this repo's analytic models.  Costs follow
the scripts' convention: running cost is passed as ``dt*Q, dt*R`` (e.g. pendulum.py:93),
terminal cost unscaled.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import systems
from .utils_derivs_interpolation import derivs_interpolation


@dataclass
class Problem:
    name: str
    system: systems.AnalyticSystem
    N: int
    x0: np.ndarray
    x_nom: np.ndarray
    Q: np.ndarray          # already scaled by dt
    R: np.ndarray          # already scaled by dt
    Qf: np.ndarray
    u_guess: np.ndarray    # (m, N-1), reference layout
    beta: float = 0.95
    delta: float = 1e-2
    gamma: float = 0.0
    keypoints: Optional[derivs_interpolation] = None
    sigma: float = 0.0     # std-dev of the x0 perturbation used to seed a batch
    extra: dict = field(default_factory=dict)

    def batch_x0(self, B: int, seed: int = 0) -> np.ndarray:
        """(B, n) initial states: x0 + sigma * N(0, I), default_rng(seed)."""
        rng = np.random.default_rng(seed)
        return self.x0[None, :] + self.sigma * rng.standard_normal((B, self.system.n))


def pendulum(N: int = 100) -> Problem:
    """pendulum.py:18-34,85-98 (script horizon is 200; BASELINE config C1 uses N=100)."""
    sysm = systems.pendulum(dt=1e-2)
    dt = sysm.dt
    return Problem("pendulum", sysm, N, np.array([0.0, 0.0]), np.array([np.pi, 0.0]),
                   dt * 0.01 * np.diag([0.0, 1.0]), dt * 0.01 * np.eye(1), 100.0 * np.eye(2),
                   np.zeros((1, N - 1)), beta=0.95, delta=1e-2, gamma=0.0, sigma=0.1)


def acrobot(N: int = 40) -> Problem:
    """acrobot.py:19-45,112-125 (script horizon 750; config C2 uses N=40)."""
    sysm = systems.acrobot(dt=4e-3)
    dt = sysm.dt
    return Problem("acrobot", sysm, N, np.zeros(4), np.array([np.pi, 0.0, 0.0, 0.0]),
                   dt * 0.01 * np.diag([0.0, 0.0, 1.0, 1.0]), dt * 0.01 * np.eye(1),
                   100.0 * np.eye(4), np.zeros((1, N - 1)), beta=0.5, delta=1e-2, sigma=0.1)


def cart_pole(N: int = 200) -> Problem:
    """cart_pole.py:21-46,100-120."""
    sysm = systems.cart_pole(dt=1e-2)
    dt = sysm.dt
    return Problem("cart_pole", sysm, N, np.zeros(4), np.array([0.0, np.pi, 0.0, 0.0]),
                   dt * np.diag([10.0, 10.0, 0.1, 0.1]), dt * 0.001 * np.eye(1),
                   np.diag([100.0, 100.0, 10.0, 10.0]), np.zeros((1, N - 1)), beta=0.9,
                   delta=1e-2, sigma=0.05)


def cart_pole_with_wall(N: int = 200, beta: float = 0.5) -> Problem:
    """cart_pole_with_wall.py:22-52,142-160 (script horizon 100; config C3 uses N=200)."""
    sysm = systems.cart_pole_with_wall(dt=1e-2)
    dt = sysm.dt
    return Problem("cart_pole_with_wall", sysm, N, np.array([0.0, np.pi + 0.5, 0.0, 0.0]),
                   np.array([0.0, np.pi, 0.0, 0.0]), dt * np.diag([0.1, 1.0, 0.01, 0.01]),
                   dt * 0.001 * np.eye(1), np.diag([200.0, 200.0, 10.0, 10.0]),
                   np.zeros((1, N - 1)), beta=beta, delta=1e-2, sigma=0.02)


# ---- quadruped (mini_cheetah-scale) -------------------------------------------------
def quadruped_stand(sysm: systems.AnalyticSystem, knee=1.6):
    """Standing pose q0 (18,) and the joint torques u_stand (12,) that hold it.

    Plays the role of ``q0``/``u_stand`` in mini_cheetah.py:41-49: level body, legs at
    (0, hip, 1.6) with the hip angle (about -0.75) that puts each foot under its hip so the
    pose is an exact equilibrium, body height such that the compliant feet carry the weight.
    """
    p = sysm.params
    mass, l1, l2, l3, rf, E, g = p[2], p[10], p[11], p[12], p[15], p[16], p[19]
    hip = float(np.arctan2(-l3 * np.sin(knee), l2 + l3 * np.cos(knee)))
    Fn = mass * g / 4.0
    depth = np.sqrt(Fn / (np.pi * E))
    for _ in range(50):  # Newton on pi E d^2 (1 - 2d/(3 rf)) = Fn
        f = np.pi * E * depth ** 2 * (1 - 2 * depth / (3 * rf)) - Fn
        df = 2 * np.pi * E * depth * (1 - depth / rf)
        depth -= f / df
    lx = -l2 * np.sin(hip) - l3 * np.sin(hip + knee)
    lz = -l2 * np.cos(hip) - l3 * np.cos(hip + knee)
    q0 = np.zeros(18)
    q0[2] = rf - depth - lz
    u = np.zeros(12)
    for leg in range(4):
        sd = 1.0 if (leg & 1) else -1.0
        q0[6 + 3 * leg: 9 + 3 * leg] = [0.0, hip, knee]
        # u = -J^T (0, 0, Fn) with abad angle 0 (csrc/models.h Quadruped::step)
        u[3 * leg: 3 * leg + 3] = [-(sd * l1) * Fn, lx * Fn, -(l3 * np.sin(hip + knee)) * Fn]
    return q0, u


def quadruped(N: int = 200, target_vel: float = 1.0, keypoints=None) -> Problem:
    """mini_cheetah.py:22-69,147-180 re-expressed for the n=36 Euler-angle model
    (script horizon 50; config C4 uses N=200).  Batch seeds perturb x0 by sigma = 0.01
    (SURVEY 8d)."""
    sysm = systems.quadruped(dt=4e-3)
    dt = sysm.dt
    q0, u_stand = quadruped_stand(sysm)
    x0 = np.hstack([q0, np.zeros(18)])
    x_nom = x0.copy()
    T = N * dt
    x_nom[0] += target_vel * T      # base x position   (mini_cheetah.py:56)
    x_nom[18] += target_vel         # base x velocity   (mini_cheetah.py:57)
    Qq_base = np.hstack([np.ones(3), 3.0 * np.ones(3)])   # position 1, orientation 3 (:60-61)
    Qv_base = np.ones(6)
    Qq_legs, Qv_legs = np.zeros(12), 0.01 * np.ones(12)
    Q = np.diag(np.hstack([Qq_base, Qq_legs, 0.01 * Qv_base, Qv_legs]))
    R = 0.01 * np.eye(12)
    Qf = np.diag(np.hstack([5 * Qq_base, 0.1 + Qq_legs, Qv_base, Qv_legs]))
    u_guess = np.repeat(u_stand[:, None], N - 1, axis=1)
    return Problem("quadruped", sysm, N, x0, x_nom, dt * Q, dt * R, Qf, u_guess, beta=0.5,
                   delta=1e-2, gamma=0.0, keypoints=keypoints, sigma=0.01,
                   extra={"u_stand": u_stand, "target_vel": target_vel})


def quadruped_quat(N: int = 50, target_vel: float = 1.0, keypoints=None) -> Problem:
    """mini_cheetah.py:22-69,147-180 in the script's own n=37 layout and horizon (T=0.2, dt=4e-3
    -> N=50): q = [quat, pos, joints], v = [omega, v, joint rates]; x_nom[4] += v*T, x_nom[22] += v."""
    sysm = systems.quadruped_quat(dt=4e-3)
    dt = sysm.dt
    q18, u_stand = quadruped_stand(sysm)
    q0 = np.hstack([[1.0, 0.0, 0.0, 0.0], q18[0:3], q18[6:18]])      # mini_cheetah.py:41-46
    x0 = np.hstack([q0, np.zeros(18)])
    x_nom = x0.copy()
    x_nom[4] += target_vel * (N * dt)     # base x position   (mini_cheetah.py:56)
    x_nom[22] += target_vel               # base x velocity   (mini_cheetah.py:57)
    Qq_base = np.ones(7)
    Qq_base[0:4] += 2                     # mini_cheetah.py:60-61
    Qv_base = np.ones(6)
    Qq_legs, Qv_legs = 0.0 * np.ones(12), 0.01 * np.ones(12)
    Q = np.diag(np.hstack([Qq_base, Qq_legs, 0.01 * Qv_base, Qv_legs]))          # :66
    R = 0.01 * np.eye(12)
    Qf = np.diag(np.hstack([5 * Qq_base, 0.1 + Qq_legs, Qv_base, Qv_legs]))      # :68
    u_guess = np.repeat(u_stand[:, None], N - 1, axis=1)
    return Problem("quadruped_quat", sysm, N, x0, x_nom, dt * Q, dt * R, Qf, u_guess, beta=0.5,
                   delta=1e-2, gamma=0.0, keypoints=keypoints, sigma=0.002,
                   extra={"u_stand": u_stand, "target_vel": target_vel})


# ---- arm + ball (kinova_gen3 / panda_fr3-scale) --------------------------------------
def arm_tip(sysm: systems.AnalyticSystem, q_arm):
    """Tool-tip position of the 7R chain of csrc/models.h ArmBall (z, y, z, y, z, y, z axes;
    link i goes d_i along the rotated local z)."""
    d, bz = sysm.params[4:11], sysm.params[18]
    R, pos = np.eye(3), np.array([0.0, 0.0, bz])
    for i in range(7):
        c, s_ = np.cos(q_arm[i]), np.sin(q_arm[i])
        Rz = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1.0]])
        Ry = np.array([[c, 0, s_], [0, 1.0, 0], [-s_, 0, c]])
        R = R @ (Rz if i % 2 == 0 else Ry)
        pos = pos + d[i] * R[:, 2]
    return pos


def arm_ball_start(sysm: systems.AnalyticSystem, ball_xy=(0.6, 0.0), press=2e-4):
    """Push pose (the role of q_push, kinova_gen3.py:47): elbow-up planar configuration whose
    tool sphere presses ``press`` metres into the ball's -x side; the ball rests on the table
    at its static penetration depth."""
    p = sysm.params
    rt, rb, mb, E, g = p[11], p[12], p[13], p[14], p[17]
    depth = np.sqrt(mb * g / (np.pi * E))
    for _ in range(50):
        f = np.pi * E * depth ** 2 * (1 - 2 * depth / (3 * rb)) - mb * g
        depth -= f / (2 * np.pi * E * depth * (1 - depth / rb))
    ball = np.array([ball_xy[0], ball_xy[1], rb - depth])
    target = ball - np.array([rt + rb - press, 0.0, 0.0])
    q = np.array([0.0, 0.9, 0.0, 1.2, 0.0, 0.6, 0.0])
    for _ in range(100):   # Gauss-Newton on the three pitch joints (planar IK)
        e = arm_tip(sysm, q) - target
        J = np.zeros((3, 3))
        for k, j in enumerate((1, 3, 5)):
            dq = np.zeros(7)
            dq[j] = 1e-6
            J[:, k] = (arm_tip(sysm, q + dq) - arm_tip(sysm, q - dq)) / 2e-6
        step = np.linalg.lstsq(J, e, rcond=None)[0]
        q[[1, 3, 5]] -= step
        if np.abs(e).max() < 1e-13:
            break
    return q, ball


def arm_ball(N: int = 400, keypoints="setInterval5", scenario="forward") -> Problem:
    """kinova_gen3.py:32-99,252-275 re-expressed for the analytic 7R arm + ball model
    (script horizon 50; config C5 uses N=400 with derivative interpolation on)."""
    sysm = systems.arm_ball(dt=1e-2)
    dt = sysm.dt
    q_arm, ball = arm_ball_start(sysm)
    ball_q = np.hstack([[1.0, 0.0, 0.0, 0.0], ball])
    x0 = np.hstack([q_arm, ball_q, np.zeros(13)])
    x_nom = x0.copy()
    if scenario == "forward":
        x_nom[11] += 0.2            # move the ball forward (kinova_gen3.py:57-58)
    else:
        x_nom[12] += 0.15           # move it to the side (kinova_gen3.py:59-60)
    Qq = np.hstack([np.zeros(7), [0, 0, 0, 0, 100, 100, 100]])      # kinova_gen3.py:74-76
    Qv = 0.1 * np.ones(13)
    Q = np.diag(np.hstack([Qq, Qv]))
    R = 0.01 * np.eye(7)
    Qfv = Qv.copy()
    Qfv[7:] *= 10.0                                                  # 10*Qv_ball, :84
    Qf = np.diag(np.hstack([Qq, Qfv]))
    if keypoints == "setInterval5":
        keypoints = derivs_interpolation("setInterval", 5, 40, 1e-4, 1e-2)
    elif keypoints == "adaptiveJerk":
        keypoints = derivs_interpolation("adaptiveJerk", 5, 40, 1e-4, 1e-2)   # kinova_gen3.py:37-42
    return Problem("arm_ball", sysm, N, x0, x_nom, dt * Q, dt * R, Qf, np.zeros((7, N - 1)),
                   beta=0.5, delta=1e-3, gamma=0.0, keypoints=keypoints, sigma=0.002)


def affine_sin(n=4, m=1, N=40, seed=0, keypoints=None) -> Problem:
    """The survey's probe problem for arbitrary (n, m)."""
    sysm = systems.random_affine_sin(n, m, seed=seed)
    rng = np.random.default_rng(seed + 1)
    return Problem(sysm.name, sysm, N, rng.standard_normal(n), np.zeros(n), 0.1 * np.eye(n),
                   0.01 * np.eye(m), 10.0 * np.eye(n), np.zeros((m, N - 1)), beta=0.5,
                   delta=1e-6, keypoints=keypoints, sigma=0.1)


CONFIGS = {
    "C1": lambda: pendulum(100),
    "C2": lambda: acrobot(40),
    "C3": lambda: cart_pole_with_wall(200, beta=0.95),
    "C4": lambda: quadruped(200),
    "C5": lambda: arm_ball(400),
}
