"""Native model descriptors accepted as the ``system`` argument of the solver.

The reference takes a discrete-time Drake ``System`` (/root/reference/ilqr.py:21-58)
and only ever uses it for (a) the state/input sizes, (b) the discrete update and its
AutoDiff clone.  Here a system is a *fixed analytic model* compiled into the CUDA
library (``csrc/models.h``), identified by ``model_id`` plus a parameter vector; the
first parameter is always the time step.  Constants are taken from the reference's
example scripts / Drake's stock models where the survey could recover them
(SURVEY.md section 8c); the contact models are this repo's closed forms.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# keep in sync with csrc/models.h (enum ModelId)
MODEL_PENDULUM = 0
MODEL_ACROBOT = 1
MODEL_CARTPOLE = 2
MODEL_CARTPOLE_WALL = 3
MODEL_QUADRUPED = 4
MODEL_ARM_BALL = 5
MODEL_QUADRUPED_QUAT = 6
MODEL_AFFINE_SIN = {(4, 1): 10, (6, 2): 11, (27, 7): 12, (36, 12): 13, (37, 12): 14}


@dataclass
class AnalyticSystem:
    """A discrete-time system x+ = f(x, u) evaluated in-kernel.

    Duck-types the two Drake calls example scripts make on the plant outside the solver:
    ``time_step()`` (/root/reference/ilqr.py:725) and ``IsDifferenceEquationSystem()``
    (/root/reference/ilqr.py:37).
    """

    name: str
    model_id: int
    n: int
    m: int
    params: np.ndarray = field(repr=False)

    @property
    def dt(self) -> float:
        return float(self.params[0])

    def time_step(self) -> float:
        return self.dt

    def IsDifferenceEquationSystem(self):
        return (True, self.dt)

    def GetSubsystemByName(self, name):
        return self

    def num_multibody_states(self) -> int:
        return self.n


def pendulum(dt=1e-2, mass=1.0, length=0.5, damping=0.1, g=9.81) -> AnalyticSystem:
    """Drake's Pendulum.urdf constants (pendulum.py:41)."""
    return AnalyticSystem("pendulum", MODEL_PENDULUM, 2, 1,
                          np.array([dt, mass, length, damping, g], dtype=np.float64))


def acrobot(dt=4e-3, m1=1.0, m2=1.0, l1=1.0, lc1=0.5, lc2=1.0, Ic1=0.083, Ic2=0.33,
            b1=0.1, b2=0.1, g=9.81) -> AnalyticSystem:
    """Drake's Acrobot.urdf constants (acrobot.py:52), elbow actuated."""
    return AnalyticSystem("acrobot", MODEL_ACROBOT, 4, 1,
                          np.array([dt, m1, m2, l1, lc1, lc2, Ic1, Ic2, b1, b2, g], dtype=np.float64))


def cart_pole(dt=1e-2, mc=10.0, mp=1.0, length=0.5, g=9.81) -> AnalyticSystem:
    """Drake's cart_pole.sdf constants (cart_pole.py:53)."""
    return AnalyticSystem("cart_pole", MODEL_CARTPOLE, 4, 1,
                          np.array([dt, mc, mp, length, g, 0.0, 0.0, 0.0, 1.0], dtype=np.float64))


def cart_pole_with_wall(dt=1e-2, mc=10.0, mp=1.0, length=0.5, g=9.81, ball_radius=0.05,
                        modulus=2e6, wall_face_x=-0.45, substeps=4) -> AnalyticSystem:
    """cart_pole_with_wall.py:22-52,64-97: ball r=.05 on the pole tip, E=2e6, wall face x=-0.45."""
    return AnalyticSystem("cart_pole_with_wall", MODEL_CARTPOLE_WALL, 4, 1,
                          np.array([dt, mc, mp, length, g, ball_radius, modulus, wall_face_x,
                                    float(substeps)], dtype=np.float64))


def quadruped(dt=4e-3, substeps=2, mass=8.252, inertia=(0.07, 0.26, 0.242),
              joint_inertia=(0.03, 0.03, 0.03), joint_damping=1.0,
              l_abad=0.062, l_thigh=0.209, l_shank=0.19, hip_x=0.19, hip_y=0.049,
              foot_radius=0.0175, modulus=1.75e5, mu=0.6, v_stiction=0.2, g=9.81) -> AnalyticSystem:
    """mini_cheetah-scale lumped quadruped (masses/lengths from mini_cheetah_mesh.urdf).

    ``modulus`` is the modulus of the model's compliant-sphere/rigid-plane closed form
    F = pi E' d^2 (1 - 2d/3R).  The script puts RIGID feet (r = 0.0175) on a COMPLIANT ground box of
    thickness 1 m with hydroelastic modulus E = 5e6 (mini_cheetah.py:75-101): the box's pressure
    field rises linearly from 0 at the surface to E at the medial plane 0.5 m down, so a foot that
    sinks d feels F = int (E/0.5)(d - rho^2/2R) 2 pi rho d rho = 2 pi E R d^2.  Matching the d^2
    term gives E' = 2 E R / (1 m) = 1.75e5, the default here (rest penetration 6 mm).  Used directly as
    E' the script's 5e6 makes the contact 29x stiffer than the script's own scene and the N=200
    solve chaotic."""
    p = [dt, float(substeps), mass, *inertia, *joint_inertia, joint_damping, l_abad, l_thigh,
         l_shank, hip_x, hip_y, foot_radius, modulus, mu, v_stiction, g]
    # masses and inertias enter the model through their reciprocals (csrc/models.h), rounded here once
    p += [1.0 / mass, *(1.0 / np.asarray(inertia, dtype=np.float64)),
          *(1.0 / np.asarray(joint_inertia, dtype=np.float64))]
    p += [2.0 / (3.0 * foot_radius)]      # the contact law's 2/(3R), same reason
    return AnalyticSystem("quadruped", MODEL_QUADRUPED, 36, 12, np.array(p, dtype=np.float64))


def quadruped_quat(**kw) -> AnalyticSystem:
    """The quadruped with a quaternion floating base in the reference's n=37 state layout
    (mini_cheetah.py:41-57); same parameters as ``quadruped``."""
    base = quadruped(**kw)
    return AnalyticSystem("quadruped_quat", MODEL_QUADRUPED_QUAT, 37, 12, base.params)


def arm_ball(dt=1e-2, substeps=4, joint_inertia=0.3, joint_damping=0.5,
             links=(0.28, 0.0, 0.42, 0.0, 0.31, 0.0, 0.17), tip_radius=0.04, ball_radius=0.1,
             ball_mass=0.5, modulus=5e5, mu=0.3, v_stiction=0.05, g=9.81, base_z=0.0,
             dissipation=5.0) -> AnalyticSystem:
    """kinova_gen3 / panda_fr3-scale 7R arm pushing a free ball on a table (n=27, m=7); ball
    radius, dissipation and friction from kinova_gen3.py:51,91-96; the modulus is 10x softer
    than the script's 5e6 so the explicit contact stays well inside its stability limit."""
    p = [dt, float(substeps), joint_inertia, joint_damping, *links, tip_radius, ball_radius,
         ball_mass, modulus, mu, v_stiction, g, base_z, dissipation]
    # parameter-only quotients of the contact laws (csrc/models.h ArmBall::contacts), rounded here once
    r_eff = tip_radius * ball_radius / (tip_radius + ball_radius)
    p += [r_eff, 2.0 / (3.0 * r_eff), 2.0 / (3.0 * ball_radius)]
    return AnalyticSystem("arm_ball", MODEL_ARM_BALL, 27, 7, np.array(p, dtype=np.float64))


def affine_sin(n: int, m: int, A: np.ndarray, B: np.ndarray, dt: float = 1.0) -> AnalyticSystem:
    """Test stub x+ = A x + B u + 0.01 sin(x) (the survey's probe dynamics), fixed (n, m) table."""
    mid = MODEL_AFFINE_SIN[(n, m)]
    A = np.asarray(A, dtype=np.float64).reshape(n, n)
    B = np.asarray(B, dtype=np.float64).reshape(n, m)
    return AnalyticSystem(f"affine_sin_{n}_{m}", mid, n, m,
                          np.concatenate([[dt], A.ravel(), B.ravel()]))


def random_affine_sin(n: int, m: int, seed: int = 0, spectral_radius: float = 0.95) -> AnalyticSystem:
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    A *= spectral_radius / max(abs(np.linalg.eigvals(A)))
    B = rng.standard_normal((n, m)) / np.sqrt(n)
    return affine_sin(n, m, A, B)
