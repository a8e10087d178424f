"""Analytic models: host build of csrc/models.h (what the oracle calls) against independent
NumPy restatements, exact AD Jacobians against central differences, physical sanity."""
import numpy as np
import pytest

from drake_ddp_b200 import problems, systems
from oracle.dynamics import HostDynamics


def np_pendulum(x, u, p):
    h, m, l, b, g = p
    a = (u[0] - b * x[1] - m * g * l * np.sin(x[0])) / (m * l * l)
    v = x[1] + h * a
    return np.array([x[0] + h * v, v])


def np_cartpole(x, u, p):
    h, mc, mp, l, g = p[:5]
    s, c = np.sin(x[1]), np.cos(x[1])
    M = np.array([[mc + mp, mp * l * c], [mp * l * c, mp * l * l]])
    rhs = np.array([u[0] + mp * l * x[3] ** 2 * s, -mp * g * l * s])
    a = np.linalg.solve(M, rhs)
    v = x[2:] + h * a
    return np.hstack([x[:2] + h * v, v])


def np_acrobot(x, u, p):
    h, m1, m2, l1, lc1, lc2, Ic1, Ic2, b1, b2, g = p
    I1, I2 = Ic1 + m1 * lc1 ** 2, Ic2 + m2 * lc2 ** 2
    q1, q2, v1, v2 = x
    c2, s2 = np.cos(q2), np.sin(q2)
    M = np.array([[I1 + I2 + m2 * l1 ** 2 + 2 * m2 * l1 * lc2 * c2, I2 + m2 * l1 * lc2 * c2],
                  [I2 + m2 * l1 * lc2 * c2, I2]])
    C = np.array([[-2 * m2 * l1 * lc2 * s2 * v2, -m2 * l1 * lc2 * s2 * v2],
                  [m2 * l1 * lc2 * s2 * v1, 0.0]])
    tau_g = np.array([-m1 * g * lc1 * np.sin(q1) - m2 * g * (l1 * np.sin(q1) + lc2 * np.sin(q1 + q2)),
                      -m2 * g * lc2 * np.sin(q1 + q2)])
    rhs = tau_g + np.array([0.0, u[0]]) - C @ np.array([v1, v2]) - np.array([b1 * v1, b2 * v2])
    a = np.linalg.solve(M, rhs)
    v = np.array([v1, v2]) + h * a
    return np.hstack([x[:2] + h * v, v])


@pytest.mark.parametrize("sysm,fn", [(systems.pendulum(), np_pendulum), (systems.cart_pole(), np_cartpole),
                                     (systems.acrobot(), np_acrobot)])
def test_host_model_matches_numpy_restatement(sysm, fn):
    dyn = HostDynamics(sysm)
    rng = np.random.default_rng(0)
    for _ in range(20):
        x, u = rng.standard_normal(sysm.n), rng.standard_normal(sysm.m)
        np.testing.assert_allclose(dyn.step(x, u), fn(x, u, sysm.params), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["pendulum", "acrobot", "cart_pole", "cart_pole_with_wall", "quadruped",
                                  "quadruped_quat", "arm_ball"])
def test_jacobian_matches_central_differences(name):
    prob = getattr(problems, name)()
    dyn = HostDynamics(prob.system)
    rng = np.random.default_rng(1)
    n, m = prob.system.n, prob.system.m
    for trial in range(3):
        x = prob.x0 + 0.05 * rng.standard_normal(n)
        u = prob.u_guess[:, 0] + 0.1 * rng.standard_normal(m)
        fx, fu = dyn.jac(x, u)
        h = 1e-6
        fxn = np.stack([(dyn.step(x + h * e, u) - dyn.step(x - h * e, u)) / (2 * h) for e in np.eye(n)], axis=1)
        fun = np.stack([(dyn.step(x, u + h * e) - dyn.step(x, u - h * e)) / (2 * h) for e in np.eye(m)], axis=1)
        scale = max(1.0, np.abs(fx).max())
        assert np.abs(fx - fxn).max() < 1e-6 * scale
        assert np.abs(fu - fun).max() < 1e-6 * max(1.0, np.abs(fu).max())


def test_wall_contact_force_closed_form():
    """Sphere/plane force pi E (d^2 - 2 d^3 / (3R)) equals the integral of p = E(1 - r/R) over
    the contact disk, and the wall pushes the tip away (+x)."""
    R, E = 0.05, 2e6
    for depth in (0.001, 0.01, 0.03):
        hh = R - depth
        rho = np.linspace(0.0, np.sqrt(R * R - hh * hh), 200001)
        integrand = E * (1 - np.sqrt(hh * hh + rho * rho) / R) * 2 * np.pi * rho
        F_num = np.sum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(rho))
        F_cf = np.pi * E * depth ** 2 * (1 - 2 * depth / (3 * R))
        assert abs(F_num - F_cf) < 1e-6 * F_cf
    sysm = systems.cart_pole_with_wall()
    dyn = HostDynamics(sysm)
    # tip ball centre 1 cm inside the contact zone, at rest, no input
    th = np.pi + 0.5
    x_cart = (-0.45 + 0.05 - 0.01) - 0.5 * np.sin(th)
    free = HostDynamics(systems.cart_pole())
    x = np.array([x_cart, th, 0.0, 0.0])
    assert dyn.step(x, [0.0])[2] > free.step(x, [0.0])[2]


def test_quadruped_standing_is_equilibrium():
    prob = problems.quadruped(50)
    dyn = HostDynamics(prob.system)
    x = prob.x0.copy()
    for _ in range(100):
        x = dyn.step(x, prob.extra["u_stand"])
    assert np.abs(x - prob.x0).max() < 1e-9


def test_pendulum_energy_without_damping():
    sysm = systems.pendulum(dt=1e-3, damping=0.0)
    dyn = HostDynamics(sysm)
    m, l, g = 1.0, 0.5, 9.81
    energy = lambda x: 0.5 * m * l * l * x[1] ** 2 - m * g * l * np.cos(x[0])
    x = np.array([1.0, 0.0])
    e0 = energy(x)
    for _ in range(2000):
        x = dyn.step(x, [0.0])
    assert abs(energy(x) - e0) < 2e-2 * abs(e0)   # symplectic Euler: bounded energy error


def test_quaternion_and_euler_quadrupeds_agree():
    """The n=37 quaternion-base model (reference layout) and the n=36 Euler-angle model are the same
    physics: identical torques give the same base/joint motion up to the O(h^2) difference of the
    two attitude integrators."""
    p37, p36 = problems.quadruped_quat(50), problems.quadruped(50)
    d37, d36 = HostDynamics(p37.system), HostDynamics(p36.system)
    rng = np.random.default_rng(0)
    xa, xb = p37.x0.copy(), p36.x0.copy()
    for _ in range(30):
        u = p37.extra["u_stand"] + 0.5 * rng.standard_normal(12)
        xa, xb = d37.step(xa, u), d36.step(xb, u)
    assert np.abs(xa[4:7] - xb[0:3]).max() < 1e-6          # base position
    assert np.abs(xa[7:19] - xb[6:18]).max() < 1e-6        # joints
    qw, qx, qy, qz = xa[:4] / np.linalg.norm(xa[:4])
    rpy = [np.arctan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy)), np.arcsin(2 * (qw * qy - qz * qx)),
           np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))]
    assert np.abs(np.array(rpy) - xb[3:6]).max() < 1e-6
    assert p37.system.n == 37 and p37.x_nom[4] > p37.x0[4] and p37.x_nom[22] == 1.0   # mini_cheetah.py:56-57
