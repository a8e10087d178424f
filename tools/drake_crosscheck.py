"""cross-validate the analytic models against Drake on a machine that has pydrake.

The solver never calls Drake (Drake's arithmetic is unpinned here because
pydrake is not installable offline).  Where pydrake IS importable this script builds the three
contact-free plants the reference's scripts build (pendulum.py:41-47, acrobot.py:52-58,
cart_pole.py:53-59: Drake's stock URDF/SDF files, discrete MultibodyPlant with the script's time
step), pushes random (x, u) through ``CalcForcedDiscreteVariableUpdate`` exactly like
``ilqr.py:223-229`` and through the host build of this repo's model, and reports the gap of the
step and of the Jacobian (Drake AutoDiff vs. the model's forward-mode AD, ilqr.py:253-270).

    python tools/drake_crosscheck.py            # prints one line per model; exit 0 without pydrake

Written without access to pydrake (round 1): treat a failure on first use as a bug in this script.
The contact models (wall, quadruped, arm+ball) are this repo's own closed forms and have no Drake
counterpart to compare with.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = {
    # name: (Drake resource, model file kind, dt, state scale)
    "pendulum": ("drake/examples/pendulum/Pendulum.urdf", 1e-2),
    "acrobot": ("drake/examples/acrobot/Acrobot.urdf", 4e-3),
    "cart_pole": ("drake/examples/multibody/cart_pole/cart_pole.sdf", 1e-2),
}


def drake_plant(resource: str, dt: float):
    from pydrake.all import DiagramBuilder, AddMultibodyPlantSceneGraph, Parser, FindResourceOrThrow
    builder = DiagramBuilder()
    plant, _ = AddMultibodyPlantSceneGraph(builder, dt)
    parser = Parser(plant)
    path = FindResourceOrThrow(resource)
    if hasattr(parser, "AddModels"):
        parser.AddModels(path)
    else:  # older Drake (pendulum.py:42)
        parser.AddModelFromFile(path)
    plant.Finalize()
    return plant


def drake_step(plant, ctx, port, x, u):
    """ilqr.py:223-229."""
    ctx.SetDiscreteState(x)
    port.FixValue(ctx, u)
    state = ctx.get_discrete_state()
    plant.CalcForcedDiscreteVariableUpdate(ctx, state)
    return state.get_vector().value().flatten().copy()


def drake_jac(plant_ad, ctx_ad, port_ad, x, u):
    """ilqr.py:253-270."""
    from pydrake.all import InitializeAutoDiff, ExtractGradient
    xu = InitializeAutoDiff(np.hstack([x, u]))
    n = x.size
    ctx_ad.SetDiscreteState(xu[:n])
    port_ad.FixValue(ctx_ad, xu[n:])
    state = ctx_ad.get_discrete_state()
    plant_ad.CalcForcedDiscreteVariableUpdate(ctx_ad, state)
    G = ExtractGradient(state.get_vector().CopyToVector())
    return G[:, :n], G[:, n:]


def crosscheck(name: str, samples: int = 50, seed: int = 0):
    from drake_ddp_b200 import systems
    from oracle.dynamics import HostDynamics
    resource, dt = CASES[name]
    sysd = getattr(systems, name)(dt=dt)
    dyn = HostDynamics(sysd)
    plant = drake_plant(resource, dt)
    ctx = plant.CreateDefaultContext()
    port = plant.get_actuation_input_port()
    plant_ad = plant.ToAutoDiffXd()
    ctx_ad = plant_ad.CreateDefaultContext()
    port_ad = plant_ad.get_actuation_input_port()
    rng = np.random.default_rng(seed)
    worst = {"step": 0.0, "fx": 0.0, "fu": 0.0}
    for _ in range(samples):
        x = rng.uniform(-1.0, 1.0, sysd.n)
        u = rng.uniform(-1.0, 1.0, sysd.m)
        xd = drake_step(plant, ctx, port, x, u)
        xm = dyn.step(x, u)
        worst["step"] = max(worst["step"], float(np.abs(xd - xm).max()))
        fxd, fud = drake_jac(plant_ad, ctx_ad, port_ad, x, u)
        fxm, fum = dyn.jac(x, u)
        worst["fx"] = max(worst["fx"], float(np.abs(fxd - fxm).max()))
        worst["fu"] = max(worst["fu"], float(np.abs(fud - fum).max()))
    return worst


def main():
    try:
        import pydrake  # noqa: F401
    except ImportError:
        print("pydrake is not importable here: nothing to cross-check (Drake parity stays unpinned)")
        return 0
    for name in CASES:
        w = crosscheck(name)
        print(f"{name}: max |f_drake - f_model| = {w['step']:.3e}, |fx| gap {w['fx']:.3e}, |fu| gap {w['fu']:.3e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
