"""Host-side logic that needs no GPU: problem fixtures, drop-in module names, batch sharding
(world_size-2 gloo run with the CPU oracle standing in for each rank's local solver)."""
import os
import sys

import numpy as np
import pytest

from drake_ddp_b200 import dist as ddist
from drake_ddp_b200 import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_shapes():
    want = {"C1": (2, 1, 100), "C2": (4, 1, 40), "C3": (4, 1, 200), "C4": (36, 12, 200), "C5": (27, 7, 400)}
    for name, (n, m, N) in want.items():
        p = problems.CONFIGS[name]()
        assert (p.system.n, p.system.m, p.N) == (n, m, N)
        assert p.Q.shape == (n, n) and p.R.shape == (m, m) and p.Qf.shape == (n, n)
        assert p.u_guess.shape == (m, N - 1) and p.x0.shape == (n,) and p.x_nom.shape == (n,)
    assert problems.CONFIGS["C5"]().keypoints.keypoint_method == "setInterval"


def test_dropin_module_names():
    """Scripts do `from ilqr import IterativeLinearQuadraticRegulator` and
    `import utils_derivs_interpolation` (pendulum.py:11, acrobot.py:115)."""
    sys.path.insert(0, ROOT)
    import utils_derivs_interpolation as udi
    cfg = udi.derivs_interpolation("adaptiveJerk", 5, 100, 7e-4, 5e-5)   # positional, acrobot.py:115
    assert cfg.minN == 5 and udi.index_tuple(1, 2).end_index == 2
    src = open(os.path.join(ROOT, "ilqr.py")).read()
    assert "IterativeLinearQuadraticRegulator" in src


def test_shard_range_covers_batch():
    for B in (1, 7, 50, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [ddist.shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = ddist.shard_sizes(B, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == B
    assert ddist.shard_sizes(1024, 8) == [128] * 8      # C4: 128 trajectories per GPU


def _worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tests.helpers import make_oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    prob = problems.pendulum(30)
    x0 = prob.batch_x0(B, seed=0)
    lo, hi = ddist.shard_range(B, rank, world)
    costs = []
    for b in range(lo, hi):
        o = make_oracle(prob, x0=x0[b])
        o.solve(max_iters=3)
        costs.append(o.trace[-1].L)
    full = ddist.all_gather_ragged(torch.tensor(costs, dtype=torch.float64), B)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_costs_equal_single_process_gloo():
    """world_size 2, gloo: sharding the batch and all-gathering per-trajectory costs reproduces
    the single-process result bit for bit (ragged shards: B=5 -> 3 + 2)."""
    import torch.multiprocessing as mp
    from tests.helpers import make_oracle
    B, world = 5, 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = problems.pendulum(30)
    x0 = prob.batch_x0(B, seed=0)
    want = []
    for b in range(B):
        o = make_oracle(prob, x0=x0[b])
        o.solve(max_iters=3)
        want.append(o.trace[-1].L)
    np.testing.assert_array_equal(got, np.array(want))


def test_drake_system_is_accepted_and_mapped_to_the_analytic_model():
    """the constructor takes the reference's ``system`` argument -- a discrete-time
    Drake System, read only through the calls ilqr.py:37-58,725 makes on it -- and selects the
    matching analytic model.  ShimSystem (the object that lets the unmodified reference run here)
    duck-types exactly those calls, so it stands in for Drake."""
    from drake_ddp_b200 import systems
    from drake_ddp_b200.ilqr import _as_system
    from oracle.pydrake_shim import ShimSystem

    class Named(ShimSystem):
        def __init__(self, sysm, name):
            super().__init__(sysm)
            self._name = name

        def get_name(self):
            return self._name

    for sysm, name in ((systems.pendulum(dt=5e-3), "Pendulum"), (systems.acrobot(), "Acrobot"),
                       (systems.cart_pole(), "CartPole"), (systems.quadruped_quat(), "mini_cheetah"),
                       (systems.quadruped(), "quadruped"), (systems.arm_ball(), "gen3")):
        got = _as_system(Named(sysm, name), input_port_index=3)
        assert (got.model_id, got.n, got.m) == (sysm.model_id, sysm.n, sysm.m)
        assert got.dt == sysm.dt
        np.testing.assert_array_equal(got.params, type(got)(got.name, got.model_id, got.n, got.m, sysm.params).params)
    assert _as_system(systems.pendulum()) is not None          # native descriptor passes through
    with pytest.raises(TypeError, match="no analytic model"):
        _as_system(Named(systems.random_affine_sin(6, 2), "mystery"))
    with pytest.raises(TypeError, match="AnalyticSystem or a discrete-time Drake System"):
        _as_system(object())
