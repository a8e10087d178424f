"""Driven by tests/test_gpu_configs.py::test_cost_gather_side_stream: CostGather (the path's one
collective) on a one-rank NCCL group: the gathered vector of every iteration equals the solver's
cost vector of that iteration, with the gathers issued back to back (ping-pong buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from drake_ddp_b200 import problems
from drake_ddp_b200.dist import CostGather
from drake_ddp_b200.ilqr import BatchedILQR

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
prob = problems.acrobot(40)
B = 6
s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
s.set_cost(prob.Q, prob.R, prob.Qf)
s.set_target(prob.x_nom)
s.set_initial_state(prob.batch_x0(B, seed=1))
s.set_initial_guess(prob.u_guess)
s.begin_solve()
g = CostGather(s, B)
seen = []
for it in range(5):
    s.iterate()
    g.issue()
    if it >= 1:                      # two gathers in flight before the first result is read
        seen.append((g.result().cpu().numpy().copy(), s.cost.copy()))
for got, want in seen:
    assert np.array_equal(got[:B], want), (got, want)
dist.destroy_process_group()
print("cost gather ok")
