"""Whole iterations of a batch problem (default C4): ms per iteration, phase times and checksums of
the results, to compare library builds for speed and bit-equality.  Scratch tool, not a test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.ilqr import BatchedILQR

B = int(os.environ.get("PB_BATCH", "1024"))
N = int(os.environ.get("PB_HORIZON", "200"))
iters = int(os.environ.get("PB_ITERS", "6"))
model = os.environ.get("PB_MODEL", "quadruped")
prob = getattr(problems, model)(N)
x0 = prob.batch_x0(B, seed=0)
s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
s.set_cost(prob.Q, prob.R, prob.Qf)
s.set_target(prob.x_nom)
s.set_initial_state(x0)
s.set_initial_guess(prob.u_guess)
s.begin_solve()
s.iterate()
torch.cuda.synchronize()
ts, ph = [], []
for i in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s._stream)
    s.iterate()
    e1.record(s._stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
    ph.append(s.timings_ms())
print({"lib": os.path.basename(_lib.LIB_PATH), "ms": [round(t, 3) for t in ts],
       "phases_last": {k: round(v, 3) for k, v in ph[-1].items()},
       "cost_sum": float(np.sum(s.cost)), "K_sum": float(np.abs(s.get(_lib.K)).sum()),
       "kappa_sum": float(np.abs(s.get(_lib.KAPPA)).sum()), "fx_sum": float(np.abs(s.get(_lib.FX)).sum())})
