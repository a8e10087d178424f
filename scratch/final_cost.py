"""North-star check: final cost of a full Solve() against the oracle (C4 problem, a few trajectories)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import _lib, problems
from tests.helpers import make_gpu, make_oracle
prob = problems.quadruped(200)
B = 4
x0 = prob.batch_x0(B, seed=0)
s = make_gpu(prob, B=B, x0=x0)
s.begin_solve()
it = 0
while s.iterate() > 0 and it < 60:
    it += 1
cost, iters, K = s.cost, s.get_int(_lib.I_ITERS), s.get(_lib.K)
for b in range(B):
    o = make_oracle(prob, x0=x0[b])
    o.solve(max_iters=60)
    Lo = o.trace[-1].L
    print(f"trajectory {b}: gpu cost {cost[b]:.12f} after {iters[b]} iterations, oracle {Lo:.12f} after {len(o.trace)}; "
          f"rel err {abs(cost[b]-Lo)/abs(Lo):.2e}; K rel err {np.abs(K[b]-o.K).max()/np.abs(o.K).max():.2e}")
