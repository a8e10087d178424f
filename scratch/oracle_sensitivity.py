"""How reproducible is a full Solve() of the C4 problem at all?  The oracle against itself with
x0 perturbed by one part in 1e15 (CPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import problems
from tests.helpers import make_oracle
prob = problems.quadruped(200)
x0 = prob.batch_x0(4, seed=0)
for b in range(4):
    res = []
    for scale in (1.0, 1.0 + 1e-15, 1.0 - 1e-15):
        o = make_oracle(prob, x0=x0[b] * scale)
        o.solve(max_iters=60)
        res.append((o.trace[-1].L, len(o.trace)))
    L0 = res[0][0]
    print(f"trajectory {b}: oracle final cost {L0:.12f} ({res[0][1]} it); perturbed x0: "
          + ", ".join(f"{L:.12f} ({n} it, rel {abs(L - L0) / abs(L0):.1e})" for L, n in res[1:]))
