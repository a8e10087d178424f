import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import _lib, problems
from tests.helpers import make_gpu, make_oracle, relerr
prob = problems.quadruped(200)
B = 1024
x0 = prob.batch_x0(B, seed=0); x0[-4:] = x0[:4]
s = make_gpu(prob, B=B, A=2, x0=x0)
s.begin_solve()
spot = [0, 1, 517]
oracles = [make_oracle(prob, x0=x0[b]) for b in spot]
Ls = [np.inf]*3
prev = np.full(B, np.inf)
for it in range(2):
    s.iterate()
    cost, status = s.cost, s.status
    ok = status != 2
    print("it", it, "ok", ok.mean(), "decreasing", np.all(cost[ok] < prev[ok]), "dup equal", np.array_equal(cost[-4:], cost[:4]), np.array_equal(s.get(_lib.K)[-4:], s.get(_lib.K)[:4]))
    for k, b in enumerate(spot):
        rec = oracles[k].iterate(Ls[k]); Ls[k] = rec.L
        print("  b", b, "cost rel", abs(cost[b]-rec.L)/abs(rec.L), "ls", s.get_int(_lib.I_LS_ITERS)[b], rec.ls_iters, "dK", relerr(s.get(_lib.K)[b], oracles[k].K), "dkappa", relerr(s.get(_lib.KAPPA)[b], oracles[k].kappa), "dfx", relerr(s.get(_lib.FX)[b], oracles[k].fx), "dx", relerr(s.get(_lib.X_BAR)[b], oracles[k].x_bar))
    prev = cost
