"""Actual GPU-vs-oracle errors per iteration (scratch): quadruped N=60, 4 iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import _lib, problems
from tests.helpers import make_gpu, make_oracle, relerr
for name, prob in (("quadruped", problems.quadruped(60)), ("quadruped200", problems.quadruped(200))):
    s = make_gpu(prob); o = make_oracle(prob)
    s.begin_solve(); L = np.inf
    for it in range(3):
        s.iterate(); rec = o.iterate(L); L = rec.L
        print(name, "it", it, "cost rel", abs(s.cost[0] - rec.L) / abs(rec.L), "K", relerr(s.get(_lib.K)[0], o.K),
              "kappa", relerr(s.get(_lib.KAPPA)[0], o.kappa), "x", relerr(s.get(_lib.X_BAR)[0], o.x_bar),
              "fx", relerr(s.get(_lib.FX)[0], o.fx), "ls", int(s.get_int(_lib.I_LS_ITERS)[0]), rec.ls_iters)
