"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import problems, _lib
from drake_ddp_b200.ilqr import BatchedILQR
name = os.environ.get("SAN_MODEL", "quadruped")
prob = getattr(problems, name)(12)
B = int(os.environ.get("SAN_B", "3"))
s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma,
                ls_parallel=int(os.environ.get("SAN_A", "8")))
s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
s.set_initial_state(prob.batch_x0(B, seed=0)); s.set_initial_guess(prob.u_guess)
s.begin_solve()
for _ in range(2):
    s.iterate()
print(name, "cost", s.cost)
