"""C1 / C3: a few iLQR iterations for an ncu launch list (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drake_ddp_b200 import problems
from drake_ddp_b200.ilqr import BatchedILQR
name, N, A = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prob = getattr(problems, name)(N) if name != "cart_pole_with_wall" else problems.cart_pole_with_wall(N, beta=0.95)
s = BatchedILQR(prob.system, prob.N, batch=1, delta=prob.delta, beta=prob.beta, gamma=prob.gamma, ls_parallel=A)
s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
s.set_initial_state(prob.x0[None].copy()); s.set_initial_guess(prob.u_guess)
s.begin_solve()
for _ in range(4):
    s.iterate()
