"""Does splitting the batch over concurrent streams help? (latency-bound rollout vs tensor-bound backward)"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.ilqr import BatchedILQR
prob = problems.quadruped(200)
def make(B, seed, stream):
    with torch.cuda.stream(stream):
        s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma, ls_parallel=8)
        s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
        s.set_initial_state(prob.batch_x0(1024, seed=0)[seed*B:(seed+1)*B]); s.set_initial_guess(prob.u_guess); s.begin_solve()
    return s
for S in (1, 2, 4):
    B = 1024 // S
    streams = [torch.cuda.Stream() for _ in range(S)]
    solvers = [make(B, i, streams[i]) for i in range(S)]
    def work(i, k):
        for _ in range(k): solvers[i].iterate()
    for phase, k in (("warm", 3), ("timed", 8)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i, k)) for i in range(S)]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if phase == "timed":
            print(f"S={S} B={B}: {dt/k*1e3:.2f} ms per iteration of all {S*B} trajectories")
    del solvers
