"""Experiment: the batch as K independent sub-batches on K streams / host threads (phases of different
sub-batches overlap on the GPU).  python scratch/two_streams.py [K ...]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from drake_ddp_b200 import _lib, problems
torch.cuda.set_device(0)
prob = problems.quadruped(200)
B, T, m = 1024, 199, 12
for K in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    Bs = B // K
    solvers = []
    for k in range(K):
        s = bench.make_solver(prob, Bs)
        x0 = prob.batch_x0(B, seed=0)[k * Bs:(k + 1) * Bs]
        u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (Bs, T, m)))
        s.reset(); s.set_initial_state(x0); s.set_initial_guess(u0); s.begin_solve()
        solvers.append(s)
    W, N = 5, 20
    def run(s, n):
        for _ in range(n):
            s.iterate()
    for s in solvers: run(s, W)
    torch.cuda.synchronize()
    it0 = [s.get_int(_lib.I_ITERS).sum() for s in solvers]
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(s, N)) for s in solvers]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    units = sum(int(s.get_int(_lib.I_ITERS).sum() - i0) for s, i0 in zip(solvers, it0))
    print(f"K={K}: {units / dt:.0f} trajectory-iterations/s, {dt / N * 1e3:.3f} ms per batch iteration")
    del solvers
