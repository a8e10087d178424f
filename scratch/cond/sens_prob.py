"""Oracle self-sensitivity of a full Solve() for any problem factory (CPU only).
usage: python scratch/cond/sens_prob.py <factory-expr> [nb] [sigma]   e.g. "problems.quadruped_quat(200)" 32 0.01"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import multiprocessing as mp
from drake_ddp_b200 import problems, systems
from tests.helpers import make_oracle

def work(args):
    expr, b, sigma = args
    prob = eval(expr)
    if sigma is not None: prob.sigma = sigma
    x0 = prob.batch_x0(b + 1, seed=0)[b]
    out = []
    for scale in (1.0, 1.0 + 8e-16, 1.0 - 8e-16):
        o = make_oracle(prob, x0=x0 * scale)
        try:
            o.solve(max_iters=100); err = None
        except RuntimeError as e:
            err = str(e)
        out.append(([r.L for r in o.trace], err, o.K.copy()))
    return b, out

if __name__ == "__main__":
    expr = sys.argv[1]; nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    sigma = float(sys.argv[3]) if len(sys.argv) > 3 else None
    t0 = time.time()
    with mp.Pool(8) as pool:
        res = pool.map(work, [(expr, b, sigma) for b in range(nb)], chunksize=1)
    worst = []
    for b, out in res:
        L0, e0, K0 = out[0]
        rels = [abs(L[-1] - L0[-1]) / abs(L0[-1]) for (L, e, K) in out[1:]]
        krel = [np.abs(K - K0).max() / np.abs(K0).max() for (L, e, K) in out[1:]]
        worst.append(max(rels))
        print(f"traj {b}: final {L0[-1]:.9f} it={len(L0)} err={e0} perturbed: " + ", ".join(
            f"it={len(L)} rel={r:.1e} K={k:.1e}" for (L, e, K), r, k in zip(out[1:], rels, krel)))
    w = np.array(worst)
    print(f"ok(<=1e-7): {(w <= 1e-7).sum()}/{len(w)}   ok(<=1e-5): {(w<=1e-5).sum()}/{len(w)}  time {time.time()-t0:.0f}s")
