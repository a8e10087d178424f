"""Conditioning study (CPU only): oracle vs oracle under a one-ulp change of x0, per iteration.
usage: python scratch/cond/sens.py [key=value ...]  (system kwargs, sigma=, N=, vel=, nb=, seed=)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import multiprocessing as mp
from drake_ddp_b200 import problems, systems
from tests.helpers import make_oracle

def build_prob(kw):
    kw = dict(kw)
    N = int(kw.pop("N", 200)); sigma = float(kw.pop("sigma", 0.002)); vel = float(kw.pop("vel", 1.0))
    kw.pop("nb", None); kw.pop("seed", None); kw.pop("verbose", None)
    delta = kw.pop("delta", None)
    prob = problems.quadruped(N, target_vel=vel)
    if kw:
        sysm = systems.quadruped(dt=4e-3, **kw)
        prob.system = sysm
        q0, u_stand = problems.quadruped_stand(sysm)
        x0 = np.hstack([q0, np.zeros(18)])
        xn = x0.copy(); xn[0] += vel * N * sysm.dt; xn[18] += vel
        prob.x0, prob.x_nom = x0, xn
        prob.u_guess = np.repeat(u_stand[:, None], N - 1, axis=1)
    prob.sigma = sigma
    if delta is not None: prob.delta = float(delta)
    return prob

def work(args):
    kw, b, seed = args
    prob = build_prob(kw)
    x0 = prob.batch_x0(max(b + 1, 4), seed=seed)[b]
    out = []
    for scale in (1.0, 1.0 + 2e-16 * 4, 1.0 - 2e-16 * 4):
        o = make_oracle(prob, x0=x0 * scale)
        try:
            o.solve(max_iters=80)
            out.append(([r.L for r in o.trace], [r.ls_iters for r in o.trace], None))
        except RuntimeError as e:
            out.append(([r.L for r in o.trace], [r.ls_iters for r in o.trace], str(e)))
    return b, out

if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        try: kw[k] = float(v) if "." in v or "e" in v else int(v)
        except ValueError: kw[k] = v
    nb = int(kw.get("nb", 8)); seed = int(kw.get("seed", 0)); verbose = int(kw.get("verbose", 0))
    t0 = time.time()
    with mp.Pool(min(8, nb)) as pool:
        res = pool.map(work, [(kw, b, seed) for b in range(nb)])
    worst = []
    for b, out in res:
        (L0, ls0, e0) = out[0]
        rels = []
        for (L, ls, e) in out[1:]:
            rels.append(abs(L[-1] - L0[-1]) / abs(L0[-1]) if L and L0 else float("nan"))
        worst.append(max(rels))
        print(f"traj {b}: final {L0[-1]:.9f} it={len(L0)} err={e0}  perturbed: " +
              ", ".join(f"it={len(L)} rel={r:.1e}" for (L, ls, e), r in zip(out[1:], rels)))
        if verbose:
            for (L, ls, e) in out[1:]:
                k = min(len(L), len(L0))
                print("    per-iter rel:", " ".join(f"{abs(L[i]-L0[i])/abs(L0[i]):.0e}" for i in range(k)))
            print("    costs:", " ".join(f"{l:.4f}" for l in L0))
            print("    ls:", ls0)
    w = np.array(worst)
    print(f"ok(<=1e-7): {(w <= 1e-7).sum()}/{len(w)}   ok(<=1e-5): {(w<=1e-5).sum()}/{len(w)}  time {time.time()-t0:.0f}s")
