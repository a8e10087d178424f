import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, multiprocessing as mp
from drake_ddp_b200 import problems, systems
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation
from tests.helpers import make_oracle
def mk(kw):
    kw = dict(kw)
    N = int(kw.pop("N", 400)); press = float(kw.pop("press", 2e-4)); sigma = float(kw.pop("sigma", 0.002)); kp = kw.pop("kp", "setInterval5")
    dx = float(kw.pop("dx", 0.2)); kw.pop("nb", None)
    prob = problems.arm_ball(N, keypoints=kp)
    sysm = systems.arm_ball(dt=1e-2, **kw)
    q_arm, ball = problems.arm_ball_start(sysm, press=press)
    x0 = np.hstack([q_arm, [1.0, 0, 0, 0], ball, np.zeros(13)])
    xn = x0.copy(); xn[11] += dx
    prob.system, prob.x0, prob.x_nom, prob.sigma = sysm, x0, xn, sigma
    if kp == "none": prob.keypoints = None
    return prob
def work(a):
    kw, b = a
    prob = mk(kw)
    x0 = prob.batch_x0(b + 1, seed=0)[b]
    res = []
    for sc in (1.0, 1 + 8e-16):
        o = make_oracle(prob, x0=x0 * sc)
        try: o.solve(max_iters=100); err = ""
        except RuntimeError: err = "LSFAIL"
        res.append(([r.L for r in o.trace], err, o.x_bar[-1, 11]))
    return b, res
if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        try: kw[k] = float(v) if ("." in v or "e" in v) else int(v)
        except ValueError: kw[k] = v
    nb = int(kw.get("nb", 16))
    with mp.Pool(8) as pool: out = pool.map(work, [(kw, b) for b in range(nb)], chunksize=1)
    its, fails, bad = [], 0, 0
    for b, res in out:
        (L, e, bx), (L2, e2, _) = res
        rel = abs(L[-1] - L2[-1]) / abs(L[-1])
        its.append(len(L)); fails += bool(e); bad += rel > 1e-7
        if nb <= 16: print(f"  {b}: it={len(L)} {e} L0={L[0]:.2f} Lf={L[-1]:.3f} ballx={bx:.3f} selfrel={rel:.0e}")
    print(f"iters: min {min(its)} med {np.median(its)} max {max(its)}; linesearch failures {fails}/{nb}; ill-conditioned {bad}/{nb}")
