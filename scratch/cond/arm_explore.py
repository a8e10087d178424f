import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from drake_ddp_b200 import problems, systems
from tests.helpers import make_oracle
def run(expr, sigma, nb=6, kp="problem"):
    prob = eval(expr); prob.sigma = sigma
    x0s = prob.batch_x0(nb, seed=0)
    for b in range(nb):
        o = make_oracle(prob, x0=x0s[b], kp=kp)
        t0=time.time()
        try:
            o.solve(max_iters=60); err=None
        except RuntimeError as e: err=str(e)
        L=[r.L for r in o.trace]
        print(f"  traj {b}: it={len(L)} err={err} costs: " + " ".join(f"{l:.3f}" for l in L[:25]), " ls:", [r.ls_iters for r in o.trace][:25], f"{time.time()-t0:.1f}s")
        ball = o.x_bar[:, 11:14]
        print("     ball start", ball[0], "end", ball[-1], "target", prob.x_nom[11:14])
if __name__ == "__main__":
    run(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv)>3 else 4, None if (len(sys.argv)>4 and sys.argv[4]=="nokp") else "problem")
