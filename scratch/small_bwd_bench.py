"""Scalar backward sweeps of the small models: register kernel vs CTA kernel, per launch (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.ilqr import BatchedILQR
for name, N, B in (("acrobot", 40, 50), ("acrobot", 400, 50), ("pendulum", 100, 1), ("cart_pole_with_wall", 200, 1)):
    for mode in ("reg", "cta"):
        os.environ.pop("DDP_SMALL_BACKWARD", None)
        if mode == "cta":
            os.environ["DDP_SMALL_BACKWARD"] = "cta"
        prob = getattr(problems, name)(N)
        s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
        s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
        s.set_initial_state(prob.batch_x0(B, seed=0) if B > 1 else prob.x0[None].copy()); s.set_initial_guess(prob.u_guess)
        s.begin_solve(); s.iterate(); s.iterate()
        ts = []
        for r in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s._stream); s.run_phase(_lib.PHASE_BACKWARD); e1.record(s._stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(name, N, B, mode, "us", round(1e3 * float(np.median(ts)), 1))
