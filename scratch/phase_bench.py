"""Time the three phases of one iLQR iteration on the C4 problem for one library build
(scratch tool, not a test):  DDP_B200_LIB=<lib.so> python scratch/phase_bench.py [phase ...]

After two full iterations (so K, kappa, fx, fu hold realistic data) every phase is run alone
`reps` times with CUDA events around it; a checksum of its outputs is printed so variants can
be compared for equality across builds."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.ilqr import BatchedILQR

B = int(os.environ.get("PB_BATCH", "1024"))
N = int(os.environ.get("PB_HORIZON", "200"))
reps = int(os.environ.get("PB_REPS", "5"))
model = os.environ.get("PB_MODEL", "quadruped")
phases = sys.argv[1:] or ["linesearch", "derivs", "backward"]
PH = {"linesearch": _lib.PHASE_LINESEARCH, "derivs": _lib.PHASE_DERIVATIVES, "backward": _lib.PHASE_BACKWARD}

prob = getattr(problems, model)(N)
x0 = prob.batch_x0(B, seed=0)
s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
s.set_cost(prob.Q, prob.R, prob.Qf)
s.set_target(prob.x_nom)
s.set_initial_state(x0)
s.set_initial_guess(prob.u_guess)
s.begin_solve()
s.iterate()
if os.environ.get("PB_DUMP"):
    K1 = s.get(_lib.K)
    np.savez(os.environ["PB_DUMP"], K8=K1[:8], Ksum=np.abs(K1).sum(axis=(1, 2, 3)), kappa=s.get(_lib.KAPPA),
             dV=s.get(_lib.DV), cost=s.cost)
    del K1
s.iterate()
torch.cuda.synchronize()
out = {"lib": os.path.basename(_lib.LIB_PATH)}
for name in phases:
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s._stream)
        s.run_phase(PH[name])
        e1.record(s._stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[name] = round(float(np.median(ts)), 4)
    if name == "backward":
        K = s.get(_lib.K)
        out["K_sum"] = float(np.abs(K).sum())
        out["kappa_sum"] = float(np.abs(s.get(_lib.KAPPA)).sum())
        out["dV_sum"] = float(np.abs(s.get(_lib.DV)).sum())
    if name == "derivs":
        out["fx_sum"] = float(np.abs(s.get(_lib.FX)).sum())
        out["fu_sum"] = float(np.abs(s.get(_lib.FU)).sum())
print(out)
if os.environ.get("PB_PROF"):
    import ctypes
    buf = (ctypes.c_longlong * 128)()
    _lib.lib().ddp_debug_bwd_profile(buf)
    print('gauss-jordan fallbacks (all launches):', buf[127], 'of', B * (N - 1), 'per launch; newton passes (all launches):', buf[126])
    a = np.array(buf[:], dtype=np.int64).reshape(2, 4, 16)[:, :, :11]
    # prof_acc[i] = cycles between the previous tick and BS_TICK(i) (backward_sym.cuh)
    dn = ["loop", "wait fx/fu", "bar0", "W other strips", "W fu strips + Quu tiles", "M tiles (2b)", "wait barV",
          "route stores", "wait barQ", "K + update"]
    vn = ["loop", "wait fx/fu", "bar0", "lx lu", "wait Quu (bar2)", "inverse", "wait barQ", "Qx Qu + wait barS",
          "kappa dV Vx"]
    for cta in range(2):
        for w in range(4):
            nm_ = dn if w < 3 else vn
            tot = a[cta, w].sum()
            print("cta", cta, f"dmma role {w}" if w < 3 else "vector warp", "cycles/step", int(tot / (N - 1)),
                  {nm: int(v / (N - 1)) for nm, v in zip(nm_, a[cta, w])})

if os.environ.get("PB_ROLLPROF"):
    import ctypes
    buf = (ctypes.c_longlong * 16)()
    _lib.lib().ddp_debug_roll_profile(buf)
    names = ["loop", "stage-wait", "feedback", "cost", "sincos", "leg", "butterfly", "base_acc", "integrate", "tail"]
    tot = sum(buf[:10])
    print("rollout cycles/step", int(tot / (N - 1)), {nm: int(v / (N - 1)) for nm, v in zip(names, buf[:10])})
