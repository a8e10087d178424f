"""Print the in-kernel phase profile of backward_sym_kernel (build with -DDDP_BWD_PROFILE into
scratch/out/lib_prof.so):  DDP_B200_LIB=scratch/out/lib_prof.so PB_BATCH=1024 python scratch/bwd_prof.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drake_ddp_b200 import _lib, problems
from drake_ddp_b200.ilqr import BatchedILQR
B = int(os.environ.get("PB_BATCH", "1024"))
prob = problems.quadruped(200)
s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
s.set_initial_state(prob.batch_x0(B, seed=0)); s.set_initial_guess(prob.u_guess)
s.begin_solve(); s.iterate(); s.iterate()
s.run_phase(_lib.PHASE_BACKWARD)
out = (ctypes.c_longlong * 128)()
L = _lib.lib(); L.ddp_debug_bwd_profile.argtypes = [ctypes.c_void_p]
assert L.ddp_debug_bwd_profile(out) == 0
a = np.array(list(out)).reshape(2, 4, 16) / 199.0
names_d = ["loop", "tma-wait", "B1", "phase1", "2a+arrive", "2b-k", "barV", "epilogue", "barQ", "phaseC"]
names_v = ["loop", "tma-wait", "B1", "lx+dots", "bar2(Quu)", "inverse", "barQ", "barS", "tail"]
for cta in range(2):
    if a[cta].sum() == 0: continue
    print(f"CTA {'5 (first wave)' if cta == 0 else '700 (second wave)'}: cycles per step")
    for role in range(4):
        nm = names_v if role == 3 else names_d
        print(f"  {'vector' if role == 3 else 'dmma%d' % role}: total {a[cta, role].sum():7.0f} | " +
              " ".join(f"{nm[i]}={a[cta, role, i]:.0f}" for i in range(len(nm))))
