"""Debug aid: fused vs AD linearization of quadruped_quat, error by row/column block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import _lib, problems
from tests.helpers import make_gpu
prob = problems.quadruped_quat(20)
B = 4
rng = np.random.default_rng(7)
x = prob.x0[None, None] + 0.05 * rng.standard_normal((B, prob.N, 37))
x[:, :, 0:4] += 0.1 * rng.standard_normal((B, prob.N, 4))
x[:, :, 19:22] += 0.5 * rng.standard_normal((B, prob.N, 3))
x[B // 2:, :, 6] += 0.05
u = prob.u_guess.T[None] + 2.0 * rng.standard_normal((B, prob.N - 1, 12))
out = {}
for mode in ("ad", "fused"):
    os.environ["DDP_QUAD_LINEARIZE"] = mode
    s = make_gpu(prob, B=B)
    s.put(_lib.X_BAR, x); s.put(_lib.U_BAR, u)
    s.run_phase(_lib.PHASE_DERIVATIVES)
    out[mode] = np.concatenate([s.get(_lib.FX), s.get(_lib.FU)], axis=-1)
a, f = out["ad"], out["fused"]
print("max abs ad", np.abs(a).max(), "max abs diff", np.abs(a - f).max())
rows = {"qt": slice(0, 4), "pos": slice(4, 7), "qj": slice(7, 19), "w": slice(19, 22), "vl": slice(22, 25), "vj": slice(25, 37)}
cols = dict(rows); cols["u"] = slice(37, 49)
for b in (0, B - 1):
    print("trajectory", b, "(contact)" if b == 0 else "(flight)")
    for rn, rs in rows.items():
        print("  rows %-3s" % rn, {cn: float("%.1e" % np.abs(a[b, :, rs, cs] - f[b, :, rs, cs]).max()) for cn, cs in cols.items()})
