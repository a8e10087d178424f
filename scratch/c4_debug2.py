import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drake_ddp_b200 import _lib, problems
from tests.helpers import make_gpu, make_oracle, relerr
prob = problems.quadruped(200)
B = 1024
x0 = prob.batch_x0(B, seed=0)
s = make_gpu(prob, B=B, A=2, x0=x0)
s.begin_solve(); s.iterate()
K = s.get(_lib.K); fx, fu, xb, ub = s.get(_lib.FX), s.get(_lib.FU), s.get(_lib.X_BAR), s.get(_lib.U_BAR)
print("cost", s.cost[[0,1,517]])
for b in (0, 517):
    o = make_oracle(prob, x0=x0[b])
    o.fx, o.fu, o.x_bar, o.u_bar = fx[b].copy(), fu[b].copy(), xb[b].copy(), ub[b].copy()
    # instrumented backward pass
    Q, R, Qf, xn = o.Q, o.R, o.Qf, o.x_nom
    Vx = 2 * Qf @ o.x_bar[-1] - 2 * xn.T @ Qf; Vxx = 2 * Qf
    print("b", b, "max|x|", abs(xb[b]).max(), "max|fx|", abs(fx[b]).max())
    for t in range(o.N - 2, -1, -1):
        f_x, f_u = o.fx[t], o.fu[t]
        Quu = 2 * R + f_u.T @ Vxx @ f_u
        Qux = f_u.T @ Vxx @ f_x
        Quu_inv = np.linalg.inv(Quu)
        Kt = Quu_inv @ Qux
        Qu = 2 * R @ o.u_bar[t] + f_u.T @ Vx
        Qx = 2 * Q @ o.x_bar[t] - 2 * xn.T @ Q + f_x.T @ Vx
        Vx = Qx - Qu.T @ Quu_inv @ Qux
        Vxx = 2 * Q + f_x.T @ Vxx @ f_x - Qux.T @ Quu_inv @ Qux
        if t % 20 == 0 or t > 190:
            ev = np.linalg.eigvalsh((Vxx + Vxx.T) / 2)
            print(f"  t {t:3d} relK {relerr(K[b, t], Kt):.2e} cond(Quu) {np.linalg.cond(Quu):.2e} max|Vxx| {abs(Vxx).max():.2e} asym {abs(Vxx - Vxx.T).max():.2e} mineig {ev.min():.2e} max|K| {abs(Kt).max():.2e}")
