"""First GPU shake-out: parity vs oracle on several problems + timings (scratch, not a test)."""
import sys, time, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from drake_ddp_b200 import problems, _lib
from drake_ddp_b200.ilqr import BatchedILQR, IterativeLinearQuadraticRegulator
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation
from oracle.dynamics import HostDynamics
from oracle.ilqr_port import IlqrOracle

print(torch.cuda.get_device_name(0))
L = _lib.lib()
for mma in (0, 1):
    tf = ctypes.c_double()
    _lib.check(L.ddp_peak_fp64(None, mma, ctypes.byref(tf)))
    print("peak fp64", "DMMA" if mma else "DFMA", tf.value, "TFLOP/s")

def make(prob, B=1, kp=None, A=None, x0=None):
    s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma,
                    derivs_keypoint_method=kp, ls_parallel=A)
    s.set_cost(prob.Q, prob.R, prob.Qf); s.set_target(prob.x_nom)
    s.set_initial_state(prob.x0 if x0 is None else x0); s.set_initial_guess(prob.u_guess)
    return s
def oracle(prob, kp=None, x0=None):
    o = IlqrOracle(HostDynamics(prob.system), prob.N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma, keypoints=kp)
    o.set_initial_state(prob.x0 if x0 is None else x0); o.set_target_state(prob.x_nom); o.set_running_cost(prob.Q, prob.R)
    o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
    return o
def rel(a, b): return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))

def compare(prob, iters, kp=None, A=None):
    try:
        s = make(prob, kp=kp, A=A); o = oracle(prob, kp=kp)
        s.begin_solve(); Lo = np.inf
        for i in range(iters):
            s.iterate(); rec = o.iterate(Lo); Lo = rec.L
            Lg = s.cost[0]; eps = s.get(_lib.EPS)[0]; ls = s.get_int(_lib.I_LS_ITERS)[0]
            kpg = s.keypoints()[0]
            print(f"  {prob.name} it{i} L gpu {Lg:.12g} cpu {rec.L:.12g} rel {abs(Lg-rec.L)/abs(rec.L):.1e} eps {eps} {rec.eps} ls {ls} {rec.ls_iters} kp_equal {kpg == rec.keypoints} nkp {len(kpg)}"
                  f" dx {rel(s.get(_lib.X_BAR)[0], o.x_bar):.1e} dK {rel(s.get(_lib.K)[0], o.K):.1e} dkappa {rel(s.get(_lib.KAPPA)[0], o.kappa):.1e} dfx {rel(s.get(_lib.FX)[0], o.fx):.1e} dfu {rel(s.get(_lib.FU)[0], o.fu):.1e} ddV {rel(s.get(_lib.DV)[0], o.dV):.1e}")
    except Exception:
        traceback.print_exc()

compare(problems.pendulum(100), 5)
compare(problems.acrobot(40), 5)
compare(problems.cart_pole(100), 4)
compare(problems.cart_pole_with_wall(100), 4)
compare(problems.affine_sin(4, 1, 40), 4)
compare(problems.affine_sin(4, 1, 40), 3, kp=derivs_interpolation('setInterval', 5, 0, 0, 0))
compare(problems.affine_sin(4, 1, 40), 3, kp=derivs_interpolation('adaptiveJerk', 2, 10, 1e-4, 0))
compare(problems.affine_sin(4, 1, 40), 3, kp=derivs_interpolation('iterativeError', 2, 0, 0, 1e-9))
compare(problems.affine_sin(6, 2, 30), 3)
compare(problems.affine_sin(27, 7, 30), 3)
compare(problems.affine_sin(37, 12, 30), 3)
compare(problems.quadruped(50), 4)
compare(problems.quadruped(50), 3, kp=derivs_interpolation('adaptiveJerk', 2, 20, 0.3, 10), A=1)
compare(problems.arm_ball(50, keypoints="setInterval5"), 3, kp=problems.arm_ball(50).keypoints)

# drop-in class
try:
    p = problems.pendulum(100)
    ilqr = IterativeLinearQuadraticRegulator(p.system, p.N)
    ilqr.SetInitialState(p.x0); ilqr.SetTargetState(p.x_nom); ilqr.SetRunningCost(p.Q, p.R); ilqr.SetTerminalCost(p.Qf); ilqr.SetInitialGuess(p.u_guess)
    x, u, t, c = ilqr.Solve(); print("class Solve:", x.shape, u.shape, t, c)
except Exception:
    traceback.print_exc()

# throughput C4
try:
    prob = problems.quadruped(200)
    for B, A in ((1024, 1), (1024, 2)):
        s = make(prob, B=B, A=A, x0=prob.batch_x0(B))
        s.begin_solve()
        for i in range(6):
            torch.cuda.synchronize(); t0 = time.time()
            na = s.iterate()
            torch.cuda.synchronize(); dt = time.time() - t0
            ls = s.get_int(_lib.I_LS_ITERS)
            print(f"C4 B={B} A={A} it{i} wall {dt*1e3:.2f} ms  {s.timings_ms()}  active {na} ls mean {ls.mean():.2f} max {ls.max()} L mean {s.cost.mean():.5f} status {np.bincount(s.status, minlength=3)}")
        del s
except Exception:
    traceback.print_exc()
