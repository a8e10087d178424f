"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): two iLQR iterations
of small quadruped (fused linearization, 8-lane rollout, symmetric backward sweep), quadruped_quat
(the same three in the n = 37 layout: odd-sized bulk-TMA tiles, single fx buffer), arm_ball (8-lane
arm rollout, generic AD + interpolation, odd-sized tiles) and pendulum (scalar kernels) problems,
with the device-side MPC re-arm on.  Driven by tests/test_gpu_parity.py::test_compute_sanitizer_clean_..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drake_ddp_b200 import problems
from drake_ddp_b200.ilqr import BatchedILQR

for name, N in (("quadruped", 12), ("quadruped_quat", 10), ("arm_ball", 10), ("pendulum", 20)):
    prob = getattr(problems, name)(N)
    B = 3
    s = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma,
                    ls_parallel=8)
    s.set_cost(prob.Q, prob.R, prob.Qf)
    s.set_target(prob.x_nom)
    s.set_initial_state(prob.batch_x0(B, seed=0))
    s.set_initial_guess(prob.u_guess)
    s.set_mpc_rearm(2)
    s.begin_solve()
    for _ in range(2):
        s.iterate()
    print(name, "cost", s.cost)
