"""Host-side time split of bench.py's end-to-end step (scratch tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from drake_ddp_b200 import _lib, problems
prob = problems.quadruped(200)
B, T, m, n = 1024, 199, 12, 36
torch.cuda.set_device(0)
solver = bench.make_solver(prob, B)
x0 = prob.batch_x0(B, seed=0)
u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, T, m)))
solver.reset(); solver.set_initial_state(x0); solver.set_initial_guess(u0); solver.begin_solve()
x0_pin = torch.from_numpy(x0.copy()).pin_memory(); u_pin = torch.from_numpy(u0.copy()).pin_memory()
cost_pin = torch.empty(B, dtype=torch.float64).pin_memory()
ex = solver.host_exchange()
ex.stage_inputs(x0_pin, u_pin)
acc = {}
def tick(name, t0):
    torch.cuda.synchronize() if False else None
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
for it in range(25):
    if it == 5: acc.clear(); tstart = time.perf_counter()
    t = time.perf_counter()
    ex.apply_inputs(); t = tick("apply", t)
    solver.iterate_linesearch(); t = tick("linesearch(sync)", t)
    ex.read_controls(u_pin); t = tick("read_controls(issue)", t)
    solver.iterate_finish_async(); t = tick("finish_async(issue)", t)
    ex.wait_controls(); t = tick("wait_controls", t)
    ex.stage_inputs(x0_pin, u_pin); t = tick("stage(issue)", t)
    solver.iterate_wait(); t = tick("iterate_wait", t)
    ex.read_state(x0_pin, cost_pin); t = tick("read_state", t)
tot = time.perf_counter() - tstart
print("ms per step:", tot / 20 * 1e3)
for k, v in acc.items(): print(f"  {k:24s} {v / 20 * 1e3:.3f} ms")
