#!/usr/bin/env python
"""Benchmark of the batched iLQR hot path (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

A *step* is one iLQR iteration (line-search rollouts + linearization + backward Riccati sweep,
/root/reference/ilqr.py:695-697) of every trajectory of the batch.  Workload: BASELINE config C4,
"1024 MPC resolves" of the quadruped (n=36, m=12, N=200), B=1024 per GPU (weak scaling;
x0 = stand pose + 0.01 N(0,I), seed = rank), run as the receding-horizon loop of
mini_cheetah.py:186-206: a trajectory whose Solve() converges (improvement <= delta, ilqr.py:692)
is shifted by replan_steps = 4, its target advances, and its next iteration is the first of the
next resolve -- on the device (ddp_set_mpc_rearm), so EVERY timed step has B active trajectories.
The CPU arm follows the same rule.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "iLQR iterations/sec (fwd+bwd) at batch=1024, horizon N=200; cost vs reference"
UNIT = "trajectory-iterations/s"
REPLAN_STEPS = 4          # mini_cheetah.py:37


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU (weak scaling)")
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--ls-parallel", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong split / other configs / convergence check")
    return ap.parse_args()


def workload(args):
    from drake_ddp_b200 import problems
    return problems.quadruped(args.horizon)


def target_advance(prob):
    """x_nom[base x] += target_vel * dt * replan_steps per resolve (mini_cheetah.py:151-156)."""
    adv = np.zeros(prob.system.n)
    adv[0 if prob.system.n == 36 else 4] = prob.extra["target_vel"] * prob.system.dt * REPLAN_STEPS
    return adv


def config_dict(args, prob, world):
    return {"workload": f"C4 quadruped (mini_cheetah-scale analytic model) n={prob.system.n} m={prob.system.m} "
                        f"N={prob.N} B={args.batch} per GPU, fp64, setInterval-1 keypoints, receding-horizon "
                        f"resolves (replan {REPLAN_STEPS} steps, moving target) re-armed on convergence: "
                        f"every step iterates all B trajectories",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "horizon": prob.N,
            "n": prob.system.n, "m": prob.system.m, "beta": prob.beta, "delta": prob.delta,
            "x0_sigma": prob.sigma, "replan_steps": REPLAN_STEPS,
            "l2": "per-step working set 7.4 GB per GPU >> 126 MB L2 (no flush needed)",
            "parallelism": f"batch-sharded x{world}"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of the reference, one trajectory per process)
# ----------------------------------------------------------------------------------------------
def _cpu_worker(conn, horizon, seeds):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    from drake_ddp_b200 import problems
    from oracle.dynamics import HostDynamics
    from oracle.ilqr_port import IlqrOracle
    prob = problems.quadruped(horizon)
    adv = target_advance(prob)
    dyn = HostDynamics(prob.system)
    solvers = []
    for seed in seeds:
        x0 = prob.batch_x0(1, seed=1000 + seed)[0]
        o = IlqrOracle(dyn, prob.N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
        o.set_initial_state(x0); o.set_target_state(prob.x_nom)
        o.set_running_cost(prob.Q, prob.R); o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
        solvers.append([o, np.inf])
    conn.send("ready")
    r = REPLAN_STEPS
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        done = 0
        for so in solvers:
            o = so[0]
            try:
                rec = o.iterate(so[1])
                so[1] = rec.L
                done += 1
                if rec.improvement <= prob.delta:     # Solve() returned: next MPC resolve (mini_cheetah.py:190-201)
                    u = o.u_bar.T
                    o.set_initial_guess(np.block([u[:, r:], np.repeat(u[:, -1][np.newaxis].T, r, axis=1)]))
                    o.set_initial_state(o.x_bar[r])
                    o.set_target_state(o.x_nom + adv)
                    so[1] = np.inf
            except RuntimeError:
                pass
        conn.send(done)


class CpuFarm:
    def __init__(self, horizon, procs, per_proc):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        for k in saved:           # one thread per worker process: the reference's execution model
            os.environ[k] = "1"
        for p in range(procs):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_cpu_worker, args=(b, horizon, list(range(p * per_proc, (p + 1) * per_proc))),
                             daemon=True)
            pr.start()
            self.conns.append(a)
            self.procs.append(pr)
        for c in self.conns:
            assert c.recv() == "ready"
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def step(self):
        for c in self.conns:
            c.send("iter")
        return sum(c.recv() for c in self.conns)

    def close(self):
        for c in self.conns:
            c.send("stop")
        for p in self.procs:
            p.join(timeout=10)


def cpu_measure(horizon, steps, warmup, per_proc=1):
    procs = max(1, min(os.cpu_count() or 1, 64))
    farm = CpuFarm(horizon, procs, per_proc)
    for _ in range(warmup):
        farm.step()
    t0 = time.perf_counter()
    units = 0
    for _ in range(steps):
        units += farm.step()
    dt = time.perf_counter() - t0
    farm.close()
    return units / dt, procs, dt, units


def cpu_single_thread(horizon, iters=6):
    """The reference's own execution model: ONE trajectory, one thread, phases timed
    the way ilqr.py:364-372,696-699 times them (forward pass = line search + derivatives)."""
    from drake_ddp_b200 import problems
    from oracle.dynamics import HostDynamics
    from oracle.ilqr_port import IlqrOracle
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)
    except Exception:
        limit = None
    prob = problems.quadruped(horizon)
    o = IlqrOracle(HostDynamics(prob.system), prob.N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
    o.set_initial_state(prob.batch_x0(1, seed=999)[0]); o.set_target_state(prob.x_nom)
    o.set_running_cost(prob.Q, prob.R); o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
    L = o.iterate(np.inf).L                       # warm-up iteration
    ph = {"linesearch": 0.0, "derivs": 0.0, "backward": 0.0}
    t_all = time.perf_counter()
    for _ in range(iters):
        t0 = time.perf_counter()
        eps, x, u, Lc, n_ls = o.linesearch(L)
        t1 = time.perf_counter()
        o.get_derivatives(x, u)
        o.u_bar, o.x_bar = u, x
        t2 = time.perf_counter()
        o.backward_pass()
        t3 = time.perf_counter()
        ph["linesearch"] += t1 - t0; ph["derivs"] += t2 - t1; ph["backward"] += t3 - t2
        L = float(Lc)
    dt = time.perf_counter() - t_all
    if limit is not None:
        limit.restore_original_limits()
    return {"value": iters / dt, "unit": UNIT, "cores": 1, "iterations": iters,
            "phase_ms_per_iteration": {k: v / iters * 1e3 for k, v in ph.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = workload(args)
    # 8 trajectories per process: the processes meet after every iteration, and with one trajectory
    # each the slowest line search of the step sets the pace (measured: 449/s against 565/s); the
    # same sample shape as the in-line cpu_baseline of the B200 arm
    per_proc = 8
    value, procs, dt, units = cpu_measure(args.horizon, args.steps, args.warmup, per_proc=per_proc)
    sample = (f"{per_proc * procs} trajectories ({per_proc} per process, {procs} processes, OMP_NUM_THREADS=1) x "
              f"{args.steps} iLQR iterations of the C4 problem after {args.warmup} warm-up iterations, same "
              f"re-arm-on-convergence rule; oracle port = numpy restatement of ilqr.py + host build of the analytic model")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, prob, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                             "single_thread": cpu_single_thread(args.horizon)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        """windows: list of (t0, t1) host-time intervals of the timed regions."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                smax = float(parts[1])
                if any(t0 <= ts <= t1 for t0, t1 in windows):
                    sm.append(float(parts[0]))
                    for nm, val in zip(names, parts[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def phase_bytes(n, m, N):
    """Algorithmic bytes per trajectory (SURVEY.md 8d / DESIGN.md), fp64."""
    T = N - 1
    return {"backward": 8 * ((n * n + n * m) * T + (n * N + m * T) + (m * n + m + 1) * T),
            "derivs": 8 * ((n * N + m * T) + (n * n + n * m) * T),
            "rollout": 8 * ((n * N + m * T + m * T + m * n * T + T) + (n * N + m * T))}


def make_solver(prob, B, ls_parallel=None, rearm=True, kp="problem"):
    from drake_ddp_b200.ilqr import BatchedILQR
    solver = BatchedILQR(prob.system, prob.N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma,
                         ls_parallel=ls_parallel, derivs_keypoint_method=prob.keypoints if kp == "problem" else kp)
    solver.set_cost(prob.Q, prob.R, prob.Qf)
    solver.set_target(prob.x_nom)
    if rearm:
        solver.set_mpc_rearm(REPLAN_STEPS, target_advance(prob))
    return solver


def timed_iterations(torch, solver, W, K, barrier, after_step=None):
    """W warm-up + K timed ddp_iterate calls; CUDA events on the solver's stream.  Returns
    (elapsed ms, trajectory-iterations done in the timed region, per-phase ms sums, host window)."""
    from drake_ddp_b200 import _lib
    for _ in range(W):
        solver.iterate()
        if after_step:
            after_step()
    it0 = solver.get_int(_lib.I_ITERS).astype(np.int64)
    phase_ms = {"linesearch": 0.0, "derivs": 0.0, "backward": 0.0}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(solver._stream)
    for _ in range(K):
        solver.iterate()
        if after_step:
            after_step()
        ms = solver.timings_ms()          # device events recorded inside the library, no extra sync
        for k in phase_ms:
            phase_ms[k] += ms[k]
    e1.record(solver._stream)
    barrier()
    t1 = time.perf_counter()
    units = int((solver.get_int(_lib.I_ITERS).astype(np.int64) - it0).sum())
    return e0.elapsed_time(e1), units, phase_ms, (t0, t1)


KERNELS = {   # (line search, derivatives, backward sweep) kernels per model
    "quadruped": ("rollout_quad8_kernel", "quad_fused_kernel", "backward_sym_kernel"),
    "quadruped_quat": ("rollout_quad8_kernel<QUAT>", "quad_quat_fused_kernel", "backward_sym_kernel"),
    "arm_ball": ("rollout_arm8_kernel", "linearize_kernel + interp_kernel", "backward_sym_kernel"),
}


def dominant(phase_ms, pb, active_per_step, K, ls_mean, hbm_peak, model="quadruped"):
    dom = max(phase_ms, key=phase_ms.get)
    launch_ms = phase_ms[dom] / K
    names = KERNELS.get(model, ("rollout_kernel", "linearize_kernel", "backward_kernel"))
    if dom == "backward":
        kernel, nbytes = names[2], pb["backward"] * active_per_step
    elif dom == "derivs":
        kernel, nbytes = names[1], pb["derivs"] * active_per_step
    else:
        kernel = names[0] + " (line-search phase, all rounds)"
        nbytes = pb["rollout"] * active_per_step * ls_mean
    ach = nbytes / (launch_ms * 1e-3) / 1e9
    return {"kernel": kernel, "achieved": ach, "frac": ach / hbm_peak, "launch_ms": launch_ms,
            "algorithmic_bytes_per_launch": nbytes}


def solve_batch_throughput(torch, prob, B, x0, A=None, max_iters=40, hbm_peak=6546.2):
    """Other configs: a fresh batch solved until every trajectory has converged or failed its line
    search (or max_iters); value = trajectory-iterations / device time of the whole batch solve."""
    from drake_ddp_b200 import _lib
    s = make_solver(prob, B, ls_parallel=A, rearm=False)
    n, m, N = prob.system.n, prob.system.m, prob.N
    u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, N - 1, m)))

    def run(record):
        s.reset(); s.set_initial_state(x0); s.set_initial_guess(u0); s.begin_solve()
        ph = {"linesearch": 0.0, "derivs": 0.0, "backward": 0.0}
        full = {"linesearch": 0.0, "derivs": 0.0, "backward": 0.0, "iterations": 0}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s._stream)
        it, n_active = 0, B
        while n_active > 0 and it < max_iters:
            was_full = n_active == B
            n_active = s.iterate()
            it += 1
            if record:
                ms = s.timings_ms()
                for k in ph:
                    ph[k] += ms[k]
                    if was_full:
                        full[k] += ms[k]
                full["iterations"] += int(was_full)
        e1.record(s._stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), it, ph, full

    run(False)                                   # warm-up solve (same work: the solve is deterministic)
    ms, iters, ph, full = run(True)
    units = int(s.get_int(_lib.I_ITERS).sum())
    status = s.status
    ls_mean = float(np.mean(s.get_int(_lib.I_LS_ITERS)))
    pb = phase_bytes(n, m, N)
    dom = dominant(ph, pb, units / iters, iters, ls_mean, hbm_peak, model=prob.system.name)
    out = {"n": n, "m": m, "N": N, "B": B, "ls_parallel": s.A, "value": units / (ms * 1e-3), "unit": UNIT,
           "batch_iterations": iters, "ms_per_batch_iteration": ms / iters,
           "mean_active_per_step": units / iters,
           "status": {"converged": int((status == 1).sum()), "linesearch_failed": int((status == 2).sum()),
                      "running": int((status == 0).sum())},
           "phase_ms_per_step": {k: v / iters for k, v in ph.items()},
           "dominant": {"kernel": dom["kernel"], "hbm_frac": dom["frac"], "launch_ms": dom["launch_ms"]}}
    # the iterations during which every trajectory of the batch was still iterating: the number to
    # compare with the headline config (phase times summed; no host gaps)
    kf = full.pop("iterations")
    if kf:
        tot = sum(full.values()) / kf
        out["all_active"] = {"iterations": kf, "phase_ms_per_step": {k: v / kf for k, v in full.items()},
                             "ms_per_batch_iteration": tot, "value": B / (tot * 1e-3), "unit": UNIT}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from drake_ddp_b200 import _lib, problems
    from drake_ddp_b200.dist import CostGather

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    # stdout carries ONE JSON line: whatever libraries print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) is sent to stderr; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION|INFO) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # the path's one collective is an 8 KB all-gather: one channel is plenty, and every NCCL CTA
        # takes residency from the rollout (1024 CTAs on 1036 slots) while it spins for its peer
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sampler = ClockSampler(local)              # started before any warm-up
    prob = workload(args)
    B, n, m, N, T = args.batch, prob.system.n, prob.system.m, prob.N, prob.N - 1
    K, W = args.steps, args.warmup
    x0 = prob.batch_x0(B, seed=rank)
    solver = make_solver(prob, B, args.ls_parallel)
    u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, T, m)))
    gather = CostGather(solver, B * world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_costs():
        # the path's one collective: per-trajectory costs of all ranks (issued on a side stream,
        # overlaps the next iteration; waited for before the next gather)
        if gather is not None:
            gather.issue()

    def fresh(s=solver, x=x0, u=u0):
        s.reset()
        s.set_initial_state(x)
        s.set_initial_guess(u)
        s.begin_solve()

    # ------------------------------------------------------------------ device-resident timing
    fresh()
    launches0 = solver.launch_count()
    elapsed_ms, units_local, phase_ms, win_dev = timed_iterations(torch, solver, W, K, barrier, gather_costs)
    launches = solver.launch_count() - launches0
    ls = solver.get_int(_lib.I_LS_ITERS)
    status = solver.get_int(_lib.I_STATUS)
    resolves = solver.get_int(_lib.I_RESOLVES)

    # ------------------------------------------------------------------ end-to-end timing
    # Public API with HOST buffers: every step uploads x0 and the control tape from pinned memory,
    # runs one iteration, and reads the new control tape, x0 and the costs back; what is read back
    # is what the next step uploads (closed loop through host memory, like the MPC loops of the
    # reference's scripts).  HostExchange overlaps the copies with the derivatives + backward pass
    # where the data is already final.  A trajectory that the device re-arms at the end of an
    # iteration keeps its own (shifted) tape over the stale uploaded row (ddp_apply_staged_inputs).
    x0_pin = torch.from_numpy(x0.copy()).pin_memory()
    u_pin = torch.from_numpy(u0.copy()).pin_memory()
    cost_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    fresh()
    ex = solver.host_exchange()

    def e2e_step():
        ex.apply_inputs()
        solver.iterate_linesearch()
        ex.read_controls(u_pin)               # tape of this iteration, under derivatives + backward
        solver.iterate_finish_async()
        ex.wait_controls()
        ex.stage_inputs(x0_pin, u_pin)        # next upload travels under the backward pass
        n_act = solver.iterate_wait()
        gather_costs()
        ex.read_state(x0_pin, cost_pin)       # costs + x0 (the device moves x0 of re-armed trajectories)
        return n_act

    ex.stage_inputs(x0_pin, u_pin)
    for _ in range(W):
        e2e_step()
    it0e = solver.get_int(_lib.I_ITERS).astype(np.int64)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e0 = time.perf_counter()
    f0.record(solver._stream)
    for _ in range(K):
        e2e_step()
    f1.record(solver._stream)
    barrier()
    t_e1 = time.perf_counter()
    e2e_ms = f0.elapsed_time(f1)
    units_e2e_local = int((solver.get_int(_lib.I_ITERS).astype(np.int64) - it0e).sum())
    clocks = sampler.stop([win_dev, (t_e0, t_e1)])

    # ------------------------------------------------------------------ strong-scaling split of C4
    # BASELINE config 4: batch = 1024 sharded over the GPUs (1024 / world per GPU)
    strong = None
    if not args.no_extras:
        Bs = max(1, 1024 // world)
        if world == 1 and B == 1024:
            s_ms, s_units, s_phase = elapsed_ms, units_local, phase_ms
        else:
            ss = make_solver(prob, Bs, args.ls_parallel)
            xs = prob.batch_x0(1024, seed=0)[rank * Bs:(rank + 1) * Bs]
            us = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (Bs, T, m)))
            fresh(ss, xs, us)
            sg = CostGather(ss, Bs * world) if world > 1 else None
            s_ms, s_units, s_phase, _ = timed_iterations(torch, ss, W, K, barrier, (sg.issue if sg else None))
            del ss
        strong = (s_ms, s_units, s_phase, Bs)

    # ------------------------------------------------------------------ reduce over ranks
    vals = [elapsed_ms, e2e_ms, float(units_local), float(units_e2e_local), float(launches)]
    if strong:
        vals += [strong[0], float(strong[1])]
    stats = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    elapsed_ms, e2e_ms = float(mx[0]), float(mx[1])
    units, units_e2e, launches_all = float(sm[2]), float(sm[3]), int(sm[4])

    if rank == 0:
        value = units / (elapsed_ms * 1e-3)
        e2e_value = units_e2e / (e2e_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        pb = phase_bytes(n, m, N)
        active_per_step = units_local / K
        dom = dominant(phase_ms, pb, active_per_step, K, float(np.mean(ls)), hbm_peak)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom["kernel"].split()[0])
        except Exception:
            pass
        tf = {}
        import ctypes
        for mma, nm in ((0, "dfma"), (1, "dmma")):
            v = ctypes.c_double()
            if _lib.lib().ddp_peak_fp64(None, mma, ctypes.byref(v)) == 0:
                tf[nm] = round(v.value, 2)
        bwd_flops = 2.0 * T * (2 * n ** 3 + 3 * n * n * m + 2 * n * m * m) * active_per_step
        whole = pb["backward"] + pb["derivs"] + pb["rollout"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, prob, world),
            "active_per_step": active_per_step,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": hbm_peak,
                         "unit": "GB/s", "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                         "launch_ms": dom["launch_ms"],
                         "whole_iteration": {
                             "algorithmic_bytes_per_trajectory_iteration": whole,
                             "achieved_GBps": whole * value / world / 1e9,
                             "frac": whole * value / world / 1e9 / hbm_peak},
                         "fp64": {"measured_peak_tflops": tf,
                                  "backward_tflops": bwd_flops / (phase_ms["backward"] / K * 1e-3) / 1e12,
                                  # DMMAs the sweep issues per step at (36, 12), DESIGN.md section 3;
                                  # 512 flop each, against the measured DMMA peak
                                  "backward_dmma_issue_frac": (
                                      _lib.BWD_DMMA_PER_STEP_36_12 * 512.0 * T * active_per_step
                                      / (phase_ms["backward"] / K * 1e-3) / 1e12 / tf["dmma"])
                                  if (n, m) == (36, 12) and tf.get("dmma") else None}},
            "phase_ms_per_step": {k: v / K for k, v in phase_ms.items()},
            "ls_iters_mean_last_step": float(np.mean(ls)), "ls_parallel": solver.A,
            "trajectory_status": {"running": int((status == 0).sum()), "converged": int((status == 1).sum()),
                                  "linesearch_failed": int((status == 2).sum()),
                                  "resolves_completed": int(resolves.sum())},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(B * (n + T * m) * 8),
                    "d2h_bytes_per_step": int(ex.d2h_bytes / max(1, ex.d2h_steps)),
                    "ms_per_step": e2e_ms / K, "active_per_step": units_e2e_local / K},
            "gpu_launches": launches_all, "clocks": clocks,
        }
        if strong:
            s_ms, s_units = float(mx[5]), float(sm[6])
            sphase = strong[2]
            line["strong"] = {"global_batch": strong[3] * world, "per_gpu": strong[3],
                              "value": s_units / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / K,
                              "phase_ms_per_step_rank0": {k: v / K for k, v in sphase.items()},
                              "limiter": "per-step latency of the 199-step sequential kernels: below one "
                                         "CTA per SM slot the rollout and backward sweeps take the same time "
                                         "for fewer trajectories (DESIGN.md section 6)"}
        if world == 1 and not args.no_extras:
            line["other_configs"] = other_configs(torch, problems, hbm_peak)
            line["cost_vs_oracle"] = converged_cost_vs_oracle(prob, args)
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload on the host cores (about 10-30 s of CPU work)
            cpu_steps, cpu_per_proc = 12, 8
            cv, procs, cdt, cunits = cpu_measure(args.horizon, cpu_steps, 1, per_proc=cpu_per_proc)
            line["cpu_baseline"] = {
                "value": cv, "unit": UNIT, "cores": procs, "kind": "port",
                "sample": f"{cpu_per_proc * procs} trajectories ({cpu_per_proc} per process, {procs} processes, 1 thread each) x {cpu_steps} "
                          f"iLQR iterations of the same C4 problem after 1 warm-up iteration, same re-arm rule "
                          f"({cunits} trajectory-iterations in {cdt:.1f} s wall)",
                "single_thread": cpu_single_thread(args.horizon)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(torch, problems, hbm_peak):
    """The other BASELINE configs as batch-solve throughput (parity for each is in tests/)."""
    out = {}
    cases = {
        "C5_arm_ball_n27_N400_B512_setInterval5": (lambda: problems.arm_ball(400), 512, None),
        "C4_n37_quat_N200_B1024": (lambda: problems.quadruped_quat(200), 1024, None),
        "C2_acrobot_N40_B50": (lambda: problems.acrobot(40), 50, 8),
        "C3_wall_N200_256_alphas": (lambda: problems.cart_pole_with_wall(200, beta=0.95), 1, 256),
        "C1_pendulum_N100_B1": (lambda: problems.pendulum(100), 1, 8),
    }
    for name, (factory, B, A) in cases.items():
        try:
            prob = factory()
            x0 = prob.batch_x0(B, seed=0) if B > 1 else prob.x0[None].copy()
            out[name] = solve_batch_throughput(torch, prob, B, x0, A=A, hbm_peak=hbm_peak)
        except Exception as e:      # a failing side config must not lose the headline line
            out[name] = {"error": repr(e)}
        torch.cuda.empty_cache()
    return out


def converged_cost_vs_oracle(prob, args, nb=8):
    """Cost vs reference at CONVERGENCE: nb seeds of the workload solved to convergence by a fresh
    solver on the GPU and by the CPU oracle (ilqr.py:669-710); worst relative difference."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    from drake_ddp_b200 import _lib
    x0 = prob.batch_x0(nb, seed=0)
    s = make_solver(prob, nb, args.ls_parallel, rearm=False)
    s.set_initial_state(x0)
    s.set_initial_guess(prob.u_guess)
    s.begin_solve()
    it = 0
    while s.iterate() > 0 and it < 200:
        it += 1
    cost, iters = s.cost, s.get_int(_lib.I_ITERS)
    with ProcessPoolExecutor(max_workers=min(nb, os.cpu_count() or 1), mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_oracle_solve, [(args.horizon, x0[b]) for b in range(nb)]))
    rel = [abs(cost[b] - res[b][0]) / abs(res[b][0]) for b in range(nb)]
    return {"trajectories": nb, "at": "convergence (improvement <= delta, ilqr.py:692)",
            "gpu": [float(c) for c in cost], "oracle": [r[0] for r in res],
            "iterations_gpu": [int(i) for i in iters], "iterations_oracle": [r[1] for r in res],
            "rel_err": float(max(rel)), "rel_err_each": [float(r) for r in rel]}


def _oracle_solve(a):
    horizon, x0 = a
    os.environ["OMP_NUM_THREADS"] = "1"
    from drake_ddp_b200 import problems
    from oracle.dynamics import HostDynamics
    from oracle.ilqr_port import IlqrOracle
    prob = problems.quadruped(horizon)
    o = IlqrOracle(HostDynamics(prob.system), prob.N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
    o.set_initial_state(x0); o.set_target_state(prob.x_nom)
    o.set_running_cost(prob.Q, prob.R); o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
    o.solve(max_iters=200)
    return float(o.trace[-1].L), len(o.trace)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
