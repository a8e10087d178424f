#!/usr/bin/env python
"""Benchmark of the batched iLQR hot path (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

A *step* is one iLQR iteration (line-search rollouts + linearization + backward Riccati
sweep, /root/reference/ilqr.py:695-697) over one batch of synthetic problems: BASELINE config
C4, the quadruped n=36, m=12, N=200, B=1024 per GPU (weak scaling; x0 = stand pose + 0.01 N(0,I),
seed = rank).  Prints ONE JSON line on rank 0.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU")
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--ls-parallel", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    from drake_ddp_b200 import problems
    return problems.quadruped(args.horizon)


def config_dict(args, prob, world):
    return {"workload": f"C4 quadruped (mini_cheetah-scale analytic model) n={prob.system.n} m={prob.system.m} "
                        f"N={prob.N} B={args.batch} per GPU, fp64, setInterval-1 keypoints",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "horizon": prob.N,
            "n": prob.system.n, "m": prob.system.m, "beta": prob.beta, "delta": prob.delta,
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
    dyn = HostDynamics(prob.system)
    solvers = []
    for seed in seeds:
        x0 = prob.batch_x0(1, seed=1000 + seed)[0]
        o = IlqrOracle(dyn, prob.N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
        o.set_initial_state(x0); o.set_target_state(prob.x_nom)
        o.set_running_cost(prob.Q, prob.R); o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
        solvers.append([o, np.inf])
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        done = 0
        for so in solvers:
            try:
                rec = so[0].iterate(so[1])
                so[1] = rec.L
                done += 1
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = workload(args)
    value, procs, dt, units = cpu_measure(args.horizon, args.steps, args.warmup, per_proc=1)
    sample = (f"{procs} trajectories (one per process, OMP_NUM_THREADS=1) x {args.steps} iLQR iterations of the "
              f"C4 problem; oracle port = numpy restatement of ilqr.py + host build of the analytic model")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, prob, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                smax = float(parts[1])
                if t0 <= ts <= t1:
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


def run_b200(args):
    import torch
    import torch.distributed as dist

    from drake_ddp_b200 import _lib
    from drake_ddp_b200.dist import all_gather_ragged
    from drake_ddp_b200.ilqr import BatchedILQR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    prob = workload(args)
    B, n, m, N, T = args.batch, prob.system.n, prob.system.m, prob.N, prob.N - 1
    K, W = args.steps, args.warmup
    x0 = prob.batch_x0(B, seed=rank)
    solver = BatchedILQR(prob.system, N, batch=B, delta=prob.delta, beta=prob.beta, gamma=prob.gamma,
                         ls_parallel=args.ls_parallel)
    solver.set_cost(prob.Q, prob.R, prob.Qf)
    solver.set_target(prob.x_nom)
    u0 = np.ascontiguousarray(np.broadcast_to(prob.u_guess.T, (B, T, m)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_costs():
        if world > 1:
            return all_gather_ragged(solver.device_tensor(_lib.COST), B * world)
        return solver.device_tensor(_lib.COST)

    def fresh():
        solver.reset()
        solver.set_initial_state(x0)
        solver.set_initial_guess(u0)
        solver.begin_solve()

    # ------------------------------------------------------------------ device-resident timing
    fresh()
    cost_after = {}
    for w in range(W):
        solver.iterate()
        gather_costs()
        cost_after[w + 1] = float(solver.cost[0])
    it0 = solver.get_int(_lib.I_ITERS).astype(np.int64)
    launches0 = solver.launch_count()
    sampler = ClockSampler(local)
    phase_ms = {"linesearch": 0.0, "derivs": 0.0, "backward": 0.0}
    ls_sum, ls_cnt = 0.0, 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record(solver._stream)
    for _ in range(K):
        solver.iterate()
        gather_costs()
        ms = solver.timings_ms()          # device events recorded inside the library, no extra sync
        for k in phase_ms:
            phase_ms[k] += ms[k]
    e1.record(solver._stream)
    barrier()
    t_host1 = time.perf_counter()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_host0, t_host1)
    launches = solver.launch_count() - launches0
    it1 = solver.get_int(_lib.I_ITERS).astype(np.int64)
    units_local = int((it1 - it0).sum())
    ls = solver.get_int(_lib.I_LS_ITERS)
    status = solver.get_int(_lib.I_STATUS)
    cost_dev = solver.cost.copy()

    # ------------------------------------------------------------------ end-to-end timing
    # Public API with HOST buffers: every step uploads x0 and the control tape from pinned
    # memory, runs one iteration, and reads the costs and the new control tape back; the tape
    # read back is the one uploaded for the next step (closed loop through host memory, like
    # the MPC loops of the reference's scripts).  HostExchange overlaps the copies with the
    # derivatives + backward pass; the bytes moved per step are unchanged.
    x0_pin = torch.from_numpy(x0).pin_memory()
    u_pin = torch.from_numpy(u0.copy()).pin_memory()
    cost_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    solver.reset()
    solver.set_initial_pinned(x0_pin, u_pin)
    solver.begin_solve()
    ex = solver.host_exchange()

    def e2e_step():
        ex.apply_inputs()
        solver.iterate_linesearch()
        ex.read_controls(u_pin)
        solver.iterate_finish_async()
        ex.wait_controls()
        ex.stage_inputs(x0_pin, u_pin)
        solver.iterate_wait()
        gather_costs()
        solver.get_into(_lib.COST, cost_pin)

    ex.stage_inputs(x0_pin, u_pin)
    for _ in range(W):
        e2e_step()
    it0e = solver.get_int(_lib.I_ITERS).astype(np.int64)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(solver._stream)
    for _ in range(K):
        e2e_step()
    f1.record(solver._stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    units_e2e_local = int((solver.get_int(_lib.I_ITERS).astype(np.int64) - it0e).sum())
    e2e_cost_match = float(np.abs(cost_pin.numpy() - cost_dev).max() / np.abs(cost_dev).max())

    # ------------------------------------------------------------------ reduce over ranks
    stats = torch.tensor([elapsed_ms, e2e_ms, float(units_local), float(units_e2e_local), float(launches)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_ms = float(mx[0]), float(mx[1])
        units, units_e2e, launches_all = float(sm[2]), float(sm[3]), int(sm[4])
    else:
        units, units_e2e, launches_all = float(units_local), float(units_e2e_local), int(launches)

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
        dom = max(phase_ms, key=phase_ms.get)
        active_per_step = units_local / K
        if dom == "backward":
            kernel, bytes_per_launch, launch_ms = "backward_mma_kernel", pb["backward"] * active_per_step, phase_ms[dom] / K
        elif dom == "derivs":
            kernel, bytes_per_launch, launch_ms = "quad_fused_kernel", pb["derivs"] * active_per_step, phase_ms[dom] / K
        else:
            # line-search phase: every round launches one rollout kernel over A candidates
            kernel = "rollout_quad8_kernel (line-search phase, all rounds)"
            bytes_per_launch = pb["rollout"] * active_per_step * float(np.mean(ls))
            launch_ms = phase_ms[dom] / K
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel.split()[0])
        except Exception:
            pass
        tf = {}
        import ctypes
        for mma, nm in ((0, "dfma"), (1, "dmma")):
            v = ctypes.c_double()
            if _lib.lib().ddp_peak_fp64(None, mma, ctypes.byref(v)) == 0:
                tf[nm] = round(v.value, 2)
        bwd_flops = 2.0 * T * (2 * n ** 3 + 3 * n * n * m + 2 * n * m * m) * active_per_step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, prob, world),
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
                         "whole_iteration": {
                             "algorithmic_bytes_per_trajectory_iteration": pb["backward"] + pb["derivs"] + pb["rollout"],
                             "achieved_GBps": (pb["backward"] + pb["derivs"] + pb["rollout"]) * value / world / 1e9,
                             "frac": (pb["backward"] + pb["derivs"] + pb["rollout"]) * value / world / 1e9 / hbm_peak},
                         "fp64": {"measured_peak_tflops": tf,
                                  "backward_tflops": bwd_flops / (phase_ms["backward"] / K * 1e-3) / 1e12,
                                  # DMMAs the sweep issues per step at (36, 12): 681 for the products (8 x 8
                                  # tile padding included) + 24 per Newton-Schulz pass, 2.8 passes on average
                                  # (DESIGN.md section 3); 512 flop each, against the measured DMMA peak
                                  "backward_dmma_issue_frac": (
                                      (681 + 24 * 2.8) * 512.0 * T * active_per_step
                                      / (phase_ms["backward"] / K * 1e-3) / 1e12 / tf["dmma"])
                                  if (n, m) == (36, 12) and tf.get("dmma") else None}},
            "phase_ms_per_step": {k: v / K for k, v in phase_ms.items()},
            "ls_iters_mean_last_step": float(np.mean(ls)), "ls_parallel": solver.A,
            "trajectory_status": {"running": int((status == 0).sum()), "converged": int((status == 1).sum()),
                                  "linesearch_failed": int((status == 2).sum())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (n + T * m) * 8),
                    "d2h_bytes_per_step": int(B * (T * m + 1) * 8), "ms_per_step": e2e_ms / K,
                    "cost_match_vs_device_resident": e2e_cost_match},
            "gpu_launches": launches_all, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload on the host cores (about 10-30 s of CPU work)
            cpu_steps, cpu_per_proc = 12, 8
            cv, procs, cdt, cunits = cpu_measure(args.horizon, cpu_steps, 1, per_proc=cpu_per_proc)
            line["cpu_baseline"] = {
                "value": cv, "unit": UNIT, "cores": procs, "kind": "port",
                "sample": f"{cpu_per_proc * procs} trajectories ({cpu_per_proc} per process, {procs} processes, 1 thread each) x {cpu_steps} "
                          f"iLQR iterations of the same C4 problem after 1 warm-up iteration "
                          f"({cunits} trajectory-iterations in {cdt:.1f} s wall)"}
            # cost vs reference: trajectory 0 re-solved by the oracle.  Compared after 2 iterations:
            # the N=200 open-loop-unstable contact problem amplifies a 1e-15 input perturbation to
            # 1e-4 after four iterations in the oracle itself (DESIGN.md "Conditioning"), so longer
            # horizons of iterations compare chaos, not implementations.
            from oracle.dynamics import HostDynamics
            from oracle.ilqr_port import IlqrOracle
            o = IlqrOracle(HostDynamics(prob.system), N, delta=prob.delta, beta=prob.beta, gamma=prob.gamma)
            o.set_initial_state(x0[0]); o.set_target_state(prob.x_nom)
            o.set_running_cost(prob.Q, prob.R); o.set_terminal_cost(prob.Qf); o.set_initial_guess(prob.u_guess)
            try:
                o.solve(max_iters=2)
                Lo = o.trace[-1].L
                line["cost_vs_oracle"] = {"trajectory": 0, "iterations": 2, "gpu": cost_after[2],
                                          "oracle": Lo, "rel_err": abs(cost_after[2] - Lo) / abs(Lo)}
            except RuntimeError as e:
                line["cost_vs_oracle"] = {"error": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
