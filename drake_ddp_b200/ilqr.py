"""Host side of the B200 iLQR: the reference's solver class surface over the CUDA C ABI.

``IterativeLinearQuadraticRegulator`` keeps the constructor, setters, ``Solve()`` return
tuple, ``SaveSolution`` and public attributes of /root/reference/ilqr.py:12-733 so the
reference's example scripts can switch to it; ``BatchedILQR`` is the same solver for B
independent trajectories (MPC resolves, initial-condition seeds), which is what the GPU is
for.  All arithmetic happens in ``libddp_b200.so`` (hand-written sm_100a kernels); torch
only provides device memory and the stream.  There is no CPU path: constructing a solver
without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np

from . import _lib
from .utils_derivs_interpolation import derivs_interpolation


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _drake_names(system):
    """Lower-cased names a Drake System / MultibodyPlant / Diagram offers for model matching."""
    names = []
    plant = system
    try:
        plant = system.GetSubsystemByName("plant")      # Diagram with a plant (cart_pole_with_wall.py:135)
    except Exception:
        pass
    for obj in (system, plant):
        for getter in ("get_name", "GetSystemName"):
            try:
                names.append(str(getattr(obj, getter)()))
            except Exception:
                pass
    try:                                                # MultibodyPlant: model instance names
        for i in range(plant.num_model_instances()):
            try:
                from pydrake.multibody.tree import ModelInstanceIndex
                names.append(str(plant.GetModelInstanceName(ModelInstanceIndex(i))))
            except Exception:
                break
    except Exception:
        pass
    n_geom = 0
    try:
        n_geom = int(plant.num_collision_geometries())
    except Exception:
        pass
    return [nm.lower() for nm in names], n_geom


def _as_system(system, input_port_index=0):
    """The ``system`` argument of the reference constructor (/root/reference/ilqr.py:21-58).

    Accepts (i) a native ``AnalyticSystem`` (has model_id/params) and (ii) a discrete-time Drake
    ``System`` -- a bare ``MultibodyPlant`` plus actuation port index (pendulum.py:74-86) or a
    ``Diagram`` with an exported input (cart_pole_with_wall.py:135-148).  A Drake object is only
    *read*, through the same calls the reference makes on it: ``IsDifferenceEquationSystem()``
    (:37), ``CreateDefaultContext().get_discrete_state_vector().size()`` (:57),
    ``get_input_port(i).size()`` (:58) and ``time_step()`` (:725); (n, m, dt, names) select the
    matching fixed analytic model of ``drake_ddp_b200.systems``, which is what the kernels
    evaluate.  Raises when no analytic model matches."""
    if hasattr(system, "model_id") and hasattr(system, "params"):
        return system
    if not (hasattr(system, "IsDifferenceEquationSystem") and hasattr(system, "CreateDefaultContext")):
        raise TypeError(
            "system must be a drake_ddp_b200.systems.AnalyticSystem or a discrete-time Drake System "
            "(IsDifferenceEquationSystem / CreateDefaultContext / get_input_port)")
    from . import systems
    res = system.IsDifferenceEquationSystem()
    is_discrete = res[0] if isinstance(res, (tuple, list)) else res
    assert is_discrete, "must be a discrete-time system"                       # ilqr.py:37
    n = int(system.CreateDefaultContext().get_discrete_state_vector().size())   # ilqr.py:57
    m = int(system.get_input_port(input_port_index).size())                     # ilqr.py:58
    dt = None
    for get in (lambda: system.time_step(), lambda: system.GetSubsystemByName("plant").time_step(),
                lambda: res[1]):
        try:
            dt = float(get())
            if dt > 0:
                break
        except Exception:
            continue
    if not dt or dt <= 0:
        raise TypeError("cannot read the time step of the Drake system")
    names, n_geom = _drake_names(system)
    has = lambda *keys: any(k in nm for nm in names for k in keys)
    if (n, m) == (2, 1):
        return systems.pendulum(dt=dt)
    if (n, m) == (4, 1):
        if has("acrobot"):
            return systems.acrobot(dt=dt)
        if has("wall") or (has("cart") and n_geom > 0):
            return systems.cart_pole_with_wall(dt=dt)
        if has("cart"):
            return systems.cart_pole(dt=dt)
    if (n, m) == (37, 12):
        return systems.quadruped_quat(dt=dt)      # mini_cheetah.py:41-52 (quaternion floating base)
    if (n, m) == (36, 12):
        return systems.quadruped(dt=dt)
    if (n, m) == (27, 7):
        return systems.arm_ball(dt=dt)            # kinova_gen3.py:68 / panda_fr3.py (7 + 7 + 13)
    raise TypeError(
        f"no analytic model for a Drake system with n={n}, m={m}, names={names}: the GPU path "
        "evaluates fixed analytic models (drake_ddp_b200.systems); pass one explicitly")


class BatchedILQR:
    """B independent iLQR problems sharing one model, horizon and cost."""

    def __init__(self, system, num_timesteps, batch=1, delta=1e-2, beta=0.95, gamma=0.0,
                 derivs_keypoint_method=None, ls_parallel=None, device=None, input_port_index=0):
        import torch

        L = _lib.lib()  # raises if the CUDA extension is missing
        if not torch.cuda.is_available():
            raise RuntimeError("drake_ddp_b200 needs a CUDA device (no CPU fallback)")
        self._torch = torch
        self._L = L
        self.system = _as_system(system, input_port_index)
        self.N, self.B = int(num_timesteps), int(batch)
        self.n, self.m = self.system.n, self.system.m
        self.T = self.N - 1
        self.delta, self.beta, self.gamma = float(delta), float(beta), float(gamma)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        n, m, npar = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(L.ddp_model_dims(self.system.model_id, ctypes.byref(n), ctypes.byref(m),
                                    ctypes.byref(npar)), "ddp_model_dims")
        assert (n.value, m.value) == (self.n, self.m), "system dims do not match the compiled model"
        params = np.ascontiguousarray(self.system.params, dtype=np.float64)
        assert params.size == npar.value, f"model expects {npar.value} parameters, got {params.size}"
        n_eps = 0
        eps = 1.0
        while eps >= 1e-8:
            n_eps += 1
            eps *= self.beta
        if ls_parallel is None:
            # Candidates evaluated speculatively per trajectory in the first line-search round.  A
            # rollout is N-1 dependent steps, so a round costs the same whether it carries one
            # candidate per trajectory or as many as fit one resident wave of the GPU: take as many
            # as that (8 for the quadruped kernels: their CTA is 8 candidates of one trajectory, and
            # 8 resolve > 90 % of the trajectories), within 2 GiB of candidate buffer.
            per_rollout = 8 * (self.N * self.n + self.T * self.m)
            # quadruped, quadruped_quat, arm_ball: the 8-lane rollouts; 8 candidates per trajectory
            # keep a 512..1024-trajectory batch within one resident wave of their CTAs
            quad = self.system.model_id in (4, 5, 6)
            lanes = 1 if (self.n + self.m) <= 8 else (8 if quad else 4)
            wave = 148 * 2048 // max(1, self.B * lanes)
            cap = 8 if quad else n_eps
            ls_parallel = max(1, min(cap, max(8, wave) if cap > 8 else 8, (2 << 30) // max(1, self.B * per_rollout)))
        self.A = max(1, min(int(ls_parallel), n_eps))
        nbytes = L.ddp_workspace_bytes(self.system.model_id, self.N, self.B, self.A)
        assert nbytes > 0, "bad (model, N, B, A)"
        with torch.cuda.device(self.device):
            self._arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            # the solver's own stream on its own device: the library switches to the arena's
            # device inside every call, so the caller may keep any GPU current
            self._stream = torch.cuda.Stream(device=self.device)
            h = ctypes.c_void_p()
            _lib.check(L.ddp_create(ctypes.byref(h), self.system.model_id, _ptr(params), params.size,
                                    self.N, self.B, self.A, ctypes.c_void_p(self._arena.data_ptr()),
                                    nbytes, ctypes.c_void_p(self._stream.cuda_stream)), "ddp_create")
        self._h = h
        _lib.check(L.ddp_set_options(h, self.delta, self.beta, self.gamma), "ddp_set_options")
        self.set_keypoint_method(derivs_keypoint_method)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and self._L is not None:
            try:
                self._L.ddp_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- configuration ------------------------------------------------------------------
    def set_keypoint_method(self, cfg):
        """derivs_keypoint_method of the reference ctor (ilqr.py:97-100)."""
        if cfg is None:
            cfg = derivs_interpolation("setInterval", 1, 0, 0, 0)
        if cfg.keypoint_method not in _lib.KP_METHODS:
            raise Exception("unknown interpolation method")  # ilqr.py:404
        self.derivs_interpolation = cfg
        _lib.check(self._L.ddp_set_keypoints(self._h, _lib.KP_METHODS[cfg.keypoint_method], int(cfg.minN),
                                             int(cfg.maxN), float(cfg.jerk_threshold),
                                             float(cfg.iterative_error_threshold)), "ddp_set_keypoints")

    def set_regularization(self, quu_reg: float):
        """Extension: Quu + quu_reg*I is inverted in the backward pass (0 = the reference, ilqr.py:654-655)."""
        _lib.check(self._L.ddp_set_regularization(self._h, float(quu_reg)), "ddp_set_regularization")

    def set_control_limits(self, u_min=None, u_max=None):
        """Extension (the reference's SetControlLimits is `pass`, ilqr.py:158-159): clamp every
        rollout control to [u_min, u_max]; None switches it off (default = reference behaviour)."""
        if u_min is None or u_max is None:
            _lib.check(self._L.ddp_set_control_limits(self._h, None, None), "ddp_set_control_limits")
            return
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(u_min, dtype=np.float64), (self.m,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(u_max, dtype=np.float64), (self.m,)))
        _lib.check(self._L.ddp_set_control_limits(self._h, _ptr(lo), _ptr(hi)), "ddp_set_control_limits")

    def set_cost(self, Q, R, Qf):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        R = np.ascontiguousarray(R, dtype=np.float64)
        Qf = np.ascontiguousarray(Qf, dtype=np.float64)
        assert Q.shape == (self.n, self.n) and R.shape == (self.m, self.m) and Qf.shape == (self.n, self.n)
        _lib.check(self._L.ddp_set_cost(self._h, _ptr(Q), _ptr(R), _ptr(Qf)), "ddp_set_cost")

    def set_target(self, x_nom):
        x_nom = np.ascontiguousarray(x_nom, dtype=np.float64)
        per = 1 if x_nom.ndim == 2 else 0
        assert x_nom.shape[-1] == self.n and (not per or x_nom.shape[0] == self.B)
        _lib.check(self._L.ddp_set_target(self._h, _ptr(x_nom), per), "ddp_set_target")

    def set_initial_state(self, x0):
        """x0: (B, n) or (n,) broadcast."""
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.ndim == 1:
            x0 = np.broadcast_to(x0, (self.B, self.n))
        x0 = np.ascontiguousarray(x0)
        assert x0.shape == (self.B, self.n)
        _lib.check(self._L.ddp_set_initial_state(self._h, _ptr(x0)), "ddp_set_initial_state")

    def set_initial_guess(self, u_guess):
        """u_guess: (B, T, m) device layout, or the reference's (m, T) broadcast over the batch."""
        u = np.asarray(u_guess, dtype=np.float64)
        if u.ndim == 2:
            assert u.shape == (self.m, self.T)
            u = np.broadcast_to(u.T, (self.B, self.T, self.m))
        u = np.ascontiguousarray(u)
        assert u.shape == (self.B, self.T, self.m)
        _lib.check(self._L.ddp_set_initial_guess(self._h, _ptr(u)), "ddp_set_initial_guess")

    def set_initial_pinned(self, x0_pinned, u_pinned):
        """Same as the two setters above but from caller-owned (pinned) buffers, no reshaping."""
        _lib.check(self._L.ddp_set_initial_state(self._h, ctypes.c_void_p(x0_pinned.data_ptr())), "x0")
        _lib.check(self._L.ddp_set_initial_guess(self._h, ctypes.c_void_p(u_pinned.data_ptr())), "u")

    def reset(self):
        _lib.check(self._L.ddp_reset(self._h), "ddp_reset")

    def mpc_shift(self, replan_steps: int):
        """Device-side receding-horizon warm start: shift the control tape by ``replan_steps``,
        pad with the last control, x0 <- x_bar[:, replan_steps] (acrobot.py:145-153)."""
        _lib.check(self._L.ddp_mpc_shift(self._h, int(replan_steps)), "ddp_mpc_shift")

    def set_mpc_rearm(self, replan_steps: int, target_advance=None):
        """The receding-horizon loop of mini_cheetah.py:186-206 on the device: a trajectory whose
        Solve() converges is shifted by ``replan_steps`` (x0 <- x_bar[:, r], tape padded, x_nom +=
        ``target_advance``) and keeps iterating as its next resolve.  0 switches it off."""
        adv = None
        if target_advance is not None:
            adv = np.ascontiguousarray(target_advance, dtype=np.float64)
            assert adv.shape == (self.n,)
        _lib.check(self._L.ddp_set_mpc_rearm(self._h, int(replan_steps), None if adv is None else _ptr(adv)),
                   "ddp_set_mpc_rearm")

    # ---- solve ----------------------------------------------------------------------------
    def begin_solve(self):
        _lib.check(self._L.ddp_begin_solve(self._h), "ddp_begin_solve")

    def iterate(self) -> int:
        """One forward + backward pass for every unconverged trajectory; returns #still active."""
        n_active = ctypes.c_int()
        _lib.check(self._L.ddp_iterate(self._h, ctypes.byref(n_active)), "ddp_iterate")
        return n_active.value

    # one iteration in three calls (include/ddp_b200.h): the line search is synchronous, the
    # derivatives + backward pass are enqueued and waited for separately, so the caller can move
    # host buffers while they run (HostExchange below)
    def iterate_linesearch(self):
        _lib.check(self._L.ddp_iterate_linesearch(self._h), "ddp_iterate_linesearch")

    def iterate_finish_async(self):
        _lib.check(self._L.ddp_iterate_finish_async(self._h), "ddp_iterate_finish_async")

    def iterate_wait(self) -> int:
        n_active = ctypes.c_int()
        _lib.check(self._L.ddp_iterate_wait(self._h, ctypes.byref(n_active)), "ddp_iterate_wait")
        return n_active.value

    def host_exchange(self):
        return HostExchange(self)

    def solve(self, max_iters=0) -> int:
        it = ctypes.c_int()
        _lib.check(self._L.ddp_solve(self._h, int(max_iters), ctypes.byref(it)), "ddp_solve")
        return it.value

    def run_phase(self, phase: int):
        _lib.check(self._L.ddp_run_phase(self._h, phase), "ddp_run_phase")

    # ---- array access -------------------------------------------------------------------
    _SHAPES = {
        _lib.X_BAR: lambda s: (s.B, s.N, s.n), _lib.U_BAR: lambda s: (s.B, s.T, s.m),
        _lib.K: lambda s: (s.B, s.T, s.m, s.n), _lib.KAPPA: lambda s: (s.B, s.T, s.m),
        _lib.DV: lambda s: (s.B, s.T), _lib.FX: lambda s: (s.B, s.T, s.n, s.n),
        _lib.FU: lambda s: (s.B, s.T, s.n, s.m), _lib.COST: lambda s: (s.B,),
        _lib.EPS: lambda s: (s.B,), _lib.IMPROVEMENT: lambda s: (s.B,),
        _lib.X0: lambda s: (s.B, s.n), _lib.X_NOM: lambda s: (s.B, s.n),
        _lib.CAND_COST: lambda s: (s.B, s.A), _lib.CAND_EXPECTED: lambda s: (s.B, s.A),
        _lib.CAND_X: lambda s: (s.B, s.A, s.N, s.n), _lib.CAND_U: lambda s: (s.B, s.A, s.T, s.m),
        _lib.CONVERGED_COST: lambda s: (s.B,),
    }

    def get(self, which: int) -> np.ndarray:
        out = np.empty(self._SHAPES[which](self), dtype=np.float64)
        _lib.check(self._L.ddp_get(self._h, which, _ptr(out)), "ddp_get")
        return out

    def get_into(self, which: int, dst_pinned):
        _lib.check(self._L.ddp_get(self._h, which, ctypes.c_void_p(dst_pinned.data_ptr())), "ddp_get")

    def put(self, which: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._SHAPES[which](self), (arr.shape, self._SHAPES[which](self))
        _lib.check(self._L.ddp_put(self._h, which, _ptr(arr)), "ddp_put")

    def get_int(self, which: int) -> np.ndarray:
        shape = (self.B, self.T) if which == _lib.I_KEYPOINTS else (self.B,)
        out = np.empty(shape, dtype=np.int32)
        _lib.check(self._L.ddp_get_int(self._h, which, _ptr(out)), "ddp_get_int")
        return out

    def device_tensor(self, which: int):
        """Zero-copy torch view of an arena array (e.g. COST for an NCCL all-gather)."""
        torch = self._torch
        ptr = self._L.ddp_device_ptr(self._h, which)
        off = ptr - self._arena.data_ptr()
        nelem = self._L.ddp_array_elems(self._h, which)
        return self._arena[off: off + 8 * nelem].view(torch.float64).view(self._SHAPES[which](self))

    def keypoints(self):
        cnt = self.get_int(_lib.I_NUM_KEYPOINTS)
        kp = self.get_int(_lib.I_KEYPOINTS)
        return [list(map(int, kp[b, :cnt[b]])) for b in range(self.B)]

    def timings_ms(self):
        ms = (ctypes.c_float * 4)()
        _lib.check(self._L.ddp_last_timings(self._h, ms), "ddp_last_timings")
        return {"linesearch": ms[0], "derivs": ms[1], "backward": ms[2], "iteration": ms[3]}

    def launch_count(self) -> int:
        return int(self._L.ddp_launch_count(self._h))

    # convenience getters in device layout
    @property
    def cost(self):
        return self.get(_lib.COST)

    @property
    def status(self):
        return self.get_int(_lib.I_STATUS)


class HostExchange:
    """Host <-> device traffic of an MPC-style loop overlapped with the iteration.

    Every iteration of such a loop (acrobot.py:142-160, mini_cheetah.py:186-206) sends x0 and a
    control tape to the solver and reads the new tape back.  The tape of an iteration is final
    after its line search, and the derivatives + backward pass that follow only read it, so

        ex.stage_inputs(x0_pinned, u_pinned)      # H2D into a staging buffer, copy stream
        loop:
            ex.apply_inputs()                     # staging -> x0, u_bar (device copy, solver stream;
                                                  # rows the device re-armed itself are kept)
            solver.iterate_linesearch()
            ex.read_controls(u_pinned)            # D2H of u_bar on the copy stream ...
            solver.iterate_finish_async()         # ... under derivatives + backward pass
            ex.wait_controls()                    # the host owns the new tape
            ex.stage_inputs(x0_pinned, u_pinned)  # next inputs travel under the backward pass
            solver.iterate_wait()

    moves the same bytes as set_initial_* / get per iteration but hides them behind the kernels.
    The staging buffer decouples the upload from u_bar, which the backward pass is still reading.
    """

    def __init__(self, solver: "BatchedILQR"):
        torch = solver._torch
        self.s = solver
        self._torch = torch
        dev = solver.device
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self.stage_x0 = torch.empty((solver.B, solver.n), dtype=torch.float64, device=dev)
        self.stage_u = torch.empty((solver.B, solver.T, solver.m), dtype=torch.float64, device=dev)
        self.ev_staged = torch.cuda.Event()
        self.ev_applied = torch.cuda.Event()
        self.ev_read = torch.cuda.Event()
        self._x0 = solver.device_tensor(_lib.X0)
        self._u = solver.device_tensor(_lib.U_BAR)
        self.ev_applied.record(solver._stream)
        self._cost = solver.device_tensor(_lib.COST)
        self.ev_patch = torch.cuda.Event()
        self.d2h_bytes = self.d2h_steps = 0

    def stage_inputs(self, x0_pinned, u_pinned):
        torch = self._torch
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.ev_applied)   # the previous staging content has been consumed
            self.stage_x0.copy_(x0_pinned, non_blocking=True)
            self.stage_u.copy_(u_pinned, non_blocking=True)
            self.ev_staged.record(self.copy_in)

    def apply_inputs(self):
        """staging -> x0, u_bar on the solver's stream (ddp_apply_staged_inputs: rows the device
        re-armed itself since the last call keep their newer data)."""
        st = self.s._stream
        st.wait_event(self.ev_staged)
        _lib.check(self.s._L.ddp_apply_staged_inputs(self.s._h, ctypes.c_void_p(self.stage_x0.data_ptr()),
                                                      ctypes.c_void_p(self.stage_u.data_ptr())),
                   "ddp_apply_staged_inputs")
        self.ev_applied.record(st)

    def read_controls(self, u_pinned):
        """Call after iterate_linesearch() (u_bar final, solver stream idle)."""
        torch = self._torch
        with torch.cuda.stream(self.copy_out):
            u_pinned.copy_(self._u, non_blocking=True)
            self.ev_read.record(self.copy_out)
        self.d2h_bytes += u_pinned.numel() * 8
        self.d2h_steps += 1

    def wait_controls(self):
        self.ev_read.synchronize()

    def read_state(self, x0_pinned, cost_pinned):
        """x0 (the device moves it when it re-arms a trajectory) and the costs, to host buffers;
        call after iterate_wait().  Blocks until both have landed."""
        torch = self._torch
        st = self.s._stream
        with torch.cuda.stream(st):
            x0_pinned.copy_(self._x0, non_blocking=True)
            cost_pinned.copy_(self._cost, non_blocking=True)
            self.ev_patch.record(st)
        self.ev_patch.synchronize()
        self.d2h_bytes += (x0_pinned.numel() + cost_pinned.numel()) * 8


class IterativeLinearQuadraticRegulator:
    """Drop-in for the reference class (/root/reference/ilqr.py:12): same constructor
    signature, setters, ``Solve()`` and result attributes; one trajectory (B=1)."""

    def __init__(self, system, num_timesteps, input_port_index=0, delta=1e-2, beta=0.95, gamma=0.0,
                 derivs_keypoint_method=None, ls_parallel=None):
        self.system = _as_system(system, input_port_index)
        self.N = num_timesteps
        self.delta, self.beta, self.gamma = delta, beta, gamma
        self.n, self.m = self.system.n, self.system.m
        self._core = BatchedILQR(self.system, num_timesteps, batch=1, delta=delta, beta=beta,
                                 gamma=gamma, derivs_keypoint_method=derivs_keypoint_method,
                                 ls_parallel=ls_parallel)
        self.derivs_interpolation = self._core.derivs_interpolation
        # ilqr.py:61-67
        self.x0 = np.zeros(self.n)
        self.Q, self.R, self.Qf = np.eye(self.n), np.eye(self.m), np.eye(self.n)
        self._u_guess = None
        self.time_getDerivs = 0
        self.percentage_derivs = 0
        self.time_backwardsPass = 0
        self.time_fp = 0

    # ---- setters: keep references, upload at Solve() like the reference reads them then ----
    def SetInitialState(self, x0):
        self.x0 = x0

    def SetTargetState(self, x_nom):
        self.x_nom = np.asarray(x_nom).reshape((self.n,))

    def SetRunningCost(self, Q, R):
        assert Q.shape == (self.n, self.n)
        assert R.shape == (self.m, self.m)
        self.Q = Q
        self.R = R

    def SetTerminalCost(self, Qf):
        assert Qf.shape == (self.n, self.n)
        self.Qf = Qf

    def SetInitialGuess(self, u_guess):
        assert u_guess.shape == (self.m, self.N - 1)
        self._u_guess = u_guess

    def SetControlLimits(self, u_min, u_max):
        pass  # no-op in the reference too (ilqr.py:158-159); BatchedILQR.set_control_limits is the extension

    # ---- results in the reference's layouts (time last, ilqr.py:70-83) ---------------------
    @property
    def x_bar(self):
        return np.ascontiguousarray(self._core.get(_lib.X_BAR)[0].T)

    @property
    def u_bar(self):
        return np.ascontiguousarray(self._core.get(_lib.U_BAR)[0].T)

    @property
    def kappa(self):
        return np.ascontiguousarray(self._core.get(_lib.KAPPA)[0].T)

    @property
    def K(self):
        return np.ascontiguousarray(self._core.get(_lib.K)[0].transpose(1, 2, 0))

    @property
    def fx(self):
        return np.ascontiguousarray(self._core.get(_lib.FX)[0].transpose(1, 2, 0))

    @property
    def fu(self):
        return np.ascontiguousarray(self._core.get(_lib.FU)[0].transpose(1, 2, 0))

    @property
    def dV_coeff(self):
        return self._core.get(_lib.DV)[0].copy()

    def _upload(self):
        c = self._core
        c.set_cost(self.Q, self.R, self.Qf)
        c.set_target(self.x_nom)
        c.set_initial_state(np.asarray(self.x0, dtype=np.float64).reshape(self.n))
        if self._u_guess is not None:
            c.set_initial_guess(self._u_guess)
            self._u_guess = None  # afterwards u_bar lives on the device (ilqr.py:375)

    def Solve(self):
        """ilqr.py:669-710: returns (x (n,N), u (m,N-1), solve_time, optimal_cost)."""
        c = self._core
        self._upload()
        print("----------------------------------------------------------------------------------------------------------------------------------")
        print("|    iter    |    cost    |    eps    |    ls    | derivs time | derivs '%'  | bp time  | fp time  |   iter time    |    time    |")
        print("----------------------------------------------------------------------------------------------------------------------------------")
        c.begin_solve()
        i = 1
        st = time.time()
        n_active = 1
        total_time = 0.0
        while n_active > 0:
            st_iter = time.time()
            n_active = c.iterate()
            status = int(c.get_int(_lib.I_STATUS)[0])
            ls_iters = int(c.get_int(_lib.I_LS_ITERS)[0])
            if status == _lib.TRAJ_LINESEARCH_FAILED:
                raise RuntimeError("linesearch failed after %s iterations" % ls_iters)  # ilqr.py:337
            L_new = float(c.get(_lib.COST)[0])
            eps = float(c.get(_lib.EPS)[0])
            ms = c.timings_ms()
            self.time_fp = ms["linesearch"] * 1e-3
            self.time_getDerivs = ms["derivs"] * 1e-3
            self.time_backwardsPass = ms["backward"] * 1e-3
            self.percentage_derivs = (int(c.get_int(_lib.I_NUM_KEYPOINTS)[0]) / (self.N - 1)) * 100
            iter_time = time.time() - st_iter
            total_time = time.time() - st
            print(f"{i:^14}{L_new:11.4f}  {eps:^12.4f}{ls_iters:^11}   {self.time_getDerivs:1.5f}         {self.percentage_derivs:.1f}       {self.time_backwardsPass:1.5f}    {self.time_fp:1.5f}      {iter_time:1.5f}          {total_time:4.2f}")
            i += 1
        return self.x_bar, self.u_bar, total_time, L_new

    def SaveSolution(self, fname):
        """ilqr.py:712-733: npz with t, x_bar[:, :-1], u_bar, K."""
        dt = self.system.GetSubsystemByName("plant").time_step()
        T = (self.N - 1) * dt
        t = np.arange(0, T, dt)
        np.savez(fname, t=t, x_bar=self.x_bar[:, :-1], u_bar=self.u_bar, K=self.K)
