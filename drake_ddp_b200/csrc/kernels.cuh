// CUDA kernels of the batched iLQR hot path (sm_100a).  One template per reference
// function; the reference line ranges each one replaces are cited at the kernel.
//
// Layouts (row-major fp64, trajectory-major, time-major tiles):
//   x_bar [B][N][n]  u_bar [B][T][m]  kappa [B][T][m]  dV [B][T]
//   K [B][T][m][n]   fx [B][T][n][n]  fu [B][T][n][m]
//   candidates: xc [B][A][N][n]  uc [B][A][T][m]  Lc, Ec [B][A]
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "models.h"

namespace ddp {

struct Dev {
  int n, m, N, T, B, A, n_eps, diag_cost;
  double delta, gamma;
  double quu_reg;  // added to the diagonal of Quu before it is inverted (0: the reference's behaviour)
  // extension (SetControlLimits is `pass` in the reference, ilqr.py:158-159): when set, the rollout
  // clamps u_t to [u_min, u_max] ([m] each); nullptr = off = the reference's behaviour
  const double *u_min, *u_max;
  double* u_lim_buf;
  const double* params;
  // the first kParamsInline model parameters by value: a kernel argument lives in the constant
  // bank, so p[k] with a literal k becomes an instruction operand (no load, no register, nothing
  // for the compiler to keep inside a loop that stores); the hand-written quadruped / arm kernels
  // read their parameters from here (models with more parameters use `params` only)
  double pm[32];
  const double *Q, *R, *Qf, *x_nom, *x0, *eps_table;
  double *x_bar, *u_bar, *K, *kappa, *dV, *fx, *fu;
  double *xc, *uc, *Lc, *Ec;
  double *L, *L_new, *eps, *improvement;
  int *ls_iters, *status, *active, *resolved, *acc, *iters, *counters, *unres;
  int *active_save, *status_save;  // ddp_run_phase (teacher-forced tests) puts these back
  // receding-horizon driver on the device (mini_cheetah.py:147-159,186-206, acrobot.py:131-162):
  // a trajectory that converges is re-armed as the next MPC resolve instead of being frozen
  int mpc_replan;          // replan_steps (0: off, a converged trajectory stops iterating)
  const double* mpc_target_adv;  // [n] added to x_nom of the trajectory at every resolve (moving target)
  double* mpc_target_adv_buf;
  int *rearm, *resolves;   // [B] flag for mpc_rearm_kernel; resolves finished so far
  int* rearm_mark;         // [B] re-armed on the device since the last ddp_apply_staged_inputs
  double* L_conv;          // [B] final cost of the last converged solve
  int bwd_flags;  // bit 0: backward_sym_kernel inverts Quu by Gauss-Jordan at every step (no Newton-Schulz)
  int* sm_slots;  // per-SM bitmask of the CTA slots in use (backward_sym_kernel deals warp roles by slot)
  // keypoints
  int kp_method, minN, maxN;
  double jerk_thr, err_thr;
  int *kplist, *kpcount, *seg_s, *seg_e;
  unsigned char *flag, *done;
  int *segs[2], *nseg[2], *evallist, *evalcount;
};

// Parameters of a model: from the kernel argument (constant bank) when they fit Dev::pm
template <class Model>
__device__ __forceinline__ const double* model_params(const Dev& d) {
  if constexpr (Model::np <= 32) return d.pm;
  else return d.params;
}

// Thread mapping per model size class.
template <class Model>
struct Cfg {
  static constexpr int n = Model::n, m = Model::m;
  static constexpr bool small = (n + m) <= 8;
  static constexpr int G_ROLL = small ? 1 : 4;      // lanes per rollout candidate
#ifndef DDP_LIN_G
#define DDP_LIN_G 16
#endif
#ifndef DDP_LIN_K
#define DDP_LIN_K 1
#endif
  static constexpr int G_LIN = small ? 1 : DDP_LIN_G;      // lanes per linearization point
  static constexpr int K_LIN = small ? (n + m) : DDP_LIN_K; // seed directions per lane and sweep
  static constexpr int P_LIN = (n + m + G_LIN * K_LIN - 1) / (G_LIN * K_LIN);  // sweeps
  static constexpr int BWD_THREADS = small ? 32 : 128;
};

__device__ __forceinline__ unsigned group_mask(int G) {
  if (G >= 32) return 0xffffffffu;
  const int lane = threadIdx.x & 31;
  return ((1u << G) - 1u) << (lane / G * G);
}

// =============================================================================================
// K1  closed-loop rollout + cost of line-search candidates.
// Replaces the body of _linesearch (/root/reference/ilqr.py:306-327) and _calc_dynamics
// (:208-231) for candidates eps_table[ls_base .. ls_base+A) of every unresolved trajectory.
// G lanes cooperate on one candidate: feedback rows are split over the lanes, the dynamics
// are evaluated redundantly by each lane (identical values, so control flow stays uniform).
// =============================================================================================
template <class Model, int G>
__global__ void __launch_bounds__(128) rollout_kernel(Dev d, int ls_base, int per_traj, int n_items) {
  constexpr int n = Model::n, m = Model::m;
  constexpr int RPL = (m + G - 1) / G;  // feedback rows per lane
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = gtid / G, lane = gtid % G;
  // work item -> (trajectory b, candidate c).  Round 0: every trajectory, candidates
  // [0, A): item = b*A + ai.  Later rounds: only the trajectories still unresolved (compacted
  // in d.unres), per_traj candidates each starting at ls_base; the slot index is the item.
  if (item >= n_items) return;
  int b, ai;
  if (ls_base == 0) {
    b = item / per_traj;
    ai = item % per_traj;
    if (!d.active[b] || d.resolved[b]) return;
  } else {
    b = d.unres[item / per_traj];
    ai = item % per_traj;
  }
  const int c = ls_base + ai;
  if (c >= d.n_eps) {
    if (lane == 0) {
      d.Lc[item] = nan("");
      d.Ec[item] = 0.0;
    }
    return;
  }
  const unsigned mask = group_mask(G);
  const int gbase = (threadIdx.x & 31) / G * G;
  const double eps = d.eps_table[c];
  const double ecoef = -eps * (1.0 - eps / 2.0);
  const int N = d.N, T = d.T;
  const double* xnom = d.x_nom + (size_t)b * n;
  double* xo = d.xc + (size_t)item * N * n;
  double* uo = d.uc + (size_t)item * T * m;

  double x[n], xn[n], u[m];
#pragma unroll
  for (int j = 0; j < n; ++j) x[j] = d.x0[(size_t)b * n + j];
#pragma unroll
  for (int j = 0; j < n; ++j)
    if (j % G == lane) xo[j] = x[j];

  double L = 0.0, E = 0.0;
  bool ok = true;
  int t = 0;
  // Small models (n + m <= 8, one lane per candidate): a step is a few dozen flops, so the global
  // loads of its operands (K_t, x_bar_t, u_bar_t, kappa_t, dV_t: L2 latency) would be most of it.
  // They are fetched one step ahead into registers (m n + n + 2 m + 1 doubles).
  constexpr bool PRE = Cfg<Model>::small && G == 1;
  double nK[PRE ? m * n : 1], nxb[PRE ? n : 1], nub[PRE ? m : 1], nkp[PRE ? m : 1], ndv = 0.0;
  auto prefetch_step = [&](int tt) {
    if constexpr (PRE) {
      const double* gK = d.K + ((size_t)b * T + tt) * m * n;
      const double* gx = d.x_bar + ((size_t)b * N + tt) * n;
      const double* gu = d.u_bar + ((size_t)b * T + tt) * m;
      const double* gk = d.kappa + ((size_t)b * T + tt) * m;
#pragma unroll
      for (int i = 0; i < m * n; ++i) nK[i] = gK[i];
#pragma unroll
      for (int i = 0; i < n; ++i) nxb[i] = gx[i];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        nub[i] = gu[i];
        nkp[i] = gk[i];
      }
      ndv = d.dV[(size_t)b * T + tt];
    }
  };
  prefetch_step(0);
  for (; t < T; ++t) {
    double cK[PRE ? m * n : 1], cxb[PRE ? n : 1], cub[PRE ? m : 1], ckp[PRE ? m : 1], cdv = 0.0;
    if constexpr (PRE) {
#pragma unroll
      for (int i = 0; i < m * n; ++i) cK[i] = nK[i];
#pragma unroll
      for (int i = 0; i < n; ++i) cxb[i] = nxb[i];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        cub[i] = nub[i];
        ckp[i] = nkp[i];
      }
      cdv = ndv;
      if (t + 1 < T) prefetch_step(t + 1);
    }
    const double* Kt = d.K + ((size_t)b * T + t) * m * n;
    const double* xb = d.x_bar + ((size_t)b * N + t) * n;
    const double* ub = d.u_bar + ((size_t)b * T + t) * m;
    const double* kp = d.kappa + ((size_t)b * T + t) * m;
    // operand accessors: the register copies (compile-time indices after unrolling) or global memory
    auto opK = [&](int i) { if constexpr (PRE) return cK[i]; else return Kt[i]; };
    auto opX = [&](int j) { if constexpr (PRE) return cxb[j]; else return xb[j]; };
    auto opU = [&](int r) { if constexpr (PRE) return cub[r]; else return ub[r]; };
    auto opP = [&](int r) { if constexpr (PRE) return ckp[r]; else return kp[r]; };
    // pull the next step's gain rows towards L2/L1 while this step computes
    if (!PRE && t + 1 < T) {
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int r = lane + G * i;
        if (r < m) {
          const char* pr = reinterpret_cast<const char*>(Kt + (size_t)m * n + (size_t)r * n);
#pragma unroll
          for (int off = 0; off < n * 8; off += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + off));
        }
      }
    }
    // u_t = u_bar_t - eps*kappa_t - K_t (x_t - x_bar_t)            (ilqr.py:313)
    double dx[n];
#pragma unroll
    for (int j = 0; j < n; ++j) dx[j] = x[j] - opX(j);
    double mine[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const int r = lane + G * i;
      double acc = 0.0;
      if (r < m) {
#pragma unroll
        for (int j = 0; j < n; ++j) acc = fma(opK(r * n + j), dx[j], acc);
        acc = opU(r) - eps * opP(r) - acc;
        if (d.u_min) acc = fmin(fmax(acc, d.u_min[r]), d.u_max[r]);   // extension, off by default
      }
      mine[i] = acc;
    }
#pragma unroll
    for (int r = 0; r < m; ++r) {
      if (G == 1) u[r] = mine[r];
      else u[r] = __shfl_sync(mask, mine[r / G], gbase + (r % G));
    }
    // x_{t+1} = f(x_t, u_t)                                          (ilqr.py:316)
    if constexpr (Model::COOP == G && G > 1) Model::step_coop(lane, mask, gbase, x, u, xn, model_params<Model>(d));
    else Model::template step<double>(x, u, xn, model_params<Model>(d));
    bool fin = true;
#pragma unroll
    for (int j = 0; j < n; ++j) fin = fin && isfinite(xn[j]);
    if (!fin) {  // the reference gets a RuntimeError from Drake: L = inf, stop   (:317-323)
      ok = false;
      break;
    }
    // running cost uses the pre-step state                         (ilqr.py:325)
    if (d.diag_cost) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const double e = x[j] - xnom[j];
        s = fma(d.Q[j * n + j] * e, e, s);
      }
#pragma unroll
      for (int r = 0; r < m; ++r) s = fma(d.R[r * m + r] * u[r], u[r], s);
      L += s;
    } else {
      double s = 0.0;
#pragma unroll 1
      for (int j = lane; j < n; j += G) {
        double row = 0.0;
        for (int k = 0; k < n; ++k) row = fma(d.Q[j * n + k], x[k] - xnom[k], row);
        s = fma(x[j] - xnom[j], row, s);
      }
#pragma unroll 1
      for (int r = lane; r < m; r += G) {
        double row = 0.0;
        for (int k = 0; k < m; ++k) row = fma(d.R[r * m + k], u[k], row);
        s = fma(u[r], row, s);
      }
      L += s;
    }
    E += ecoef * (PRE ? cdv : d.dV[(size_t)b * T + t]);        //  (ilqr.py:326)
#pragma unroll
    for (int r = 0; r < m; ++r)
      if (r % G == lane) uo[(size_t)t * m + r] = u[r];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      x[j] = xn[j];
      if (j % G == lane) xo[(size_t)(t + 1) * n + j] = xn[j];
    }
  }
  // terminal cost                                                   (ilqr.py:327)
  if (ok) {
    if (d.diag_cost) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const double e = x[j] - xnom[j];
        s = fma(d.Qf[j * n + j] * e, e, s);
      }
      L += s;
    } else {
      double s = 0.0;
#pragma unroll 1
      for (int j = lane; j < n; j += G) {
        double row = 0.0;
        for (int k = 0; k < n; ++k) row = fma(d.Qf[j * n + k], x[k] - xnom[k], row);
        s = fma(x[j] - xnom[j], row, s);
      }
      L += s;
    }
    if (!d.diag_cost && G > 1) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) L += __shfl_xor_sync(mask, L, o);
    }
  } else {
    L = INFINITY;
  }
  if (lane == 0) {
    d.Lc[item] = L;
    d.Ec[item] = E;
  }
}

// =============================================================================================
// K2  first-satisfying pick (ilqr.py:329-337) over the A candidates of this round.
// =============================================================================================
__global__ void pick_kernel(Dev d, int ls_base, int per_traj, int n_traj) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_traj) return;
  const int b = (ls_base == 0) ? j : d.unres[j];
  if (ls_base == 0 && (!d.active[b] || d.resolved[b])) return;
  const double Llast = d.L[b];
  for (int ai = 0; ai < per_traj; ++ai) {
    const int c = ls_base + ai;
    if (c >= d.n_eps) break;
    const size_t slot = (size_t)j * per_traj + ai;
    const double Lcand = d.Lc[slot];
    const double improvement = Llast - Lcand;
    if (improvement > d.gamma * d.Ec[slot]) {
      d.acc[b] = (int)slot;
      d.resolved[b] = 1;
      d.eps[b] = d.eps_table[c];
      d.ls_iters[b] = c + 1;
      d.L_new[b] = Lcand;
      return;
    }
  }
  d.acc[b] = -1;
  if (ls_base + per_traj >= d.n_eps) {  // eps < 1e-8: RuntimeError("linesearch failed ...")  (:337)
    d.status[b] = 2;
    d.active[b] = 0;
    d.resolved[b] = 1;
    d.ls_iters[b] = d.n_eps;
  }
}
// compact the trajectories that are still unresolved into d.unres (order is irrelevant:
// candidates are independent), count in counters[0]
__global__ void unresolved_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  if (d.active[b] && !d.resolved[b]) d.unres[atomicAdd(&d.counters[0], 1)] = b;
}

// u_bar <- u, x_bar <- x of the accepted candidate (ilqr.py:375-376).
__global__ void commit_kernel(Dev d) {
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  const int slot = d.acc[b];
  if (slot < 0) return;
  const size_t nx = (size_t)d.N * d.n, nu = (size_t)d.T * d.m;
  const double* xs = d.xc + (size_t)slot * nx;
  const double* us = d.uc + (size_t)slot * nu;
  double* xd = d.x_bar + (size_t)b * nx;
  double* ud = d.u_bar + (size_t)b * nu;
  for (size_t i = blockIdx.y * blockDim.x + threadIdx.x; i < nx + nu; i += (size_t)gridDim.y * blockDim.x) {
    if (i < nx) xd[i] = xs[i];
    else ud[i - nx] = us[i - nx];
  }
}
// mark the commit of trajectory b as consumed
__global__ void commit_done_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < d.B) d.acc[b] = -1;
}

// Receding-horizon warm start on the device (acrobot.py:145-153, mini_cheetah.py:190-198):
// u_bar <- [u_bar[:, r:], last column repeated r times];  x0 <- x_bar[:, r].
__global__ void mpc_shift_kernel(Dev d, int r) {
  const int b = blockIdx.x;
  const int m = d.m, T = d.T, n = d.n;
  double* u = d.u_bar + (size_t)b * T * m;
  {
    // in-place shift of the [T][m] tape by r rows, all threads: chunks of blockDim elements in
    // increasing order; a chunk's sources lie at or after its own destinations, so "read all, sync,
    // write all" per chunk never loses a value.  Rows t >= T - r repeat the (original) last row.
    __shared__ double last[32];
    for (int j = threadIdx.x; j < m; j += blockDim.x) last[j] = u[(size_t)(T - 1) * m + j];
    __syncthreads();
    const int total = T * m, shift = r * m;
    for (int base = 0; base < total; base += blockDim.x) {
      const int i = base + threadIdx.x;
      double v = 0.0;
      if (i < total) v = (i + shift < total) ? u[i + shift] : last[i % m];
      __syncthreads();
      if (i < total) u[i] = v;
      __syncthreads();
    }
  }
  double* x0 = const_cast<double*>(d.x0) + (size_t)b * n;
  const double* xr = d.x_bar + ((size_t)b * d.N + r) * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) x0[j] = xr[j];
}

// The resolve step of the MPC loops for the trajectories finish_iter_kernel flagged: the control
// tape moves up by r steps and is padded with its last column, x0 <- x_bar[:, r], the target
// advances (mini_cheetah.py:151-156,193-198).  K, kappa, x_bar, dV stay as they are: the next
// Solve() on the same object starts from them (stale-state semantics).
__global__ void mpc_rearm_kernel(Dev d) {
  const int b = blockIdx.x;
  if (!d.rearm[b]) return;
  const int m = d.m, T = d.T, n = d.n, r = d.mpc_replan;
  double* u = d.u_bar + (size_t)b * T * m;
  {
    // in-place shift of the [T][m] tape by r rows, all threads: chunks of blockDim elements in
    // increasing order; a chunk's sources lie at or after its own destinations, so "read all, sync,
    // write all" per chunk never loses a value.  Rows t >= T - r repeat the (original) last row.
    __shared__ double last[32];
    for (int j = threadIdx.x; j < m; j += blockDim.x) last[j] = u[(size_t)(T - 1) * m + j];
    __syncthreads();
    const int total = T * m, shift = r * m;
    for (int base = 0; base < total; base += blockDim.x) {
      const int i = base + threadIdx.x;
      double v = 0.0;
      if (i < total) v = (i + shift < total) ? u[i + shift] : last[i % m];
      __syncthreads();
      if (i < total) u[i] = v;
      __syncthreads();
    }
  }
  double* x0 = const_cast<double*>(d.x0) + (size_t)b * n;
  double* xnom = const_cast<double*>(d.x_nom) + (size_t)b * n;
  const double* xr = d.x_bar + ((size_t)b * d.N + r) * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    x0[j] = xr[j];
    if (d.mpc_target_adv) xnom[j] += d.mpc_target_adv[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    d.rearm[b] = 0;
    d.rearm_mark[b] = 1;
  }
}

// SetInitialState / SetInitialGuess of an MPC loop that keeps its buffers on the host
// (mini_cheetah.py:148-149 inside the loop of :190-201), from a staging copy already on the device:
// x0 <- xs, u_bar <- us, except for the trajectories the device itself re-armed since the last call
// (their x0 / tape are the newer ones: the staged rows were read back before the shift).
__global__ void apply_staged_kernel(Dev d, const double* xs, const double* us) {
  const int b = blockIdx.x;
  if (d.rearm_mark[b]) {
    __syncthreads();
    if (threadIdx.x == 0) d.rearm_mark[b] = 0;
    return;
  }
  const size_t nu = (size_t)d.T * d.m;
  const double* src = us + (size_t)b * nu;
  double* dst = d.u_bar + (size_t)b * nu;
  for (size_t i = threadIdx.x; i < nu; i += blockDim.x) dst[i] = src[i];
  double* x0 = const_cast<double*>(d.x0) + (size_t)b * d.n;
  for (int i = threadIdx.x; i < d.n; i += blockDim.x) x0[i] = xs[(size_t)b * d.n + i];
}

// reset per-iteration line-search state
__global__ void begin_iter_kernel(Dev d, int force_all) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) {
    d.counters[0] = 0;
    d.counters[1] = 0;
  }
  if (b >= d.B) return;
  if (force_all) {
    d.active[b] = 1;
  }
  d.resolved[b] = 0;
  d.acc[b] = -1;
}

// improvement = L - L_new; L = L_new; stop when improvement <= delta (ilqr.py:692,706-708)
__global__ void finish_iter_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  if (!d.active[b]) return;
  const double imp = d.L[b] - d.L_new[b];
  d.improvement[b] = imp;
  d.L[b] = d.L_new[b];
  d.iters[b] += 1;
  if (imp > d.delta) {
    atomicAdd(&d.counters[1], 1);
  } else if (d.mpc_replan > 0) {
    // Solve() returned; the MPC loop calls it again on the same object for the next resolve
    d.L_conv[b] = d.L_new[b];
    d.resolves[b] += 1;
    d.rearm[b] = 1;
    d.L[b] = INFINITY;               // ilqr.py:681-682
    d.improvement[b] = INFINITY;
    atomicAdd(&d.counters[1], 1);
  } else {
    d.L_conv[b] = d.L_new[b];
    d.active[b] = 0;
    d.status[b] = 1;
  }
}

// =============================================================================================
// K3  keypoint selection (ilqr.py:417-539) -> ascending list per trajectory.
// =============================================================================================
// get_keypoints_set_interval, ilqr.py:417-432
// (CTA per trajectory, a thread per keypoint: the list is written coalesced)
__global__ void kp_set_interval_kernel(Dev d) {
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  int* list = d.kplist + (size_t)b * d.T;
  const int cnt = (d.N - 2) / d.minN + 1;           // t = 0, minN, 2 minN, ... < N - 1
  for (int i = threadIdx.x; i < cnt; i += blockDim.x)
    list[i] = (i == cnt - 1) ? d.N - 2 : i * d.minN;   // the last keypoint is replaced by N - 2 (:428-430)
  if (threadIdx.x == 0) d.kpcount[b] = cnt;
}
// calc_jerk_profile + threshold test, ilqr.py:470-486,454-455: flag[b][t] = any_i jerk[t,i] > thr
__global__ void jerk_flag_kernel(Dev d) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)d.B * (d.N - 3)) return;
  const int b = (int)(gid / (d.N - 3)), t = (int)(gid % (d.N - 3));
  if (!d.active[b]) return;
  const int n = d.n, dof = n / 2;
  const double* x = d.x_bar + ((size_t)b * d.N + t) * n + dof;
  bool any = false;
  for (int i = 0; i < dof; ++i) {
    const double a1 = x[2 * n + i] - x[n + i];
    const double a2 = x[n + i] - x[i];
    any = any || ((a1 - a2) > d.jerk_thr);
  }
  d.flag[(size_t)b * d.N + t] = any ? 1 : 0;
}
// get_keypoints_adaptive_jerk counter scan, ilqr.py:447-466
__global__ void jerk_scan_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || !d.active[b]) return;
  int* list = d.kplist + (size_t)b * d.T;
  const unsigned char* flag = d.flag + (size_t)b * d.N;
  int cnt = 0, counter = 0;
  list[cnt++] = 0;
  for (int t = 0; t < d.N - 3; ++t) {
    counter += 1;
    if (counter >= d.minN && flag[t]) {
      if (cnt < d.T) list[cnt++] = t;
      counter = 0;
    }
    if (counter >= d.maxN) {
      if (cnt < d.T) list[cnt++] = t;
      counter = 0;
    }
  }
  if (list[cnt - 1] != d.N - 2) list[cnt - 1] = d.N - 2;
  d.kpcount[b] = cnt;
}
// segment lookup for the interpolation: t in [kp_i, kp_{i+1}) -> (kp_i, kp_{i+1}); else -1.
// One CTA per trajectory, a thread per step: binary search in the (ascending) keypoint list, so the
// two index rows are written coalesced (a thread per trajectory scanning its row took 0.12 ms at C5).
__global__ void segments_kernel(Dev d) {
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  const int* list = d.kplist + (size_t)b * d.T;
  int* ss = d.seg_s + (size_t)b * d.T;
  int* se = d.seg_e + (size_t)b * d.T;
  const int cnt = d.kpcount[b];
  for (int t = threadIdx.x; t < d.T; t += blockDim.x) {
    int s0 = -1, e0 = 0;
    if (cnt >= 2 && t >= list[0] && t < list[cnt - 1]) {
      int lo = 0, hi = cnt - 1;          // invariant: list[lo] <= t < list[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] <= t) lo = mid;
        else hi = mid;
      }
      s0 = list[lo];
      e0 = list[hi];
    }
    ss[t] = s0;
    if (s0 >= 0) se[t] = e0;
  }
}

// ---- iterativeError (ilqr.py:488-593), level-synchronous ---------------------------------
__global__ void ie_init_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || !d.active[b]) return;
  for (int t = 0; t < d.N; ++t) d.done[(size_t)b * d.N + t] = 0;
  d.segs[0][(size_t)b * 2 * d.T + 0] = 0;
  d.segs[0][(size_t)b * 2 * d.T + 1] = d.N - 2;
  d.nseg[0][b] = 1;
  d.nseg[1][b] = 0;
}
// collect the indices whose Jacobian this level needs (start/mid/end of every segment longer
// than minN), ilqr.py:557-577
__global__ void ie_collect_kernel(Dev d, int cur) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || !d.active[b]) return;
  const int ns = d.nseg[cur][b];
  const int* segs = d.segs[cur] + (size_t)b * 2 * d.T;
  int* ev = d.evallist + (size_t)b * d.T;
  unsigned char* done = d.done + (size_t)b * d.N;
  int ne = 0;
  for (int i = 0; i < ns; ++i) {
    const int s = segs[2 * i], e = segs[2 * i + 1];
    if (e - s <= d.minN) continue;
    const int idx[3] = {s, (s + e) / 2, e};
    for (int k = 0; k < 3; ++k)
      if (!done[idx[k]]) {
        done[idx[k]] = 1;
        ev[ne++] = idx[k];
      }
  }
  d.evalcount[b] = ne;
}
// check_one_matrix_error (ilqr.py:579-591) for every segment of the level; split the bad ones
// (ilqr.py:517-521).  One block per trajectory.
__global__ void ie_check_kernel(Dev d, int cur) {
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  const int n = d.n, nn = n * n;
  const int ns = d.nseg[cur][b];
  const int* segs = d.segs[cur] + (size_t)b * 2 * d.T;
  int* out = d.segs[cur ^ 1] + (size_t)b * 2 * d.T;
  __shared__ double red[32];
  __shared__ int nout;
  if (threadIdx.x == 0) nout = 0;
  __syncthreads();
  for (int i = 0; i < ns; ++i) {
    const int s = segs[2 * i], e = segs[2 * i + 1];
    if (e - s <= d.minN) continue;
    const int mid = (s + e) / 2;
    const double* fs = d.fx + ((size_t)b * d.T + s) * nn;
    const double* fm = d.fx + ((size_t)b * d.T + mid) * nn;
    const double* fe = d.fx + ((size_t)b * d.T + e) * nn;
    double acc = 0.0;
    for (int k = threadIdx.x; k < nn; k += blockDim.x) {
      const double lin = (fe[k] + fs[k]) / 2.0;
      const double df = lin - fm[k];
      acc += df * df;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (blockDim.x + 31) / 32; ++w) tot += red[w];
      if (tot / (2.0 * n) > d.err_thr) {
        out[2 * nout] = s;
        out[2 * nout + 1] = mid;
        out[2 * nout + 2] = mid;
        out[2 * nout + 3] = e;
        nout += 2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    d.nseg[cur ^ 1][b] = nout;
    d.nseg[cur][b] = 0;
  }
}
// keypoints = every index whose Jacobian was evaluated, ascending (ilqr.py:535-537)
__global__ void ie_finish_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || !d.active[b]) return;
  int* list = d.kplist + (size_t)b * d.T;
  int cnt = 0;
  for (int t = 0; t < d.N - 1; ++t)
    if (d.done[(size_t)b * d.N + t]) list[cnt++] = t;
  d.kpcount[b] = cnt;
  d.evalcount[b] = 0;
}

// =============================================================================================
// K4  dynamics linearization at listed timesteps (replaces _calc_dynamics_partials,
// ilqr.py:233-272, and the keypoint loop :409-411).  Forward-mode AD with the n+m seed
// directions spread over the G lanes of a group (K per lane); lane L owns directions
// g = k*G + L so that stores of one Jacobian row are contiguous across lanes.
// =============================================================================================
#ifndef DDP_LIN_MINB
#define DDP_LIN_MINB 2
#endif
template <class Model, int G, int K, int PASSES>
__global__ void __launch_bounds__(128, (Model::n > 8 ? DDP_LIN_MINB : 1)) linearize_kernel(Dev d, const int* list, const int* count) {
  constexpr int n = Model::n, m = Model::m;
  typedef Dual<K> D;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gtid / G;
  const int lane = (int)(gtid % G);
  if (item >= (size_t)d.B * d.T) return;
  const int b = (int)(item / d.T), i = (int)(item % d.T);
  if (!d.active[b] || i >= count[b]) return;
  const int t = list[(size_t)b * d.T + i];
  const double* xp = d.x_bar + ((size_t)b * d.N + t) * n;
  const double* up = d.u_bar + ((size_t)b * d.T + t) * m;
  double* fx = d.fx + ((size_t)b * d.T + t) * n * n;
  double* fu = d.fu + ((size_t)b * d.T + t) * n * m;
  // the n+m seed directions are covered in PASSES sweeps of G*K directions each: fewer
  // directions per lane keep the dual state in registers at the price of recomputing values
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {
    D xs[n], us[m], out[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      xs[j].v = xp[j];
#pragma unroll
      for (int k = 0; k < K; ++k) xs[j].d[k] = ((pass * K + k) * G + lane == j) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int j = 0; j < m; ++j) {
      us[j].v = up[j];
#pragma unroll
      for (int k = 0; k < K; ++k) us[j].d[k] = ((pass * K + k) * G + lane == n + j) ? 1.0 : 0.0;
    }
    Model::template step<D>(xs, us, out, model_params<Model>(d));
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int g = (pass * K + k) * G + lane;
      if (g < n) {
#pragma unroll
        for (int r = 0; r < n; ++r) fx[r * n + g] = out[r].d[k];
      } else if (g < n + m) {
#pragma unroll
        for (int r = 0; r < n; ++r) fu[r * m + (g - n)] = out[r].d[k];
      }
    }
  }
}

// =============================================================================================
// K5  interpolate_derivatives (ilqr.py:596-621): fx_j = fx_s + (fx_e - fx_s)*(j-s)/(e-s)
// =============================================================================================
__global__ void interp_kernel(Dev d) {
  const int b = blockIdx.x / d.T, t = blockIdx.x % d.T;
  if (!d.active[b]) return;
  const int s = d.seg_s[(size_t)b * d.T + t];
  if (s < 0 || s == t) return;
  const int e = d.seg_e[(size_t)b * d.T + t];
  const double w = (double)(t - s), den = (double)(e - s);
  const int nn = d.n * d.n, nm = d.n * d.m;
  const double* fs = d.fx + ((size_t)b * d.T + s) * nn;
  const double* fe = d.fx + ((size_t)b * d.T + e) * nn;
  double* ft = d.fx + ((size_t)b * d.T + t) * nn;
  for (int k = threadIdx.x; k < nn; k += blockDim.x) ft[k] = fs[k] + (fe[k] - fs[k]) * w / den;
  const double* us = d.fu + ((size_t)b * d.T + s) * nm;
  const double* ue = d.fu + ((size_t)b * d.T + e) * nm;
  double* ut = d.fu + ((size_t)b * d.T + t) * nm;
  for (int k = threadIdx.x; k < nm; k += blockDim.x) ut[k] = us[k] + (ue[k] - us[k]) * w / den;
}

// =============================================================================================
// K6  backward Riccati sweep (replaces _backward_pass, ilqr.py:623-667, with the cost
// partials of :161-206).  One CTA per trajectory; Vxx, Vx and the per-step tiles live in
// shared memory; sequential over t = N-2 .. 0.
// =============================================================================================
template <int n, int m>
struct BwdSmem {
  double Vxx[n * n], W[n * n], Fx[n * n];
  double Fu[n * m], Wu[n * m], Qux[m * n], Kt[m * n];
  double Quu[m * m], Inv[m * m];
  double Vx[n], Qx[n], xb[n];
  double Qu[m], g[m], kap[m], ub[m];
  double Q2[n * n], R2[m * m];   // 2 Q, 2 R (lxx, luu), read once
};

// explicit inverse of the m x m matrix A into Inv: in-place Gauss-Jordan with partial pivoting,
// one warp, lane r keeps row r in registers.  Rows are never physically swapped: the pivot lane
// of every column is remembered and the permutation is undone when the result is stored.  The
// pivot search is two REDUX max ops on the IEEE bit pattern of |a| plus a ballot (lowest lane
// wins ties); the pivot row is broadcast with shuffles.  Stands in for np.linalg.inv(Quu)
// (ilqr.py:655).  m <= 32.
template <int m>
__device__ void invert_warp(const double* A, double* Inv) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  double a[m];
#pragma unroll
  for (int j = 0; j < m; ++j) a[j] = (lane < m) ? A[lane * m + j] : 0.0;
  bool used = (lane >= m);  // lanes that may no longer serve as pivot rows
  int mycol = -1;           // pivot column this lane's row was used for
  int rc[m];                // pivot lane of each column (warp-uniform)
#pragma unroll
  for (int c = 0; c < m; ++c) {
    const unsigned long long key = used ? 0ull : (unsigned long long)__double_as_longlong(fabs(a[c])) + 1ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(full, hi);
    const unsigned ml = __reduce_max_sync(full, hi == mh ? lo : 0u);
    const int arg = __ffs(__ballot_sync(full, hi == mh && lo == ml)) - 1;
    rc[c] = arg;
    const bool isp = (lane == arg);
    if (isp) {
      used = true;
      mycol = c;
    }
    const double piv = 1.0 / __shfl_sync(full, a[c], arg);
    const double f = a[c];
#pragma unroll
    for (int j = 0; j < m; ++j) {
      const double sj = (j == c) ? piv : __shfl_sync(full, a[j], arg) * piv;
      if (isp) a[j] = sj;
      else a[j] = (j == c) ? (-f * sj) : fma(-f, sj, a[j]);
    }
  }
  // stored S[r][c'] holds inverse entry (mycol(r), rc[c']); A may alias Inv: every lane loaded
  // its row before the first shuffle, so all reads precede these writes
  __syncwarp();
  if (mycol >= 0) {
#pragma unroll
    for (int j = 0; j < m; ++j) Inv[mycol * m + rc[j]] = a[j];
  }
  __syncwarp();
}

template <class Model, int NT>
__global__ void __launch_bounds__(NT) backward_kernel(Dev d) {
  constexpr int n = Model::n, m = Model::m;
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<n, m>& s = *reinterpret_cast<BwdSmem<n, m>*>(smem_raw);
  const int tid = threadIdx.x;
  const int N = d.N, T = d.T;
  const double* Q = d.Q;
  const double* R = d.R;
  const double* Qf = d.Qf;
  const double* xnom = d.x_nom + (size_t)b * n;

  // Vx, Vxx <- terminal cost partials at x_bar[:, -1]            (ilqr.py:638, 203-204)
  {
    const double* xl = d.x_bar + ((size_t)b * N + (N - 1)) * n;
    for (int i = tid; i < n * n; i += NT) s.Vxx[i] = 2.0 * Qf[i];
    for (int i = tid; i < n; i += NT) {
      double a = 0.0, c = 0.0;
      for (int j = 0; j < n; ++j) {
        a = fma(2.0 * Qf[i * n + j], xl[j], a);
        c = fma(2.0 * xnom[j], Qf[j * n + i], c);
      }
      s.Vx[i] = a - c;
    }
  }
  __syncthreads();

  // The tiles of a step (fx, fu, x_bar, u_bar) are fetched one step ahead into registers (each
  // thread its own elements): for these small shapes a step is a few hundred flops and the L2
  // latency of a blocking load would be a third of it.
  constexpr int PFX = (n * n + NT - 1) / NT, PFU = (n * m + NT - 1) / NT, PX = (n + NT - 1) / NT, PU = (m + NT - 1) / NT;
  double pfx[PFX], pfu[PFU], pxb[PX], pub[PU];
  auto fetch_tiles = [&](int tt) {
    const double* gfx = d.fx + ((size_t)b * T + tt) * n * n;
    const double* gfu = d.fu + ((size_t)b * T + tt) * n * m;
#pragma unroll
    for (int k = 0; k < PFX; ++k) pfx[k] = (tid + k * NT < n * n) ? gfx[tid + k * NT] : 0.0;
#pragma unroll
    for (int k = 0; k < PFU; ++k) pfu[k] = (tid + k * NT < n * m) ? gfu[tid + k * NT] : 0.0;
#pragma unroll
    for (int k = 0; k < PX; ++k) pxb[k] = (tid + k * NT < n) ? d.x_bar[((size_t)b * N + tt) * n + tid + k * NT] : 0.0;
#pragma unroll
    for (int k = 0; k < PU; ++k) pub[k] = (tid + k * NT < m) ? d.u_bar[((size_t)b * T + tt) * m + tid + k * NT] : 0.0;
  };
  fetch_tiles(T - 1);
  for (int i = tid; i < n * n; i += NT) s.Q2[i] = 2.0 * Q[i];
  for (int i = tid; i < m * m; i += NT) s.R2[i] = 2.0 * R[i];

  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int k = 0; k < PFX; ++k)
      if (tid + k * NT < n * n) s.Fx[tid + k * NT] = pfx[k];
#pragma unroll
    for (int k = 0; k < PFU; ++k)
      if (tid + k * NT < n * m) s.Fu[tid + k * NT] = pfu[k];
#pragma unroll
    for (int k = 0; k < PX; ++k)
      if (tid + k * NT < n) s.xb[tid + k * NT] = pxb[k];
#pragma unroll
    for (int k = 0; k < PU; ++k)
      if (tid + k * NT < m) s.ub[tid + k * NT] = pub[k];
    __syncthreads();
    if (t > 0) fetch_tiles(t - 1);   // consumed at the top of the next step
    // W = Vxx fx, Wu = Vxx fu
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, k = idx % n;
      double a = 0.0;
#pragma unroll 4
      for (int j = 0; j < n; ++j) a = fma(s.Vxx[i * n + j], s.Fx[j * n + k], a);
      s.W[idx] = a;
    }
    for (int idx = tid; idx < n * m; idx += NT) {
      const int i = idx / m, r = idx % m;
      double a = 0.0;
#pragma unroll 4
      for (int j = 0; j < n; ++j) a = fma(s.Vxx[i * n + j], s.Fu[j * m + r], a);
      s.Wu[idx] = a;
    }
    // Qx = lx + fx' Vx ; Qu = lu + fu' Vx                         (ilqr.py:651-652,180-181)
    for (int k = tid; k < n; k += NT) {
      double a = 0.0, c = 0.0;
      if (d.diag_cost) {
        a = s.Q2[k * n + k] * s.xb[k];
        c = 2.0 * xnom[k] * Q[k * n + k];
      } else {
        for (int j = 0; j < n; ++j) {
          a = fma(s.Q2[k * n + j], s.xb[j], a);
          c = fma(2.0 * xnom[j], Q[j * n + k], c);
        }
      }
      double q = a - c;
      for (int i = 0; i < n; ++i) q = fma(s.Fx[i * n + k], s.Vx[i], q);
      s.Qx[k] = q;
    }
    for (int r = tid; r < m; r += NT) {
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(s.R2[r * m + j], s.ub[j], a);
      for (int i = 0; i < n; ++i) a = fma(s.Fu[i * m + r], s.Vx[i], a);
      s.Qu[r] = a;
    }
    __syncthreads();
    // Qxx = lxx + fx' W (into Vxx) ; Qux = fu' W ; Quu = luu + fu' Wu   (ilqr.py:653-656)
    for (int idx = tid; idx < n * n; idx += NT) {
      const int k = idx / n, l = idx % n;
      double a = s.Q2[idx];
#pragma unroll 4
      for (int i = 0; i < n; ++i) a = fma(s.Fx[i * n + k], s.W[i * n + l], a);
      s.Vxx[idx] = a;
    }
    for (int idx = tid; idx < m * n; idx += NT) {
      const int r = idx / n, l = idx % n;
      double a = 0.0;
#pragma unroll 4
      for (int i = 0; i < n; ++i) a = fma(s.Fu[i * m + r], s.W[i * n + l], a);
      s.Qux[idx] = a;
    }
    for (int idx = tid; idx < m * m; idx += NT) {
      const int r = idx / m, q = idx % m;
      double a = s.R2[idx];
      for (int i = 0; i < n; ++i) a = fma(s.Fu[i * m + r], s.Wu[i * m + q], a);
      s.Quu[idx] = (r == q) ? a + d.quu_reg : a;
    }
    __syncthreads();
    if constexpr (m == 1) {                                      // ilqr.py:655
      if (tid == 0) s.Inv[0] = 1.0 / s.Quu[0];                   // what the pivoting inverse does for a scalar
    } else {
      if (tid < 32) invert_warp<m>(s.Quu, s.Inv);
    }
    __syncthreads();
    // kappa = Quu^-1 Qu ; K = Quu^-1 Qux ; g = Qu' Quu^-1          (ilqr.py:659-663)
    for (int r = tid; r < m; r += NT) {
      double a = 0.0, c = 0.0;
      for (int j = 0; j < m; ++j) {
        a = fma(s.Inv[r * m + j], s.Qu[j], a);
        c = fma(s.Qu[j], s.Inv[j * m + r], c);
      }
      s.kap[r] = a;
      s.g[r] = c;
    }
    for (int idx = tid; idx < m * n; idx += NT) {
      const int r = idx / n, l = idx % n;
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(s.Inv[r * m + j], s.Qux[j * n + l], a);
      s.Kt[idx] = a;
    }
    __syncthreads();
    // outputs + value update                                       (ilqr.py:659-667)
    double* gK = d.K + ((size_t)b * T + t) * m * n;
    for (int idx = tid; idx < m * n; idx += NT) gK[idx] = s.Kt[idx];
    for (int r = tid; r < m; r += NT) d.kappa[((size_t)b * T + t) * m + r] = s.kap[r];
    if (tid == 0) {
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(s.g[j], s.Qu[j], a);
      d.dV[(size_t)b * T + t] = a;
    }
    for (int k = tid; k < n; k += NT) {
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(s.g[j], s.Qux[j * n + k], a);
      s.Vx[k] = s.Qx[k] - a;
    }
    for (int idx = tid; idx < n * n; idx += NT) {
      const int k = idx / n, l = idx % n;
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(s.Qux[j * n + k], s.Kt[j * n + l], a);
      s.Vxx[idx] -= a;
    }
    __syncthreads();
  }
}

// =============================================================================================
// K6 for the smallest models (n <= 4, m = 1: pendulum, acrobot, cart-pole): the same sweep with a
// THREAD per trajectory and every matrix in registers.  A step is ~40-250 flops: the CTA version
// above spends most of its microsecond per step on seven barriers and shared-memory round trips
// between phases of a handful of operations each.  Same formulas and the same order of operations
// as backward_kernel (results are bit-identical); the next step's tiles are pulled towards L1
// while the current step computes.  Used for B <= 8 (per-thread tile loads do not coalesce across
// trajectories); DDP_SMALL_BACKWARD=cta selects the CTA kernel.
// =============================================================================================
template <class Model>
__global__ void __launch_bounds__(32) backward_small_kernel(Dev d) {
  constexpr int n = Model::n, m = Model::m;
  static_assert(n <= 4 && m == 1, "register-resident sweep: n <= 4, scalar control");
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || !d.active[b]) return;
  const int N = d.N, T = d.T;
  const double* Q = d.Q;
  const double* Qf = d.Qf;
  const double* xnom = d.x_nom + (size_t)b * n;
  const double R2 = 2.0 * d.R[0];
  const bool diag = d.diag_cost != 0;
  double Q2[n][n], xn[n];
#pragma unroll
  for (int i = 0; i < n; ++i) {
    xn[i] = xnom[i];
#pragma unroll
    for (int j = 0; j < n; ++j) Q2[i][j] = 2.0 * Q[i * n + j];
  }
  // Vx, Vxx <- terminal cost partials at x_bar[:, -1]            (ilqr.py:638, 203-204)
  double Vxx[n][n], Vx[n];
  {
    const double* xl = d.x_bar + ((size_t)b * N + (N - 1)) * n;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double a = 0.0, c = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        Vxx[i][j] = 2.0 * Qf[i * n + j];
        a = fma(2.0 * Qf[i * n + j], xl[j], a);
        c = fma(2.0 * xn[j], Qf[j * n + i], c);
      }
      Vx[i] = a - c;
    }
  }
  for (int t = T - 1; t >= 0; --t) {
    const double* gfx = d.fx + ((size_t)b * T + t) * n * n;
    const double* gfu = d.fu + ((size_t)b * T + t) * n * m;
    const double* gxb = d.x_bar + ((size_t)b * N + t) * n;
    double fx[n][n], fu[n], xb[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      fu[i] = gfu[i];
      xb[i] = gxb[i];
#pragma unroll
      for (int j = 0; j < n; ++j) fx[i][j] = gfx[i * n + j];
    }
    const double ub = d.u_bar[(size_t)b * T + t];
    if (t > 0) {   // next step's tiles towards L1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gfx - n * n));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gfu - n));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gxb - n));
    }
    // W = Vxx fx, Wu = Vxx fu
    double W[n][n], Wu[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < n; ++j) a = fma(Vxx[i][j], fx[j][k], a);
        W[i][k] = a;
      }
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) a = fma(Vxx[i][j], fu[j], a);
      Wu[i] = a;
    }
    // Qx = lx + fx' Vx ; Qu = lu + fu' Vx                         (ilqr.py:651-652,180-181)
    double Qx[n], Qu;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      double a = 0.0, c = 0.0;
      if (diag) {
        a = Q2[k][k] * xb[k];
        c = 2.0 * xn[k] * Q[k * n + k];
      } else {
#pragma unroll
        for (int j = 0; j < n; ++j) {
          a = fma(Q2[k][j], xb[j], a);
          c = fma(2.0 * xn[j], Q[j * n + k], c);
        }
      }
      double q = a - c;
#pragma unroll
      for (int i = 0; i < n; ++i) q = fma(fx[i][k], Vx[i], q);
      Qx[k] = q;
    }
    {
      double a = 0.0;
      a = fma(R2, ub, a);
#pragma unroll
      for (int i = 0; i < n; ++i) a = fma(fu[i], Vx[i], a);
      Qu = a;
    }
    // Qxx = lxx + fx' W (into Vxx) ; Qux = fu' W ; Quu = luu + fu' Wu   (ilqr.py:653-656)
    double Qux[n], Quu;
#pragma unroll
    for (int k = 0; k < n; ++k)
#pragma unroll
      for (int l = 0; l < n; ++l) {
        double a = Q2[k][l];
#pragma unroll
        for (int i = 0; i < n; ++i) a = fma(fx[i][k], W[i][l], a);
        Vxx[k][l] = a;
      }
#pragma unroll
    for (int l = 0; l < n; ++l) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) a = fma(fu[i], W[i][l], a);
      Qux[l] = a;
    }
    {
      double a = R2;
#pragma unroll
      for (int i = 0; i < n; ++i) a = fma(fu[i], Wu[i], a);
      Quu = a + d.quu_reg;
    }
    const double Inv = 1.0 / Quu;                                  // ilqr.py:655
    // kappa = Quu^-1 Qu ; K = Quu^-1 Qux ; g = Qu' Quu^-1 ; dV = g Qu      (ilqr.py:659-663)
    const double kap = fma(Inv, Qu, 0.0), g = fma(Qu, Inv, 0.0);
    double Kt[n];
#pragma unroll
    for (int l = 0; l < n; ++l) Kt[l] = fma(Inv, Qux[l], 0.0);
    double* gK = d.K + ((size_t)b * T + t) * m * n;
#pragma unroll
    for (int l = 0; l < n; ++l) gK[l] = Kt[l];
    d.kappa[(size_t)b * T + t] = kap;
    d.dV[(size_t)b * T + t] = fma(g, Qu, 0.0);
    // Vx = Qx - g Qux ; Vxx = Qxx - Qux' K                          (ilqr.py:666-667)
#pragma unroll
    for (int k = 0; k < n; ++k) {
      Vx[k] = Qx[k] - fma(g, Qux[k], 0.0);
#pragma unroll
      for (int l = 0; l < n; ++l) Vxx[k][l] -= fma(Qux[k], Kt[l], 0.0);
    }
  }
}

// =============================================================================================
// fp64 pipe microbenchmarks (bench.py states the fp64 roofline next to the HBM one)
// =============================================================================================
__global__ void peak_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double b = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void peak_dmma_kernel(double* out, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a = threadIdx.x * 1e-3, b = 0.5;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

}  // namespace ddp
