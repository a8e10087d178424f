// K1 for the quadruped model: closed-loop rollout + cost with 8 lanes per line-search candidate.
//
// Same contract as rollout_kernel (kernels.cuh): the body of _linesearch
// (/root/reference/ilqr.py:306-327) and _calc_dynamics (:208-231) for candidates
// eps_table[ls_base .. ls_base + per_traj) of every unresolved trajectory.  A rollout is 199
// dependent steps, so its time is the latency of one step times N-1 as long as every candidate
// is resident at once; the generic kernel spends ~19 k cycles per step with 4 lanes per
// candidate.  Here a candidate gets 8 lanes (8192 candidates = 2048 warps = one resident wave at
// <= 146 registers) and every serial piece of a step is split over them:
//   * feedback u = u_bar - eps kappa - K (x - x_bar): lane pair p owns rows p, p+4, p+8, each
//     lane one half of the columns, halves combined with one shuffle;
//   * the 15 sines / cosines a substep needs (3 per leg + roll, pitch, yaw): two per lane,
//     shared by shuffle; the leg dynamics proper (Quadruped::leg_trig, the same template the
//     linearization differentiates) on the even lane of each pair; loads summed with the
//     butterfly (leg0 + leg1) + (leg2 + leg3) like Quadruped::substep;
//   * the integrator and the running cost: state / control entries dealt round-robin.
// The state lives in shared memory (36 doubles per candidate).  Values equal
// Quadruped::step<double>() up to the association of the feedback sum; costs are summed per
// lane and combined once at the end.
#pragma once
#include "kernels.cuh"

namespace ddp {

constexpr int kRqLanes = 8;        // lanes per candidate
constexpr int kRqCands = 8;        // candidates per CTA (64 threads)

struct RqCandSmem {
  double x[36];     // current state
  double vn[18];    // staged v+ of the substep
  double acc[18];   // accelerations of the substep
  double u[12];     // controls of the step
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}

// SHARED: first line-search round with per_traj == kRqCands: the 8 candidates of a CTA belong to
// one trajectory, so its per-step operands (K_t, x_bar_t, u_bar_t, kappa_t: 3936 bytes) are
// staged once per CTA into a double buffer with cp.async, one step ahead, instead of being
// fetched by every candidate through L1.
constexpr int kRqStage = 432 + 36 + 12 + 12;   // doubles per staged step

template <bool SHARED>
__global__ void __launch_bounds__(kRqLanes * kRqCands, 7)
rollout_quad8_kernel(Dev d, int ls_base, int per_traj, int n_items) {
  typedef Quadruped Qd;
  constexpr int n = 36, m = 12;
  __shared__ RqCandSmem sm[kRqCands];
  __shared__ __align__(16) double stage_buf[SHARED ? 2 : 1][SHARED ? kRqStage : 2];
  const int cand = threadIdx.x >> 3, lane = threadIdx.x & 7;
  const int item = blockIdx.x * kRqCands + cand;
  bool alive = true;   // SHARED: a finished candidate keeps taking part in the CTA barriers
  if (item >= n_items) {
    if (!SHARED) return;
    alive = false;
  }
  int b, ai;
  if (SHARED) {        // launcher guarantees ls_base == 0, per_traj == kRqCands
    b = blockIdx.x;
    ai = cand;
    if (!d.active[b] || d.resolved[b]) return;   // uniform over the CTA
  } else if (ls_base == 0) {
    b = item / per_traj;
    ai = item % per_traj;
    if (!d.active[b] || d.resolved[b]) return;
  } else {
    b = d.unres[item / per_traj];
    ai = item % per_traj;
  }
  const int c = ls_base + ai;
  if (alive && c >= d.n_eps) {
    if (lane == 0) {
      d.Lc[item] = nan("");
      d.Ec[item] = 0.0;
    }
    if (!SHARED) return;
    alive = false;
  }
  const int wl = threadIdx.x & 31, gbase = wl & ~7;
  const unsigned mask = 0xFFu << gbase;
  RqCandSmem& s = sm[cand];
  const double* p = d.params;
  const int sub_n = (int)p[1];
  const double h = p[0] / sub_n;
  const double eps = d.eps_table[min(c, d.n_eps - 1)];
  const double ecoef = -eps * (1.0 - eps / 2.0);
  const int N = d.N, T = d.T;
  const double* xnom = d.x_nom + (size_t)b * n;
  double* xo = d.xc + (size_t)item * N * n;
  double* uo = d.uc + (size_t)item * T * m;
  const bool diag = d.diag_cost != 0;

  // lane roles
  const int leg = lane >> 1, odd = lane & 1;
  const double sx = (leg < 2) ? 1.0 : -1.0, sd = (leg & 1) ? 1.0 : -1.0;
  const int hh = odd;                 // column half of the feedback rows
  const int prow = lane >> 1;         // feedback rows prow, prow + 4, prow + 8

  for (int j = lane; j < n; j += kRqLanes) {
    const double v = d.x0[(size_t)b * n + j];
    s.x[j] = v;
    if (alive) xo[j] = v;
  }
  __syncwarp(mask);
  auto stage = [&](int t, int buf) {   // SHARED: all threads of the CTA
    char* dst = reinterpret_cast<char*>(stage_buf[buf]);
    const char* gK = reinterpret_cast<const char*>(d.K + ((size_t)b * d.T + t) * m * n);
    for (int ch = threadIdx.x; ch < 216; ch += kRqLanes * kRqCands) cp_async16(dst + 16 * ch, gK + 16 * ch);
    const int q = threadIdx.x;
    if (q < 18) cp_async16(dst + 3456 + 16 * q, reinterpret_cast<const char*>(d.x_bar + ((size_t)b * d.N + t) * n) + 16 * q);
    else if (q < 24) cp_async16(dst + 3744 + 16 * (q - 18), reinterpret_cast<const char*>(d.u_bar + ((size_t)b * d.T + t) * m) + 16 * (q - 18));
    else if (q < 30) cp_async16(dst + 3840 + 16 * (q - 24), reinterpret_cast<const char*>(d.kappa + ((size_t)b * d.T + t) * m) + 16 * (q - 24));
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (SHARED) stage(0, 0);

  // cost weights of this lane's entries (diagonal costs)
  double L = 0.0, E = 0.0;
  bool ok = true;
  for (int t = 0; t < T; ++t) {
    if (SHARED) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();   // step t staged and visible; everybody is done with the other buffer
      if (t + 1 < T) stage(t + 1, (t + 1) & 1);
      if (!alive) continue;
    }
    const double* Kt = SHARED ? stage_buf[t & 1] : d.K + ((size_t)b * T + t) * m * n;
    const double* xb = SHARED ? stage_buf[t & 1] + 432 : d.x_bar + ((size_t)b * N + t) * n;
    const double* ub = SHARED ? stage_buf[t & 1] + 468 : d.u_bar + ((size_t)b * T + t) * m;
    const double* kp = SHARED ? stage_buf[t & 1] + 480 : d.kappa + ((size_t)b * T + t) * m;
    if (!SHARED && t + 1 < T) {   // pull the next step's gain half-rows towards L1 while this step computes
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const char* pr = reinterpret_cast<const char*>(Kt + (size_t)m * n + (size_t)(prow + 4 * i) * n + 18 * hh);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 128));
      }
    }
    // ---- u_t = u_bar_t - eps*kappa_t - K_t (x_t - x_bar_t)            (ilqr.py:313) ----------
    {
      double dx[18];
#pragma unroll
      for (int j = 0; j < 18; ++j) dx[j] = s.x[18 * hh + j] - xb[18 * hh + j];
      double a3[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double* Kr = Kt + (size_t)(prow + 4 * i) * n + 18 * hh;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < 18; j += 2) {
          a0 = fma(Kr[j], dx[j], a0);
          a1 = fma(Kr[j + 1], dx[j + 1], a1);
        }
        a3[i] = a0 + a1;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double other = __shfl_xor_sync(mask, a3[i], 1);
        const double lo = odd ? other : a3[i], hi = odd ? a3[i] : other;   // first half + second half
        const int r = prow + 4 * i;
        const double u = ub[r] - eps * kp[r] - (lo + hi);
        if (!odd) s.u[r] = u;
      }
    }
    __syncwarp(mask);
    // ---- running cost uses the pre-step state                         (ilqr.py:325) ----------
    {
      double sacc = 0.0;
      if (diag) {
        for (int j = lane; j < n; j += kRqLanes) {
          const double e = s.x[j] - xnom[j];
          sacc = fma(d.Q[j * n + j] * e, e, sacc);
        }
        for (int r = lane; r < m; r += kRqLanes) sacc = fma(d.R[r * m + r] * s.u[r], s.u[r], sacc);
      } else {
        for (int j = lane; j < n; j += kRqLanes) {
          double row = 0.0;
          for (int k = 0; k < n; ++k) row = fma(d.Q[j * n + k], s.x[k] - xnom[k], row);
          sacc = fma(s.x[j] - xnom[j], row, sacc);
        }
        for (int r = lane; r < m; r += kRqLanes) {
          double row = 0.0;
          for (int k = 0; k < m; ++k) row = fma(d.R[r * m + k], s.u[k], row);
          sacc = fma(s.u[r], row, sacc);
        }
      }
      L += sacc;
      for (int r = lane; r < m; r += kRqLanes) uo[(size_t)t * m + r] = s.u[r];
    }
    const double ua = s.u[3 * leg], uh = s.u[3 * leg + 1], uk = s.u[3 * leg + 2];
    // ---- x_{t+1} = f(x_t, u_t)                                          (ilqr.py:316) ----------
    for (int it = 0; it < sub_n; ++it) {
      // two sincos per lane: even lane of leg l: abad, hip + knee; odd lane: hip, one base angle
      const double qh_ = s.x[7 + 3 * leg];
      const double ang1 = odd ? qh_ : s.x[6 + 3 * leg];
      const double ang2 = odd ? s.x[3 + (leg < 3 ? leg : 0)] : (qh_ + s.x[8 + 3 * leg]);
      double s1, c1, s2, c2;
      sincos_(ang1, &s1, &c1);
      sincos_(ang2, &s2, &c2);
      Qd::BasePose<double> B;
      B.sr = __shfl_sync(mask, s2, gbase + 1);
      B.cr = __shfl_sync(mask, c2, gbase + 1);
      B.sp = __shfl_sync(mask, s2, gbase + 3);
      B.cp = __shfl_sync(mask, c2, gbase + 3);
      const double sy = __shfl_sync(mask, s2, gbase + 5), cy = __shfl_sync(mask, c2, gbase + 5);
      Qd::base_pose_trig(sy, cy, B);
      const double sh = __shfl_sync(mask, s1, wl | 1), ch = __shfl_sync(mask, c1, wl | 1);
      double f[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      double vb[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) vb[k] = s.x[18 + k];
      if (!odd) {
        Qd::LegOut<double> o;
        Qd::leg_trig(sx, sd, s1, c1, sh, ch, s2, c2, s.x[24 + 3 * leg], s.x[25 + 3 * leg], s.x[26 + 3 * leg], ua, uh, uk,
                     s.x[2], vb, B, p, o);
        f[0] = o.Fx; f[1] = o.Fy; f[2] = o.Fz; f[3] = o.Tx; f[4] = o.Ty; f[5] = o.Tz;
        s.acc[6 + 3 * leg] = o.a0;
        s.acc[7 + 3 * leg] = o.a1;
        s.acc[8 + 3 * leg] = o.a2;
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(mask, f[k], 1);   // odd lanes hold 0
        f[k] += __shfl_xor_sync(mask, f[k], 2);
        f[k] += __shfl_xor_sync(mask, f[k], 4);
      }
      if (lane == 0) {
        double accb[18];
        Qd::base_acc(f[0], f[1], f[2], f[3], f[4], f[5], vb, p, accb);
#pragma unroll
        for (int k = 0; k < 6; ++k) s.acc[k] = accb[k];
      }
      __syncwarp(mask);
      // semi-implicit Euler (Quadruped::integrate): v+ first, then q+ = q + h N(q) v+
      for (int i = lane; i < 18; i += kRqLanes) s.vn[i] = s.x[18 + i] + h * s.acc[i];
      __syncwarp(mask);
      {
        const double tp = B.sp / B.cp;
        const double wyz = B.sr * s.vn[4] + B.cr * s.vn[5];
        double qn[3];
        int cnt = 0;
        for (int i = lane; i < 18; i += kRqLanes, ++cnt) {
          double rate = s.vn[i];
          if (i == 3) rate = s.vn[3] + tp * wyz;
          if (i == 4) rate = B.cr * s.vn[4] - B.sr * s.vn[5];
          if (i == 5) rate = wyz / B.cp;
          qn[cnt] = s.x[i] + h * rate;
        }
        __syncwarp(mask);   // every read of the old state done
        cnt = 0;
        for (int i = lane; i < 18; i += kRqLanes, ++cnt) {
          s.x[i] = qn[cnt];
          s.x[18 + i] = s.vn[i];
        }
      }
      __syncwarp(mask);
    }
    bool fin = true;
    for (int j = lane; j < n; j += kRqLanes) fin = fin && isfinite(s.x[j]);
    if (!__all_sync(mask, fin)) {  // the reference gets a RuntimeError from Drake: L = inf, stop (:317-323)
      ok = false;
      if (!SHARED) break;
      alive = false;   // keep taking part in the CTA barriers
      continue;
    }
    E += ecoef * d.dV[(size_t)b * T + t];                      //  (ilqr.py:326)
    for (int j = lane; j < n; j += kRqLanes) xo[(size_t)(t + 1) * n + j] = s.x[j];
  }
  // terminal cost                                                   (ilqr.py:327)
  if (ok) {
    double sacc = 0.0;
    if (diag) {
      for (int j = lane; j < n; j += kRqLanes) {
        const double e = s.x[j] - xnom[j];
        sacc = fma(d.Qf[j * n + j] * e, e, sacc);
      }
    } else {
      for (int j = lane; j < n; j += kRqLanes) {
        double row = 0.0;
        for (int k = 0; k < n; ++k) row = fma(d.Qf[j * n + k], s.x[k] - xnom[k], row);
        sacc = fma(s.x[j] - xnom[j], row, sacc);
      }
    }
    L += sacc;
    L += __shfl_xor_sync(mask, L, 1);
    L += __shfl_xor_sync(mask, L, 2);
    L += __shfl_xor_sync(mask, L, 4);
  } else {
    L = INFINITY;
  }
  if (lane == 0 && (alive || !ok)) {
    d.Lc[item] = L;
    d.Ec[item] = E;
  }
}

}  // namespace ddp
