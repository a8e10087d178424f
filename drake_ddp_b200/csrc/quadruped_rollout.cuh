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
//
// QUAT = true is the same kernel for QuadrupedQuat, the reference's own state layout
// (mini_cheetah.py:41-57, n = 37: q = [quat, pos, joints], v = [w_world, v_lin, joint rates]): the
// legs, the butterfly, the feedback split (19 + 18 columns) and the cost are shared; only the base
// differs -- rotation from the normalised quaternion instead of three sines / cosines, the
// angular velocity taken to the body frame for the legs and Euler's equations, the angular
// acceleration taken back to the world frame, and the quaternion rate instead of Euler rates.
#pragma once
#include "kernels.cuh"

namespace ddp {

constexpr int kRqLanes = 8;        // lanes per candidate
constexpr int kRqCands = 8;        // candidates per CTA (64 threads)

struct RqCandSmem {
  double x[38];     // current state (36 or 37 entries)
  double vn[18];    // staged v+ of the substep
  double acc[18];   // accelerations of the substep
  double u[12];     // controls of the step
  double cw[12][8]; // per-lane constants of the diagonal cost: weights of its 5 state and 2 control entries, 5 targets
                    // (24 registers that the 128-register build would spill and reload every step)
};

// state layout of the two base parameterisations and the layout of one staged step
template <bool QUAT>
struct RqLayout {
  static constexpr int n = QUAT ? 37 : 36;
  static constexpr int NQ = QUAT ? 19 : 18;   // first velocity entry
  static constexpr int JQ = QUAT ? 7 : 6;     // first joint angle
  static constexpr int JV = NQ + 6;           // first joint rate
  static constexpr int PZ = QUAT ? 6 : 2;     // base height
  static constexpr int HL = (n + 1) / 2;      // columns of the first half of a feedback row (18 / 19)
  // staged operands of one step (doubles): K_t | x_bar_t (padded to even) | u_bar_t | kappa_t
  static constexpr int XB = 12 * n, UB = XB + ((n + 1) & ~1), KP = UB + 12, STAGE = KP + 12;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}

// SHARED: first line-search round with per_traj == kRqCands: the 8 candidates of a CTA belong to
// one trajectory, so its per-step operands (K_t, x_bar_t, u_bar_t, kappa_t: 3936 bytes) are
// staged once per warp (4 candidates) into a double buffer with cp.async, one step ahead,
// instead of being fetched by every candidate through L1.

#ifdef DDP_ROLL_PROFILE
__device__ long long g_roll_prof[16];
#define ROLL_TICK(i)                                 \
  do {                                               \
    if (rprof) {                                     \
      const long long now_ = clock64();              \
      racc[i] += now_ - rlast;                       \
      rlast = now_;                                  \
    }                                                \
  } while (0)
#else
#define ROLL_TICK(i)
#endif

// MINB: resident CTAs per SM the register budget is sized for.  7 (128 registers: four warps per
// sub-partition's 16 K registers) keeps a 1024-trajectory round in one wave; a launch of at most
// 3 CTAs per SM (small batches, the strong split over 8 GPUs, later line-search rounds) takes the
// 3-CTA build, which spills nothing on the serial path.
template <bool SHARED, bool QUAT, int MINB>
__global__ void __launch_bounds__(kRqLanes * kRqCands, MINB)
rollout_quad8_kernel(Dev d, int ls_base, int per_traj, int n_items) {
  typedef Quadruped Qd;
  typedef RqLayout<QUAT> Ly;
  constexpr int n = Ly::n, m = 12, NQ = Ly::NQ, JQ = Ly::JQ, JV = Ly::JV, HL = Ly::HL;
  constexpr int kRqStage = Ly::STAGE;
  __shared__ RqCandSmem sm[kRqCands];
  // [warp][buffer][operands of one step]: every warp (4 candidates) keeps its own copy, so only
  // warp-level synchronisation is needed
  __shared__ __align__(16) double stage_all[SHARED ? 2 : 1][SHARED ? 2 : 1][SHARED ? kRqStage : 2];
  double (*stage_buf)[SHARED ? kRqStage : 2] = stage_all[SHARED ? (threadIdx.x >> 5) : 0];
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
  const double* p = d.pm;   // constant bank (Dev::pm)
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
  auto stage = [&](int t, int buf) {   // SHARED: all threads of the warp
    char* dst = reinterpret_cast<char*>(stage_buf[buf]);
    const char* gK = reinterpret_cast<const char*>(d.K + ((size_t)b * d.T + t) * m * n);
#pragma unroll
    for (int i = 0; i < (6 * n + 31) / 32; ++i) {   // unrolled: a loop branch per copy sits on the step's issue path
      const int ch = wl + 32 * i;
      if (ch < 6 * n) cp_async16(dst + 16 * ch, gK + 16 * ch);
    }
    const char* gx = reinterpret_cast<const char*>(d.x_bar + ((size_t)b * d.N + t) * n);
    const char* gu = reinterpret_cast<const char*>(d.u_bar + ((size_t)b * d.T + t) * m);
    const char* gk = reinterpret_cast<const char*>(d.kappa + ((size_t)b * d.T + t) * m);
    const int q = wl;
    if (!QUAT) {
      if (q < 18) cp_async16(dst + 8 * Ly::XB + 16 * q, gx + 16 * q);
      else if (q < 24) cp_async16(dst + 8 * Ly::UB + 16 * (q - 18), gu + 16 * (q - 18));
      else if (q < 30) cp_async16(dst + 8 * Ly::KP + 16 * (q - 24), gk + 16 * (q - 24));
    } else {
      // a 37-entry state row starts 16-byte aligned only at every other step: 8-byte copies
      for (int ch = q; ch < n; ch += 32) cp_async8(dst + 8 * (Ly::XB + ch), gx + 8 * ch);
      if (q >= 8 && q < 14) cp_async16(dst + 8 * Ly::UB + 16 * (q - 8), gu + 16 * (q - 8));
      else if (q >= 14 && q < 20) cp_async16(dst + 8 * Ly::KP + 16 * (q - 14), gk + 16 * (q - 14));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (SHARED) stage(0, 0);

  // diagonal cost: weights and targets of this lane's entries (states lane + 8k, controls lane + 8k)
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int j = lane + kRqLanes * k;
    s.cw[k][lane] = (j < n) ? d.Q[j * n + j] : 0.0;
    s.cw[5 + k][lane] = (j < n) ? xnom[j] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int r = lane + kRqLanes * k;
    s.cw[10 + k][lane] = (r < m) ? d.R[r * m + r] : 0.0;
  }
  // reciprocal mass / inertia of the base acceleration this lane computes (lanes 0..5), fetched once:
  // a ternary over global loads inside the step compiles to branches
  const double rcp_l = (lane < 3) ? p[20] : (lane == 3) ? p[21] : (lane == 4) ? p[22] : p[23];
  const bool l_b0 = (lane & 1) != 0, l_b1 = (lane & 2) != 0, l_b2 = (lane & 4) != 0;
  double L = 0.0, E = 0.0;
  bool ok = true;
#ifdef DDP_ROLL_PROFILE
  const bool rprof = SHARED && (blockIdx.x == 3 && threadIdx.x == 0);
  long long racc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long rlast = clock64();
#endif
  for (int t = 0; t < T; ++t) {
    ROLL_TICK(0);
    if (SHARED) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();      // step t staged and visible; the whole warp is done with the other buffer
      if (t + 1 < T) stage(t + 1, (t + 1) & 1);
      if (!alive) continue;
    }
    ROLL_TICK(1);
    const double* Kt = SHARED ? stage_buf[t & 1] : d.K + ((size_t)b * T + t) * m * n;
    const double* xb = SHARED ? stage_buf[t & 1] + Ly::XB : d.x_bar + ((size_t)b * N + t) * n;
    const double* ub = SHARED ? stage_buf[t & 1] + Ly::UB : d.u_bar + ((size_t)b * T + t) * m;
    const double* kp = SHARED ? stage_buf[t & 1] + Ly::KP : d.kappa + ((size_t)b * T + t) * m;
    const double dv_t = d.dV[(size_t)b * T + t];   // issued early, consumed at the end of the step
    if (!SHARED && t + 1 < T) {   // pull the next step's gain half-rows towards L1 while this step computes
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const char* pr = reinterpret_cast<const char*>(Kt + (size_t)m * n + (size_t)(prow + 4 * i) * n + HL * hh);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 128));
      }
    }
    // ---- u_t = u_bar_t - eps*kappa_t - K_t (x_t - x_bar_t)            (ilqr.py:313) ----------
    {
      // this lane's half of the columns: [HL * hh, HL * hh + HL), clipped to n (19 + 18 at n = 37)
      double dx[HL];
#pragma unroll
      for (int j = 0; j < HL; ++j) {
        const bool in = (2 * HL == n) || (HL * hh + j < n);
        dx[j] = in ? (s.x[HL * hh + j] - xb[HL * hh + j]) : 0.0;
      }
      double a3[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double* Kr = Kt + (size_t)(prow + 4 * i) * n + HL * hh;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < HL; j += 2) {
          const bool in0 = (2 * HL == n) || (HL * hh + j < n);
          a0 = fma(in0 ? Kr[j] : 0.0, dx[j], a0);
          if (j + 1 < HL) {
            const bool in1 = (2 * HL == n) || (HL * hh + j + 1 < n);
            a1 = fma(in1 ? Kr[j + 1] : 0.0, dx[j + 1], a1);
          }
        }
        a3[i] = a0 + a1;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double other = __shfl_xor_sync(mask, a3[i], 1);
        const double lo = odd ? other : a3[i], hi = odd ? a3[i] : other;   // first half + second half
        const int r = prow + 4 * i;
        double u = ub[r] - eps * kp[r] - (lo + hi);
        if (d.u_min) u = fmin(fmax(u, d.u_min[r]), d.u_max[r]);   // extension, off by default
        if (!odd) s.u[r] = u;
      }
    }
    __syncwarp(mask);
    ROLL_TICK(2);
    // ---- running cost uses the pre-step state                         (ilqr.py:325) ----------
    {
      double sacc = 0.0;
      if (diag) {
        double pt[7];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int j = lane + kRqLanes * k;
          const double e = ((j < n) ? s.x[j] : 0.0) - s.cw[5 + k][lane];
          pt[k] = (s.cw[k][lane] * e) * e;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int r = lane + kRqLanes * k;
          const double uu = (r < m) ? s.u[r] : 0.0;
          pt[5 + k] = (s.cw[10 + k][lane] * uu) * uu;
        }
        sacc = ((pt[0] + pt[1]) + (pt[2] + pt[3])) + ((pt[4] + pt[5]) + pt[6]);
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
#pragma unroll
      for (int k = 0; k < (m + kRqLanes - 1) / kRqLanes; ++k) {
        const int r = lane + kRqLanes * k;
        if (r < m) uo[(size_t)t * m + r] = s.u[r];
      }
    }
    ROLL_TICK(3);
    const double ua = s.u[3 * leg], uh = s.u[3 * leg + 1], uk = s.u[3 * leg + 2];
    // ---- x_{t+1} = f(x_t, u_t)                                          (ilqr.py:316) ----------
    bool fin_step = true;
    for (int it = 0; it < sub_n; ++it) {
      // two sincos per lane: even lane of leg l: abad, hip + knee; odd lane: hip, one base angle
      // (QUAT: no base angles; the odd lanes' second pair is not used)
      const double qh_ = s.x[JQ + 1 + 3 * leg];
      const double ang1 = odd ? qh_ : s.x[JQ + 3 * leg];
      const double ang2 = (!QUAT && odd) ? s.x[3 + (leg < 3 ? leg : 0)] : (qh_ + s.x[JQ + 2 + 3 * leg]);
      double s1, c1, s2, c2;
      sincos_(ang1, &s1, &c1);
      sincos_(ang2, &s2, &c2);
      ROLL_TICK(4);
      Qd::BasePose<double> B;
      double vb[6];   // [world linear velocity | body angular velocity]: what the legs want
      double icp = 1.0, tp = 0.0;
      if (!QUAT) {
        B.sr = __shfl_sync(mask, s2, gbase + 1);
        B.cr = __shfl_sync(mask, c2, gbase + 1);
        B.sp = __shfl_sync(mask, s2, gbase + 3);
        B.cp = __shfl_sync(mask, c2, gbase + 3);
        const double sy = __shfl_sync(mask, s2, gbase + 5), cy = __shfl_sync(mask, c2, gbase + 5);
        Qd::base_pose_trig(sy, cy, B);
#pragma unroll
        for (int k = 0; k < 6; ++k) vb[k] = s.x[18 + k];
        icp = 1.0 / B.cp;   // a division is ~25 dependent instructions: started here it runs under the leg chain
        tp = B.sp * icp;
      } else {
        // rotation matrix of the normalised quaternion (QuadrupedQuat::step), every lane
        const double q0 = s.x[0], q1 = s.x[1], q2 = s.x[2], q3 = s.x[3];
        const double inn = inv_sqrt_(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        const double a = q0 * inn, b_ = q1 * inn, c_ = q2 * inn, d_ = q3 * inn;
        B.R00 = 1.0 - 2.0 * (c_ * c_ + d_ * d_); B.R01 = 2.0 * (b_ * c_ - a * d_); B.R02 = 2.0 * (b_ * d_ + a * c_);
        B.R10 = 2.0 * (b_ * c_ + a * d_); B.R11 = 1.0 - 2.0 * (b_ * b_ + d_ * d_); B.R12 = 2.0 * (c_ * d_ - a * b_);
        B.R20 = 2.0 * (b_ * d_ - a * c_); B.R21 = 2.0 * (c_ * d_ + a * b_); B.R22 = 1.0 - 2.0 * (b_ * b_ + c_ * c_);
        B.sr = 0.0; B.cr = 1.0; B.sp = 0.0; B.cp = 1.0;
        const double w0 = s.x[19], w1 = s.x[20], w2 = s.x[21];
        vb[0] = s.x[22]; vb[1] = s.x[23]; vb[2] = s.x[24];
        vb[3] = B.R00 * w0 + B.R10 * w1 + B.R20 * w2;
        vb[4] = B.R01 * w0 + B.R11 * w1 + B.R21 * w2;
        vb[5] = B.R02 * w0 + B.R12 * w1 + B.R22 * w2;
      }
      const double sh = __shfl_sync(mask, s1, wl | 1), ch = __shfl_sync(mask, c1, wl | 1);
      double f[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (!odd) {
        Qd::LegOut<double> o;
        Qd::leg_trig(sx, sd, s1, c1, sh, ch, s2, c2, s.x[JV + 3 * leg], s.x[JV + 1 + 3 * leg], s.x[JV + 2 + 3 * leg], ua, uh,
                     uk, s.x[Ly::PZ], vb, B, p, o);
        f[0] = o.Fx; f[1] = o.Fy; f[2] = o.Fz; f[3] = o.Tx; f[4] = o.Ty; f[5] = o.Tz;
        s.acc[6 + 3 * leg] = o.a0;
        s.acc[7 + 3 * leg] = o.a1;
        s.acc[8 + 3 * leg] = o.a2;
      }
      ROLL_TICK(5);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(mask, f[k], 1);   // odd lanes hold 0
        f[k] += __shfl_xor_sync(mask, f[k], 2);
        f[k] += __shfl_xor_sync(mask, f[k], 4);
      }
      ROLL_TICK(6);
      // base accelerations: entry k on lane k (numerator and reciprocal mass / inertia are selected
      // branch-free)
      const double Ix = p[3], Iy = p[4], Iz = p[5], grav = p[19];
      const double n3 = f[3] - (Iz - Iy) * vb[4] * vb[5];
      const double n4 = f[4] - (Ix - Iz) * vb[5] * vb[3];
      const double n5 = f[5] - (Iy - Ix) * vb[3] * vb[4];
      if (!QUAT) {   // Quadruped::base_acc: [linear | body angular]
        // entry `lane` of (f0, f1, f2, n3, n4, n5) by a tree of selects on the lane's bits: the nested
        // ternary over six values compiles to a branch ladder on the serial path of the substep
        const double s01 = l_b0 ? f[1] : f[0], s23 = l_b0 ? n3 : f[2], s45 = l_b0 ? n5 : n4;
        const double s03 = l_b1 ? s23 : s01;
        const double num = l_b2 ? s45 : s03;
        double a = num * rcp_l;
        if (lane == 2) a -= grav;
        if (lane < 6) s.acc[lane] = a;
      } else {       // QuadrupedQuat::step: [world angular = R (body angular) | linear]
        const double ab0 = n3 * p[21], ab1 = n4 * p[22], ab2 = n5 * p[23];
        const double r0 = (lane == 0) ? B.R00 : (lane == 1) ? B.R10 : B.R20;
        const double r1 = (lane == 0) ? B.R01 : (lane == 1) ? B.R11 : B.R21;
        const double r2 = (lane == 0) ? B.R02 : (lane == 1) ? B.R12 : B.R22;
        const double arot = r0 * ab0 + r1 * ab1 + r2 * ab2;
        double alin = ((lane == 3) ? f[0] : (lane == 4) ? f[1] : f[2]) * p[20];
        if (lane == 5) alin -= grav;
        if (lane < 6) s.acc[lane] = (lane < 3) ? arot : alin;
      }
      __syncwarp(mask);
      ROLL_TICK(7);
      // semi-implicit Euler: v+ first, then q+ = q + h N(q) v+; velocity entries lane, lane + 8,
      // lane + 16; the base attitude rows get the new angular velocity by shuffle
      {
        double vn[3], qn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = lane + kRqLanes * k;
          vn[k] = (i < 18) ? (s.x[NQ + i] + h * s.acc[i]) : 0.0;
        }
        if (!QUAT) {
          const double w3 = __shfl_sync(mask, vn[0], gbase + 3), w4 = __shfl_sync(mask, vn[0], gbase + 4),
                       w5 = __shfl_sync(mask, vn[0], gbase + 5);
          const double wyz = B.sr * w4 + B.cr * w5;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int i = lane + kRqLanes * k;
            double rate = vn[k];
            if (k == 0) {
              if (lane == 3) rate = w3 + tp * wyz;
              if (lane == 4) rate = B.cr * w4 - B.sr * w5;
              if (lane == 5) rate = wyz * icp;
            }
            qn[k] = (i < 18) ? (s.x[i] + h * rate) : 0.0;
            fin_step = fin_step && isfinite(qn[k]) && isfinite(vn[k]);
          }
          __syncwarp(mask);   // every read of the old state done
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int i = lane + kRqLanes * k;
            if (i < 18) {
              s.x[i] = qn[k];
              s.x[18 + i] = vn[k];
            }
          }
        } else {
          // qdot = 0.5 (0, w_W) (x) q with the NEW angular velocity; quaternion entry e on lane e
          const double w0 = __shfl_sync(mask, vn[0], gbase + 0), w1 = __shfl_sync(mask, vn[0], gbase + 1),
                       w2 = __shfl_sync(mask, vn[0], gbase + 2);
          const double q0 = s.x[0], q1 = s.x[1], q2 = s.x[2], q3 = s.x[3];
          const double e0 = q0 + (0.5 * h) * (-(w0 * q1) - w1 * q2 - w2 * q3);
          const double e1 = q1 + (0.5 * h) * (w0 * q0 + w1 * q3 - w2 * q2);
          const double e2 = q2 + (0.5 * h) * (w1 * q0 + w2 * q1 - w0 * q3);
          const double e3 = q3 + (0.5 * h) * (w2 * q0 + w0 * q2 - w1 * q1);
          const double e01 = l_b0 ? e1 : e0, e23 = l_b0 ? e3 : e2;
          const double qe = l_b1 ? e23 : e01;
          if (lane < 4) fin_step = fin_step && isfinite(qe);
          // position / joint entry that goes with velocity entry i >= 3: q index i + 1
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int i = lane + kRqLanes * k;
            qn[k] = (i >= 3 && i < 18) ? (s.x[i + 1] + h * vn[k]) : 0.0;
            fin_step = fin_step && isfinite(qn[k]) && isfinite(vn[k]);
          }
          __syncwarp(mask);   // every read of the old state done
          if (lane < 4) s.x[lane] = qe;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int i = lane + kRqLanes * k;
            if (i < 18) {
              s.x[NQ + i] = vn[k];
              if (i >= 3) s.x[i + 1] = qn[k];
            }
          }
        }
      }
      __syncwarp(mask);
      ROLL_TICK(8);
    }
    if (!__all_sync(mask, fin_step)) {  // the reference gets a RuntimeError from Drake: L = inf, stop (:317-323)
      ok = false;
      if (!SHARED) break;
      alive = false;   // keep taking part in the warp barriers
      continue;
    }
    E += ecoef * dv_t;                      //  (ilqr.py:326)
#pragma unroll
    for (int k = 0; k < (n + kRqLanes - 1) / kRqLanes; ++k) {
      const int j = lane + kRqLanes * k;
      if (j < n) xo[(size_t)(t + 1) * n + j] = s.x[j];
    }
    ROLL_TICK(9);
  }
#ifdef DDP_ROLL_PROFILE
  if (rprof)
    for (int i = 0; i < 12; ++i) g_roll_prof[i] = racc[i];
#endif
  // terminal cost                                                   (ilqr.py:327)
  if (ok) {
    double sacc = 0.0;
    if (diag) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int j = lane + kRqLanes * k;
        const double e = ((j < n) ? s.x[j] : 0.0) - s.cw[5 + k][lane];
        const double wfk = (j < n) ? d.Qf[j * n + j] : 0.0;
        sacc = fma(wfk * e, e, sacc);
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
