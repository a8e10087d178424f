// K1 for the arm + ball model (config C5, kinova_gen3.py): closed-loop rollout + cost with 8 lanes
// per line-search candidate.
//
// Same contract as rollout_kernel (kernels.cuh): the body of _linesearch
// (/root/reference/ilqr.py:306-327) and _calc_dynamics (:208-231) for candidates
// eps_table[ls_base .. ls_base + per_traj) of every unresolved trajectory.  N = 400 steps of 4
// substeps are 1596 dependent substeps, so a round costs the latency of one substep times 1596; the
// generic kernel evaluates the whole 27-state step on every one of its 4 lanes (255 registers and
// spills: seven sines / cosines, the kinematic chain with all seven Jacobian columns, the contacts)
// and spends ~10 k cycles per substep.  Here a candidate gets 8 lanes, lane i < 7 owns joint i:
//   * feedback u_i = u_bar_i - eps kappa_i - K_i (x - x_bar): row i on lane i, and u_i never
//     leaves the lane (the joint torque balance is lane-local too);
//   * sine / cosine of joint i on lane i, shared through shared memory; the kinematic chain
//     (ArmBall::fk_joint, the model's own template) runs on every lane, which keeps the joint axis
//     and origin of ITS joint only; Jacobian column, torque balance and integration of joint i on
//     lane i; the tool-tip velocity is a butterfly sum over the lanes;
//   * contacts (ArmBall::contacts) and the free ball (ArmBall::ball_integrate) are evaluated by
//     every lane on its own register copy of the ball state: no exchange at all;
//   * running cost and the candidate's state / control tape dealt over the lanes.
// Values equal ArmBall::step<double>() up to the association of the tool-tip velocity sum, of the
// feedback sum and of the cost sums.  Diagonal cost weights only (anything else takes the generic
// kernel).
#pragma once
#include "kernels.cuh"

namespace ddp {

constexpr int kRaLanes = 8;        // lanes per candidate
constexpr int kRaCands = 16;       // candidates per CTA (128 threads)

struct alignas(16) RaCandSmem {
  double dx[28];       // x_t - x_bar_t at the start of the step (feedback operand)
  double sc[14];       // sin, cos of the seven joint angles of the substep
};

// MINB: resident CTAs per SM the register budget is sized for.  A launch of at most 2 CTAs per SM
// (C5: 512 trajectories x 8 candidates = 256 CTAs) takes the 2-CTA build: 255 registers, nothing
// spilled on the serial path; larger launches keep 4 CTAs per SM so that they stay one wave.
template <int MINB>
__global__ void __launch_bounds__(kRaLanes * kRaCands, MINB)
rollout_arm8_kernel(Dev d, int ls_base, int per_traj, int n_items) {
  typedef ArmBall Ab;
  constexpr int n = 27, m = 7;
  __shared__ RaCandSmem sm[kRaCands];
  const int cand = threadIdx.x >> 3, lane = threadIdx.x & 7;
  const int item = blockIdx.x * kRaCands + cand;
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
  const int wl = threadIdx.x & 31, gbase = wl & ~7;
  const unsigned mask = 0xFFu << gbase;
  RaCandSmem& s = sm[cand];
  const double* p = d.pm;   // constant bank (Dev::pm)
  const int sub_n = (int)p[1];
  const double h = p[0] / sub_n;
  const double Ij = p[2], bj = p[3];
  const double* dl = p + 4;
  const double rb = p[12], mb = p[13], bz = p[18];
  const double Ib = 0.4 * mb * rb * rb;
  const double eps = d.eps_table[c];
  const double ecoef = -eps * (1.0 - eps / 2.0);
  const int N = d.N, T = d.T;
  const double* xnom = d.x_nom + (size_t)b * n;
  double* xo = d.xc + (size_t)item * N * n;
  double* uo = d.uc + (size_t)item * T * m;
  const bool joint = lane < 7;
  const int jl = joint ? lane : 0;   // joint of this lane (lane 7: a copy of joint 0 that is never stored)

  // registers: this lane's joint, and the whole ball state (every lane)
  const double* x0 = d.x0 + (size_t)b * n;
  double qj = x0[jl], vj = x0[14 + jl];
  double qb[7], vb[6];
#pragma unroll
  for (int k = 0; k < 7; ++k) qb[k] = x0[7 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) vb[k] = x0[21 + k];
  // this lane's share of the state vector for the cost and the tapes: its joint angle and rate, and
  // ball entries 7 + lane (angle part) and 21 + lane (rate part, lanes 0..5); lane 7 takes no ball entry
  // diagonal cost weights / targets of those entries
  const double wq_j = d.Q[jl * n + jl], wq_v = d.Q[(14 + jl) * n + 14 + jl];
  const double wf_j = d.Qf[jl * n + jl], wf_v = d.Qf[(14 + jl) * n + 14 + jl];
  const double xn_j = xnom[jl], xn_v = xnom[14 + jl];
  const int bq = 7 + jl, bv_ = 21 + (lane < 6 ? lane : 0);
  const double wq_bq = d.Q[bq * n + bq], wq_bv = (lane < 6) ? d.Q[bv_ * n + bv_] : 0.0;
  const double wf_bq = d.Qf[bq * n + bq], wf_bv = (lane < 6) ? d.Qf[bv_ * n + bv_] : 0.0;
  const double xn_bq = xnom[bq], xn_bv = xnom[bv_];
  const double wr = d.R[jl * m + jl];
  // ball entry `k` of this lane out of the register copy (branch-free select)
  auto pick7 = [&](const double* a, int k) {
    double r = a[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) r = (k == i) ? a[i] : r;
    return r;
  };
  auto pick6 = [&](const double* a, int k) {
    double r = a[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) r = (k == i) ? a[i] : r;
    return r;
  };
  // The state entries this lane owns (joint angle / rate, ball entries 7 + lane and 21 + lane): it
  // writes them to the candidate tape and publishes x_t - x_bar_t of them to shared memory for the
  // feedback rows; its four entries of x_bar_{t+1}, its u_bar / kappa entry and dV of the next step
  // are fetched one step ahead into registers, so no global load sits on the serial path but the
  // gain row (prefetched to L1 a step ahead).
  double xbn[4], ubn, kpn, dvn;   // operands of the step about to start
  auto fetch = [&](int t) {
    const double* xb = d.x_bar + ((size_t)b * N + t) * n;
    xbn[0] = xb[jl];
    xbn[1] = xb[14 + jl];
    xbn[2] = xb[7 + jl];
    xbn[3] = xb[21 + (lane < 6 ? lane : 0)];
    if (t < T) {
      ubn = d.u_bar[((size_t)b * T + t) * m + jl];
      kpn = d.kappa[((size_t)b * T + t) * m + jl];
      dvn = d.dV[(size_t)b * T + t];
    }
  };
  auto publish = [&](int t) {   // state at step t (x_bar_t in xbn)
    const double bqv = pick7(qb, lane), bvv = pick6(vb, lane < 6 ? lane : 0);
    if (joint) {
      s.dx[lane] = qj - xbn[0];
      s.dx[14 + lane] = vj - xbn[1];
      s.dx[7 + lane] = bqv - xbn[2];
      xo[(size_t)t * n + lane] = qj;
      xo[(size_t)t * n + 14 + lane] = vj;
      xo[(size_t)t * n + 7 + lane] = bqv;
      if (lane < 6) {
        s.dx[21 + lane] = bvv - xbn[3];
        xo[(size_t)t * n + 21 + lane] = bvv;
      }
    }
  };
  fetch(0);
  publish(0);
  __syncwarp(mask);

  double L = 0.0, E = 0.0;
  bool ok = true;
  for (int t = 0; t < T; ++t) {
    const double ub_t = ubn, kp_t = kpn, dv_t = dvn;
    fetch(t + 1);      // consumed at the end of the step
    // the gain row comes straight from global memory (L2 / L1: the candidates of a trajectory share
    // it); staging it through shared memory with cp.async was measured slower (4.8 -> 5.1 ms)
    const double* Kr = d.K + ((size_t)b * T + t) * m * n + (size_t)jl * n;
    if (t + 1 < T) {   // pull the next step's row (216 bytes, any alignment) towards L1
      const char* pr = reinterpret_cast<const char*>(Kr + (size_t)m * n);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pr));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 128));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 208));
    }
    // ---- u_i = u_bar_i - eps*kappa_i - K_i (x_t - x_bar_t)            (ilqr.py:313) ----------
    double u;
    {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int j = 0; j < n; j += 3) {
        a0 = fma(Kr[j], s.dx[j], a0);
        a1 = fma(Kr[j + 1], s.dx[j + 1], a1);
        a2 = fma(Kr[j + 2], s.dx[j + 2], a2);
      }
      u = ub_t - eps * kp_t - ((a0 + a1) + a2);
      if (d.u_min) u = fmin(fmax(u, d.u_min[jl]), d.u_max[jl]);   // extension, off by default
    }
    // ---- running cost uses the pre-step state                         (ilqr.py:325) ----------
    {
      const double e0 = qj - xn_j, e1 = vj - xn_v;
      const double e2 = pick7(qb, lane) - xn_bq, e3 = pick6(vb, lane < 6 ? lane : 0) - xn_bv;
      double sacc = (wq_bq * e2) * e2 + (wq_bv * e3) * e3;   // lane 7: zero weights would be wrong: masked below
      if (!joint) sacc = 0.0;
      if (joint) sacc += ((wq_j * e0) * e0 + (wq_v * e1) * e1) + (wr * u) * u;
      L += sacc;
      if (joint) uo[(size_t)t * m + lane] = u;
    }
    // ---- x_{t+1} = f(x_t, u_t)                                          (ilqr.py:316) ----------
    for (int it = 0; it < sub_n; ++it) {
      double sn, cs;
      sincos_(qj, &sn, &cs);
      if (joint) {
        s.sc[2 * lane] = sn;
        s.sc[2 * lane + 1] = cs;
      }
      __syncwarp(mask);
      // kinematic chain on every lane; each keeps the origin and axis of its own joint
      Ab::Frame<double> F;
      Ab::fk_init(0.0, bz, F);
      double ox = 0.0, oy = 0.0, oz = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const double2 sc = *reinterpret_cast<const double2*>(&s.sc[2 * i]);
        const bool mine = joint && (i == lane);
        ox = mine ? F.px : ox;
        oy = mine ? F.py : oy;
        oz = mine ? F.pz : oz;
        double axi, ayi, azi;
        Ab::fk_joint(i, sc.x, sc.y, dl[i], F, axi, ayi, azi);
        ax = mine ? axi : ax;
        ay = mine ? ayi : ay;
        az = mine ? azi : az;
      }
      __syncwarp(mask);   // every read of sc done before the next substep overwrites it
      // Jacobian column of this joint (lane 7: axis 0 -> column 0) and the tool-tip velocity
      const double dx = F.px - ox, dy = F.py - oy, dz = F.pz - oz;
      const double Jx = ay * dz - az * dy, Jy = az * dx - ax * dz, Jz = ax * dy - ay * dx;
      double tvx = Jx * vj, tvy = Jy * vj, tvz = Jz * vj;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        tvx += __shfl_xor_sync(mask, tvx, o);
        tvy += __shfl_xor_sync(mask, tvy, o);
        tvz += __shfl_xor_sync(mask, tvz, o);
      }
      Ab::Loads<double> Ld;
      Ab::contacts(F.px, F.py, F.pz, tvx, tvy, tvz, qb[4], qb[5], qb[6], vb, vb + 3, p, Ld);
      const double tau = u + Jx * Ld.ftx + Jy * Ld.fty + Jz * Ld.ftz - bj * vj;
      vj = vj + (h / Ij) * tau;
      qj = qj + h * vj;
      Ab::ball_integrate(qb, vb, Ld, h, mb, Ib);
    }
    bool fin = !joint || (isfinite(qj) && isfinite(vj));
#pragma unroll
    for (int k = 0; k < 7; ++k) fin = fin && isfinite(qb[k]);
#pragma unroll
    for (int k = 0; k < 6; ++k) fin = fin && isfinite(vb[k]);
    if (!__all_sync(mask, fin)) {  // the reference gets a RuntimeError from Drake: L = inf, stop (:317-323)
      ok = false;
      break;
    }
    E += ecoef * dv_t;                      //  (ilqr.py:326)
    publish(t + 1);
    __syncwarp(mask);
  }
  // terminal cost                                                   (ilqr.py:327)
  if (ok) {
    const double e0 = qj - xn_j, e1 = vj - xn_v;
    const double e2 = pick7(qb, lane) - xn_bq, e3 = pick6(vb, lane < 6 ? lane : 0) - xn_bv;
    double sacc = (wf_bq * e2) * e2 + (wf_bv * e3) * e3;
    if (!joint) sacc = 0.0;
    if (joint) sacc += (wf_j * e0) * e0 + (wf_v * e1) * e1;
    L += sacc;
    L += __shfl_xor_sync(mask, L, 1);
    L += __shfl_xor_sync(mask, L, 2);
    L += __shfl_xor_sync(mask, L, 4);
  } else {
    L = INFINITY;
  }
  if (lane == 0) {
    d.Lc[item] = L;
    d.Ec[item] = E;
  }
}

}  // namespace ddp
