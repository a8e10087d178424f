// K4 for the quadruped model: structured linearization fused into one kernel.
//
// Replaces _calc_dynamics_partials (/root/reference/ilqr.py:233-272) over the keypoints
// (:409-411) for Quadruped::step (models.h).  The generic linearize_kernel pushes all n+m = 48
// seed directions through both substeps (4 legs x 48 directions x 2 substeps = 384 dual leg
// evaluations per point).  A leg's loads depend on 16 local inputs only (base height and
// attitude, base twist, its own three joints and rates), so here
//   1. per substep, the 4 x 16 (leg, local direction) pairs are dealt to the 32 lanes of a warp,
//      two directions per lane, and Quadruped::leg_trig is evaluated with Dual<2>: the same
//      templates as the rollout, 64 dual leg evaluations per substep instead of 192; each of
//      the 15 sines / cosines a substep needs is computed by one lane and shared by shuffle;
//   2. every lane scatters its derivatives straight into the substep Jacobian in shared memory
//      (base rows are summed over the legs with the same two-stage butterfly the rollout uses);
//   3. the two substeps are chained on the fp64 tensor pipe: only the velocity rows need a
//      product, and with the position rows of substep 1 folded in (D1q = E1 + h N1 D1v) it is
//      18 x 18 x 48 (90 DMMAs); the position rows follow from q+ = q + h N(q) v+ element-wise.
// One warp per point, all intermediates in that warp's slice of shared memory; fx, fu are
// written once.  Exact derivative of Quadruped::step: checked against the generic AD kernel
// and the host AD (tests/test_gpu_parity.py).  Two substeps only (the model's setting); other
// settings use the generic kernel.
#pragma once
#include "backward_sym.cuh"

namespace ddp {

constexpr int kQfLd = 52;     // leading dimension of the Jacobians in shared memory (52 = 4 mod 16:
                              // conflict-free 8 x 4 and 4 x 8 DMMA operand fetches)
#ifndef QF_WARPS
#define QF_WARPS 4
#endif
constexpr int kQfWarps = QF_WARPS;   // warps (points in flight) per CTA
#ifndef QF_MINB
#define QF_MINB 2
#endif

// local input j (0..15) of leg l -> global column of [fx | fu] (0..47): base height, roll / pitch /
// yaw, base twist, the leg's three joint angles, its three joint rates
__host__ __device__ inline int qf_gcol(int l, int j) {
  if (j == 0) return 2;
  if (j < 4) return 2 + j;
  if (j < 10) return 18 + (j - 4);
  if (j < 13) return 6 + 3 * l + (j - 10);
  return 24 + 3 * l + (j - 13);
}

struct QfWarpSmem {
  double D1v[20 * kQfLd];     // velocity rows of the substep-1 Jacobian d v1/d(q, v, u), 18 x 48 (+2 pad rows)
  double D2v[18 * kQfLd];     // velocity rows of the substep-2 Jacobian, 18 x 48
  double v2s[3 * 48];         // rows 3..5 of the chained velocity rows (Euler-rate coupling)
  double st[3][36];           // x_t, state after substep 1, after substep 2
};

#ifndef QF_MAXNREG
#define QF_MAXNREG 0
#endif
#ifndef QF_ND
#define QF_ND 2      // seed directions a lane carries per pass (2: one pass; 1: two passes, fewer registers)
#endif
#if QF_MAXNREG
__global__ void __maxnreg__(QF_MAXNREG)
#else
__global__ void __launch_bounds__(kQfWarps * 32, QF_MINB)
#endif
quad_fused_kernel(Dev d, const int* list, const int* count, int n_items) {
  typedef Quadruped Qd;
  constexpr int ND = QF_ND, NPASS = 2 / ND;
  typedef Dual<ND> D2;
  constexpr int LD = kQfLd;
  extern __shared__ __align__(16) unsigned char qf_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  QfWarpSmem& s = reinterpret_cast<QfWarpSmem*>(qf_raw)[warp];
  const unsigned full = 0xffffffffu;
  const double* p = d.pm;   // constant bank (Dev::pm)
  const double h = p[0] / 2.0;
  const double Ix = p[3], Iy = p[4], Iz = p[5];
  const int T = d.T;

  // lane -> (leg, pair of local directions) and the global columns of the two directions
  const int leg = lane >> 3, dp = lane & 7;
  const bool shared_dir = dp < 5;   // base directions: every leg contributes to the base rows
  const double sx = (leg < 2) ? 1.0 : -1.0, sd = (leg & 1) ? 1.0 : -1.0;

  // the Jacobian buffers keep their zero pattern across points: only the structural nonzeros
  // are rewritten.  The direct u -> joint acceleration terms are constant.
  for (int i = lane; i < 20 * LD; i += 32) s.D1v[i] = 0.0;
  for (int i = lane; i < 18 * LD; i += 32) s.D2v[i] = 0.0;
  __syncwarp();
  if (lane < 12) {
    s.D1v[(6 + lane) * LD + 36 + lane] = h * p[24 + lane % 3];
    s.D2v[(6 + lane) * LD + 36 + lane] = h * p[24 + lane % 3];
  }
  __syncwarp();

  // Points are taken grid-stride; the indices of the next point are fetched at the top of the
  // current one and its state and controls after the first substep, so that the global-memory
  // latency hides behind the arithmetic (8 warps per SM cannot hide it otherwise).
  // per-lane constants of the gyroscopic entries (lanes 0..5): row 3 + lane / 2, column and angular
  // velocity component by the cross-product pattern, coefficient -h (I_a - I_b) / I_row
  int gy_off = 0, gy_w = 0;
  double gy_coef = 0.0;
  {
    const int rr = 3 + (lane % 6) / 2;
    const int cc_[6] = {22, 23, 23, 21, 21, 22};
    const int ww_[6] = {2, 1, 0, 2, 1, 0};          // which of (w3, w4, w5) multiplies
    const double dI[3] = {Iz - Iy, Ix - Iz, Iy - Ix};
    int cc = 22, ww = 2;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (lane % 6 == k) { cc = cc_[k]; ww = ww_[k]; }
    gy_off = rr * LD + cc;
    gy_w = ww;
    gy_coef = -h * ((rr == 3) ? dI[0] : (rr == 4) ? dI[1] : dI[2]) * p[18 + rr];
  }
  // the angle whose sine / cosine this lane computes: x[ang_a] (+ x[ang_b] for hip + knee)
  const int ang_a = (dp == 0) ? 6 + 3 * leg : (dp <= 2) ? 7 + 3 * leg : (dp == 3) ? 3 : (dp == 4) ? 4 : 5;
  const int ang_b = 8 + 3 * leg;
  const double ang_wb = (dp == 2) ? 1.0 : 0.0;
  const int stride = gridDim.x * kQfWarps;
  auto fetch_idx = [&](int it, int& bb, int& tt) -> bool {
    if (it >= n_items) return false;
    bb = it / T;
    const int i = it % T;
    if (!d.active[bb] || i >= count[bb]) return false;
    tt = list[(size_t)bb * T + i];
    return true;
  };
  struct Pt {
    double x0, x1, ua, uh, uk;
  };
  auto fetch_pt = [&](int bb, int tt) {
    Pt r;
    const double* xp = d.x_bar + ((size_t)bb * d.N + tt) * 36;
    const double* up = d.u_bar + ((size_t)bb * T + tt) * 12;
    r.x0 = xp[lane];
    r.x1 = (lane < 4) ? xp[32 + lane] : 0.0;
    r.ua = up[3 * leg];
    r.uh = up[3 * leg + 1];
    r.uk = up[3 * leg + 2];
    return r;
  };
  int item = blockIdx.x * kQfWarps + warp;
  int b = 0, t = 0;
  bool ok = fetch_idx(item, b, t);
  Pt cur = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (ok) cur = fetch_pt(b, t);
  while (item < n_items) {
    int nb = 0, nt = 0;
    const bool nok = fetch_idx(item + stride, nb, nt);
    Pt nxt = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (!ok) {
      if (nok) nxt = fetch_pt(nb, nt);
    } else {
    const size_t bt = (size_t)b * T + t;
    s.st[0][lane] = cur.x0;
    if (lane < 4) s.st[0][32 + lane] = cur.x1;
    const double ua = cur.ua, uh = cur.uh, uk = cur.uk;
    __syncwarp();

    double trig[2][4];   // sin/cos of roll and pitch at the start of each substep
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const double* xin = s.st[sub];
      double* xout = s.st[sub + 1];
      double* Dv = (sub == 0) ? s.D1v : s.D2v;   // 18 x 48 velocity rows of this substep
      // ---- dual evaluation of this lane's leg along its two local directions (ND per pass) -------
      // one sincos per lane, shared by shuffle: lanes dp = 0..2 of a leg group take abad, hip,
      // hip + knee of their leg, dp = 3..5 roll, pitch, yaw (same values in every group)
      double sn, cs;
      {
        const double ang = xin[ang_a] + ang_wb * xin[ang_b];
        sincos_(ang, &sn, &cs);
      }
      const int gl = lane & ~7;   // first lane of this leg's group
      double f[6], av[3], vbv[6];
#ifdef QF_SKIP_LEG
      for (int pass = 0; pass < 0; ++pass) {
#else
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass) {
#endif
        int jd[ND], gc[ND];      // local directions of this pass and their global columns
#pragma unroll
        for (int e = 0; e < ND; ++e) {
          jd[e] = 2 * dp + pass * ND + e;
          gc[e] = qf_gcol(leg, jd[e]);   // (hoisting these two out of the point loop costs two registers: spills)
        }
        auto seed = [&](double v, int j) {
          D2 r;
          r.v = v;
#pragma unroll
          for (int e = 0; e < ND; ++e) r.d[e] = (j == jd[e]) ? 1.0 : 0.0;
          return r;
        };
        // dual sine / cosine of an angle whose value pair comes from lane `src`; the angle is local
        // input ja (+ jb for hip + knee)
        auto trig_dual = [&](int src, int ja, int jb, D2& sD, D2& cD) {
          const double sv = __shfl_sync(full, sn, src), cv = __shfl_sync(full, cs, src);
          sD.v = sv;
          cD.v = cv;
#pragma unroll
          for (int e = 0; e < ND; ++e) {
            const double we = ((jd[e] == ja) || (jd[e] == jb)) ? 1.0 : 0.0;
            sD.d[e] = cv * we;
            cD.d[e] = -sv * we;
          }
        };
        D2 vb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) vb[k] = seed(xin[18 + k], 4 + k);
        const D2 pz = seed(xin[2], 0);
        Qd::BasePose<D2> B;
        D2 sy, cy;
        trig_dual(gl + 3, 1, -1, B.sr, B.cr);
        trig_dual(gl + 4, 2, -1, B.sp, B.cp);
        trig_dual(gl + 5, 3, -1, sy, cy);
        Qd::base_pose_trig(sy, cy, B);
        trig[sub][0] = B.sr.v; trig[sub][1] = B.cr.v; trig[sub][2] = B.sp.v; trig[sub][3] = B.cp.v;
        D2 sa, ca, sh, ch, sk, ck;
        trig_dual(gl + 0, 10, -1, sa, ca);
        trig_dual(gl + 1, 11, -1, sh, ch);
        trig_dual(gl + 2, 11, 12, sk, ck);
        Qd::LegOut<D2> o;
        Qd::leg_trig(sx, sd, sa, ca, sh, ch, sk, ck, seed(xin[24 + 3 * leg], 13), seed(xin[25 + 3 * leg], 14),
                     seed(xin[26 + 3 * leg], 15), D2(ua), D2(uh), D2(uk), pz, vb, B, p, o);
        // ---- scatter: joint rows of this leg, base rows summed over the legs -------------------
        {
          const D2* ja[3] = {&o.a0, &o.a1, &o.a2};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = 6 + 3 * leg + k;
#pragma unroll
            for (int e = 0; e < ND; ++e) Dv[r * LD + gc[e]] = h * ja[k]->d[e] + ((gc[e] == 18 + r) ? 1.0 : 0.0);
          }
          const D2* fo[6] = {&o.Fx, &o.Fy, &o.Fz, &o.Tx, &o.Ty, &o.Tz};
          const double inv[6] = {p[20], p[20], p[20], p[21], p[22], p[23]};
#pragma unroll
          for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int e = 0; e < ND; ++e) {
              double own = fo[r]->d[e], sum = own;
              sum += __shfl_xor_sync(full, sum, 8);
              sum += __shfl_xor_sync(full, sum, 16);
              const double gs = shared_dir ? sum : own;
              if (!shared_dir || leg == 0) Dv[r * LD + gc[e]] = (h * inv[r]) * gs + ((gc[e] == 18 + r) ? 1.0 : 0.0);
            }
          }
        }
        if (pass == NPASS - 1) {
          f[0] = o.Fx.v; f[1] = o.Fy.v; f[2] = o.Fz.v; f[3] = o.Tx.v; f[4] = o.Ty.v; f[5] = o.Tz.v;
          av[0] = o.a0.v; av[1] = o.a1.v; av[2] = o.a2.v;
#pragma unroll
          for (int k = 0; k < 6; ++k) vbv[k] = vb[k].v;
        }
      }
      // ---- primal state after the substep (same arithmetic as Quadruped::integrate) ------------
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(full, f[k], 8);
        f[k] += __shfl_xor_sync(full, f[k], 16);
      }
      double accb[18];
      Qd::base_acc(f[0], f[1], f[2], f[3], f[4], f[5], vbv, p, accb);
      {
        const int jl = (lane >= 6 && lane < 18) ? (lane - 6) / 3 : 0, jk = (lane >= 6 && lane < 18) ? (lane - 6) % 3 : 0;
        const double a0 = __shfl_sync(full, av[0], 8 * jl), a1 = __shfl_sync(full, av[1], 8 * jl),
                     a2 = __shfl_sync(full, av[2], 8 * jl);
        double acc = (jk == 0) ? a0 : ((jk == 1) ? a1 : a2);
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (lane == k) acc = accb[k];
        const int li = (lane < 18) ? lane : 0;
        const double vnew = xin[18 + li] + h * acc;
        const double w3 = __shfl_sync(full, vnew, 3), w4 = __shfl_sync(full, vnew, 4), w5 = __shfl_sync(full, vnew, 5);
        const double sr = trig[sub][0], cr = trig[sub][1], sp = trig[sub][2], cp = trig[sub][3];
        const double icp = 1.0 / cp;
        const double tp = sp * icp, wyz = sr * w4 + cr * w5;
        double rate = vnew;
        if (lane == 3) rate = w3 + tp * wyz;
        if (lane == 4) rate = cr * w4 - sr * w5;
        if (lane == 5) rate = wyz * icp;
        if (lane < 18) {
          xout[18 + lane] = vnew;
          xout[lane] = xin[lane] + h * rate;
        }
      }
      __syncwarp();
      // gyroscopic terms of the base rotation rows (d/d omega of the omega x I omega term): entry
      // (gy_off) += gy_coef * omega[gy_w], constants per lane (branch-free)
      if (lane < 6) Dv[gy_off] += gy_coef * xin[21 + gy_w];
      __syncwarp();
      if (sub == 0) {
        if (nok) nxt = fetch_pt(nb, nt);   // next point's state and controls: consumed a substep later
      }
    }

#ifndef QF_SKIP_CHAIN
    // ---- chain ----------------------------------------------------------------------------------
    // With Dv2 = [Aq | Av | Au] (18 x (18+18+12)), D1v the velocity rows of substep 1 and
    // D1q = E1 + h N1 D1v its position rows (E1 = [I + h M1 | 0 | 0], N1 the Euler-rate matrix):
    //   v2 = Aq D1q + Av D1v + [0 0 Au] = [Aq E1 | 0 | Au] + (Av + h Aq N1) D1v,
    // so the position rows of D1 are never formed and the product has K = 18: 90 DMMAs.
    // M1 and N1 differ from the identity only in rows / columns 3..5 (roll, pitch, yaw).
    const double* D1v = s.D1v;
    // A' = Av + h Aq N1 is formed in the A-operand fetch of the product (N1 is the identity but for
    // columns 4 and 5, which sit in k-step kk = 1 on lanes tg = 0, 1: per-lane weights, branch-free);
    // Aq E1 then replaces Aq in place (columns 3 and 4 only) for the epilogue.
    const double sr0 = trig[0][0], cr0 = trig[0][1], sp0 = trig[0][2], cp0 = trig[0][3];
    const double icp0 = 1.0 / cp0, tp0 = sp0 * icp0, w40 = s.st[1][22], w50 = s.st[1][23];
    const double wyz0 = sr0 * w40 + cr0 * w50, wr0 = cr0 * w40 - sr0 * w50;
    double acc[3][6][2];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    {
      const double N34 = tp0 * sr0, N35 = tp0 * cr0, N44 = cr0, N45 = -sr0, N54 = sr0 * icp0, N55 = cr0 * icp0;
      const bool c4 = tg == 0, c5 = tg == 1;      // in k-step 1: column 4 + tg
      const double k3 = c4 ? N34 : (c5 ? N35 : 0.0), k4 = c4 ? N44 : (c5 ? N45 : 0.0), k5 = c4 ? N54 : (c5 ? N55 : 0.0);
      const double* pq[3];   // row of [Aq | Av | Au] of this lane's A-fragment row
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) pq[mt] = s.D2v + min(8 * mt + g, 17) * LD;
      const double* pb = D1v + tg * LD + g;
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const bool kin = (kk < 4) || (tg < 2);   // k = 4 kk + tg < 18
        double a[3], bb[6];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const double* row = pq[mt];
          double aqn;
          if (kk == 1) {
            const double own = (c4 || c5) ? 0.0 : row[4 + tg];
            aqn = own + (row[3] * k3 + row[4] * k4 + row[5] * k5);
          } else {
            aqn = kin ? row[4 * kk + tg] : 0.0;
          }
          a[mt] = kin ? (row[18 + 4 * kk + tg] + h * aqn) : 0.0;
        }
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) bb[nt] = kin ? pb[4 * kk * LD + 8 * nt] : 0.0;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int nt = 0; nt < 6; ++nt) dmma(acc[mt][nt], a[mt], bb[nt]);
      }
      __syncwarp();   // every fetch of Aq done
      // Aq E1, in place of Aq (columns 3 and 4 only)
      if (lane < 18) {
        double* row = s.D2v + lane * LD;
        const double a3 = row[3], a4 = row[4], a5 = row[5];
        row[3] = a3 * (1.0 + h * tp0 * wr0) + a4 * (-h * wyz0) + a5 * (h * wr0 * icp0);
        row[4] = a3 * (h * wyz0 * (icp0 * icp0)) + a4 + a5 * (h * wyz0 * sp0 * (icp0 * icp0));
      }
      __syncwarp();
    }
    double* fx = d.fx + bt * 36 * 36;
    double* fu = d.fu + bt * 36 * 12;
    auto store2 = [&](int r, int c, double v0, double v1) {   // rows of [fx | fu], c even
      if (c < 36) *reinterpret_cast<double2*>(fx + r * 36 + c) = make_double2(v0, v1);
      else *reinterpret_cast<double2*>(fu + r * 12 + (c - 36)) = make_double2(v0, v1);
    };
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      const int r = 8 * mt + g;
      if (r < 18) {
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int c = 8 * nt + 2 * tg;
          double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
          if (c < 18 || c >= 36) {   // + [Aq E1 | 0 | Au]
            v0 += s.D2v[r * LD + c];
            v1 += s.D2v[r * LD + c + 1];
          }
          store2(18 + r, c, v0, v1);
          if (r >= 3 && r < 6) {   // Euler-rate rows need rows 3..5 together: stage them
            s.v2s[(r - 3) * 48 + c] = v0;
            s.v2s[(r - 3) * 48 + c + 1] = v1;
          } else {                 // q2 = D1q + h v2 with D1q = I + h D1v on these rows
            const double q0 = ((c == r) ? 1.0 : 0.0) + h * D1v[r * LD + c];
            const double q1 = ((c + 1 == r) ? 1.0 : 0.0) + h * D1v[r * LD + c + 1];
            store2(r, c, q0 + h * v0, q1 + h * v1);
          }
        }
      }
    }
    __syncwarp();
    {
      // rows 3..5: q2 = (I + h M2) D1q + h N2 v2 with the Euler-rate matrices of substep 2
      const double sr = trig[1][0], cr = trig[1][1], sp = trig[1][2], cp = trig[1][3];
      const double icp = 1.0 / cp, tp = sp * icp, wy = s.st[2][22], wz = s.st[2][23];
      const double wyz = sr * wy + cr * wz, wr = cr * wy - sr * wz;
      // the step-independent coefficients of the two Euler-rate maps, hoisted out of the column loop
      const double m33 = 1.0 + h * tp * wr, m34 = h * wyz * (icp * icp), m43 = -h * wyz, m53 = h * wr * icp,
                   m54 = h * wyz * sp * (icp * icp);
      // D1q rows 3..5 = E1 + h N1 D1v with the Euler-rate matrices of substep 1 
      const double sr1 = trig[0][0], cr1 = trig[0][1], sp1 = trig[0][2], cp1 = trig[0][3];
      const double icp1 = 1.0 / cp1, tp1 = sp1 * icp1, wy1 = s.st[1][22], wz1 = s.st[1][23];
      const double wyz1 = sr1 * wy1 + cr1 * wz1, wr1 = cr1 * wy1 - sr1 * wz1;
      const double e33 = 1.0 + h * tp1 * wr1, e34 = h * wyz1 * (icp1 * icp1), e43 = -h * wyz1, e53 = h * wr1 * icp1,
                   e54 = h * wyz1 * sp1 * (icp1 * icp1);
      for (int c = lane; c < 48; c += 32) {
        const double f3 = D1v[3 * LD + c], f4 = D1v[4 * LD + c], f5 = D1v[5 * LD + c];
        double d3 = h * (f3 + tp1 * (sr1 * f4 + cr1 * f5));
        double d4 = h * (cr1 * f4 - sr1 * f5);
        double d5 = h * ((sr1 * f4 + cr1 * f5) * icp1);
        if (c == 3) {
          d3 += e33;
          d4 += e43;
          d5 += e53;
        }
        if (c == 4) {
          d3 += e34;
          d4 += 1.0;
          d5 += e54;
        }
        if (c == 5) d5 += 1.0;
        const double e3 = s.v2s[c], e4 = s.v2s[48 + c], e5 = s.v2s[96 + c];
        const double q3 = m33 * d3 + m34 * d4 + h * (e3 + tp * (sr * e4 + cr * e5));
        const double q4 = m43 * d3 + d4 + h * (cr * e4 - sr * e5);
        const double q5 = m53 * d3 + m54 * d4 + d5 + h * ((sr * e4 + cr * e5) * icp);
        if (c < 36) {
          fx[3 * 36 + c] = q3;
          fx[4 * 36 + c] = q4;
          fx[5 * 36 + c] = q5;
        } else {
          fu[3 * 12 + c - 36] = q3;
          fu[4 * 12 + c - 36] = q4;
          fu[5 * 12 + c - 36] = q5;
        }
      }
    }
    __syncwarp();
#endif
    }
    ok = nok;
    b = nb;
    t = nt;
    cur = nxt;
    item += stride;
  }
}

}  // namespace ddp
