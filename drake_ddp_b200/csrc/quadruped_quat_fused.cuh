// K4 for the quadruped in the reference's own state layout (QuadrupedQuat, n = 37,
// mini_cheetah.py:41-57): structured linearization fused into one kernel.
//
// Replaces _calc_dynamics_partials (/root/reference/ilqr.py:233-272) over the keypoints (:409-411)
// for QuadrupedQuat::step (models.h).  Same idea as quad_fused_kernel (quadruped_fused.cuh): a
// leg's loads depend on 16 local inputs only, so per substep the 4 x 16 (leg, local direction)
// pairs are dealt to the 32 lanes of a warp and Quadruped::leg_trig is evaluated with Dual<2>.
// With a quaternion base the local inputs are INTERMEDIATE quantities, not state entries:
//     zeta = [ base height | theta (3): body-frame rotation of R | v_lin (3) | omega_body (3) |
//              the leg's joints (3) | its joint rates (3) ]
// so the dual evaluation yields Z = d(body-frame accelerations) / d zeta (18 x 48, the column
// layout of the Euler-angle kernel) and the substep Jacobian in state coordinates follows by
//   * columns:  theta <- G(q) dq  with  G = 2/|q| [ -b | a I - [b]x ]  (q/|q| = (a, b); the
//     normalisation of the quaternion is inside G: G q = 0),  omega_body = R' omega_world
//     (d omega_body = R' d omega_world + [omega_body]x d theta);
//   * rows:  omega_world+ = omega_world + h R ab  (ab: body-frame angular acceleration), i.e.
//     I + h R d(ab) - h R [ab]x d theta;  q+ = q + h/2 Omega(omega+) q;  pos / joints: identity + h v+.
// The two substeps are chained on the fp64 tensor pipe: velocity rows only,
// V2 = J2v[:, x] J1 + [0 | J2v[:, u]] (18 x 37 x 49, 210 DMMAs), position rows follow row-wise.
// One warp per point.  Exact derivative of QuadrupedQuat::step: checked against the generic AD
// kernel (tests/test_gpu_parity.py).  Two substeps only (the model's setting).
#pragma once
#include "quadruped_fused.cuh"

namespace ddp {

// entry i (0..3 / 0..2) of a few values that live in registers, as a tree of selects on the bits
// of i: a nested ternary over more than two values compiles to a branch ladder (and these sit on
// the serial path of a point)
__device__ __forceinline__ double qq_sel4(int i, double a, double b, double c, double d) {
  const bool b0 = (i & 1) != 0, b1 = (i & 2) != 0;
  const double lo = b0 ? b : a, hi = b0 ? d : c;
  return b1 ? hi : lo;
}
__device__ __forceinline__ double qq_sel3(int i, double a, double b, double c) { return qq_sel4(i, a, b, c, c); }


struct QqWarpSmem {
  double J1[40 * kQfLd];    // Jacobian of substep 1, rows = entries of x1 (37; rows 37..39 stay zero), cols = [x0 (37) | u (12)]
  double Z[20 * kQfLd];     // d(body-frame v+) / d zeta of the current substep; substep 2: transformed in place to J2's velocity rows
  double w2s[3 * kQfLd];    // rows omega+ of the chained result (the quaternion rows need all three)
  double st[3][38];         // x_t, state after substep 1, after substep 2
};

#ifndef QQ_MINB
#define QQ_MINB 2
#endif

__global__ void __launch_bounds__(kQfWarps * 32, QQ_MINB)
quad_quat_fused_kernel(Dev d, const int* list, const int* count, int n_items) {
  typedef Quadruped Qd;
  typedef Dual<2> D2;
  constexpr int LD = kQfLd, n = 37, NX = 37, NC = 49;
  extern __shared__ __align__(16) unsigned char qq_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  QqWarpSmem& s = reinterpret_cast<QqWarpSmem*>(qq_raw)[warp];
  const unsigned full = 0xffffffffu;
  const double* p = d.pm;   // constant bank (Dev::pm)
  const double h = p[0] / 2.0;
  const double Ix = p[3], Iy = p[4], Iz = p[5], grav = p[19];
  const int T = d.T;

  const int leg = lane >> 3, dp = lane & 7;
  const bool shared_dir = dp < 5;   // base directions: every leg contributes to the base rows
  const double sx = (leg < 2) ? 1.0 : -1.0, sd = (leg & 1) ? 1.0 : -1.0;
  const int c0 = lane, c1 = lane + 32;        // the two state / control columns this lane assembles
  const bool has1 = c1 < NC;

  for (int i = lane; i < 40 * LD; i += 32) s.J1[i] = 0.0;
  for (int i = lane; i < 20 * LD; i += 32) s.Z[i] = 0.0;
  __syncwarp();

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
  const int ang_a = (dp == 0) ? 7 + 3 * leg : 8 + 3 * leg;
  const int ang_b = 9 + 3 * leg;
  const double ang_wb = (dp >= 2) ? 1.0 : 0.0;
  // columns in Z of this lane's two local directions (per-lane constants: qf_gcol branches)
  const int gc_lane[2] = {qf_gcol(leg, 2 * dp), qf_gcol(leg, 2 * dp + 1)};
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
    const double* xp = d.x_bar + ((size_t)bb * d.N + tt) * n;
    const double* up = d.u_bar + ((size_t)bb * T + tt) * 12;
    r.x0 = xp[lane];
    r.x1 = (lane < 5) ? xp[32 + lane] : 0.0;
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
    if (lane < 5) s.st[0][32 + lane] = cur.x1;
    const double ua = cur.ua, uh = cur.uh, uk = cur.uk;
    // the direct u -> joint acceleration terms (constant) of the intermediate Jacobian
    if (lane < 12) s.Z[(6 + lane) * LD + 36 + lane] = h * p[24 + lane % 3];
    __syncwarp();

#pragma unroll 1
    for (int sub = 0; sub < 2; ++sub) {
      const double* xin = s.st[sub];
      double* xout = s.st[sub + 1];
      double* Z = s.Z;
      // ---- base: rotation of the normalised quaternion, body-frame twist -----------------------
      const double q0 = xin[0], q1 = xin[1], q2 = xin[2], q3 = xin[3];
      const double inn = inv_sqrt_(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
      const double qa = q0 * inn, qb = q1 * inn, qc = q2 * inn, qd = q3 * inn;
      double R[3][3];
      R[0][0] = 1.0 - 2.0 * (qc * qc + qd * qd); R[0][1] = 2.0 * (qb * qc - qa * qd); R[0][2] = 2.0 * (qb * qd + qa * qc);
      R[1][0] = 2.0 * (qb * qc + qa * qd); R[1][1] = 1.0 - 2.0 * (qb * qb + qd * qd); R[1][2] = 2.0 * (qc * qd - qa * qb);
      R[2][0] = 2.0 * (qb * qd - qa * qc); R[2][1] = 2.0 * (qc * qd + qa * qb); R[2][2] = 1.0 - 2.0 * (qb * qb + qc * qc);
      const double w0 = xin[19], w1 = xin[20], w2 = xin[21];
      double vloc[6];
      vloc[0] = xin[22]; vloc[1] = xin[23]; vloc[2] = xin[24];
      vloc[3] = R[0][0] * w0 + R[1][0] * w1 + R[2][0] * w2;
      vloc[4] = R[0][1] * w0 + R[1][1] * w1 + R[2][1] * w2;
      vloc[5] = R[0][2] * w0 + R[1][2] * w1 + R[2][2] * w2;
      // ---- dual evaluation of this lane's leg along its two local directions ---------------------
      double sn, cs;
      {
        const double ang = xin[ang_a] + ang_wb * xin[ang_b];
        sincos_(ang, &sn, &cs);
      }
      const int gl = lane & ~7;   // first lane of this leg's group
      double f[6], av[3];
      {
        int jd[2], gc[2];      // local directions of this lane and their columns in Z
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          jd[e] = 2 * dp + e;
          gc[e] = gc_lane[e];
        }
        auto seed = [&](double v, int j) {
          D2 r;
          r.v = v;
#pragma unroll
          for (int e = 0; e < 2; ++e) r.d[e] = (j == jd[e]) ? 1.0 : 0.0;
          return r;
        };
        auto trig_dual = [&](int src, int ja, int jb, D2& sD, D2& cD) {
          const double sv = __shfl_sync(full, sn, src), cv = __shfl_sync(full, cs, src);
          sD.v = sv;
          cD.v = cv;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double we = ((jd[e] == ja) || (jd[e] == jb)) ? 1.0 : 0.0;
            sD.d[e] = cv * we;
            cD.d[e] = -sv * we;
          }
        };
        D2 vb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) vb[k] = seed(vloc[k], 4 + k);
        const D2 pz = seed(xin[6], 0);
        // R (I + [theta]x): d R / d theta_k = R [e_k]x, column pattern per k
        Qd::BasePose<D2> B;
        auto rot_dual = [&](int i, D2& r0, D2& r1, D2& r2) {
          r0.v = R[i][0]; r1.v = R[i][1]; r2.v = R[i][2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = jd[e];
            r0.d[e] = (k == 2) ? -R[i][2] : ((k == 3) ? R[i][1] : 0.0);
            r1.d[e] = (k == 1) ? R[i][2] : ((k == 3) ? -R[i][0] : 0.0);
            r2.d[e] = (k == 1) ? -R[i][1] : ((k == 2) ? R[i][0] : 0.0);
          }
        };
        rot_dual(0, B.R00, B.R01, B.R02);
        rot_dual(1, B.R10, B.R11, B.R12);
        rot_dual(2, B.R20, B.R21, B.R22);
        B.sr = D2(0.0); B.cr = D2(1.0); B.sp = D2(0.0); B.cp = D2(1.0);
        D2 sa, ca, sh, ch, sk, ck;
        trig_dual(gl + 0, 10, -1, sa, ca);
        trig_dual(gl + 1, 11, -1, sh, ch);
        trig_dual(gl + 2, 11, 12, sk, ck);
        Qd::LegOut<D2> o;
        Qd::leg_trig(sx, sd, sa, ca, sh, ch, sk, ck, seed(xin[25 + 3 * leg], 13), seed(xin[26 + 3 * leg], 14),
                     seed(xin[27 + 3 * leg], 15), D2(ua), D2(uh), D2(uk), pz, vb, B, p, o);
        // ---- scatter into Z: joint rows of this leg, base rows summed over the legs -------------
        {
          const D2* ja[3] = {&o.a0, &o.a1, &o.a2};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = 6 + 3 * leg + k;
#pragma unroll
            for (int e = 0; e < 2; ++e) Z[r * LD + gc[e]] = h * ja[k]->d[e] + ((gc[e] == 18 + r) ? 1.0 : 0.0);
          }
          const D2* fo[6] = {&o.Fx, &o.Fy, &o.Fz, &o.Tx, &o.Ty, &o.Tz};
          const double inv[6] = {p[20], p[20], p[20], p[21], p[22], p[23]};
#pragma unroll
          for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double own = fo[r]->d[e], sum = own;
              sum += __shfl_xor_sync(full, sum, 8);
              sum += __shfl_xor_sync(full, sum, 16);
              const double gs = shared_dir ? sum : own;
              // rows 0..2: v_lin+ (identity on its own column); rows 3..5: h d(ab), no identity
              if (!shared_dir || leg == 0)
                Z[r * LD + gc[e]] = (h * inv[r]) * gs + ((r < 3 && gc[e] == 18 + r) ? 1.0 : 0.0);
            }
          }
        }
        f[0] = o.Fx.v; f[1] = o.Fy.v; f[2] = o.Fz.v; f[3] = o.Tx.v; f[4] = o.Ty.v; f[5] = o.Tz.v;
        av[0] = o.a0.v; av[1] = o.a1.v; av[2] = o.a2.v;
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(full, f[k], 8);
        f[k] += __shfl_xor_sync(full, f[k], 16);
      }
      // body-frame angular acceleration (Euler's equations) and the primal state after the substep
      const double ab0 = (f[3] - (Iz - Iy) * vloc[4] * vloc[5]) * p[21];
      const double ab1 = (f[4] - (Ix - Iz) * vloc[5] * vloc[3]) * p[22];
      const double ab2 = (f[5] - (Iy - Ix) * vloc[3] * vloc[4]) * p[23];
      const double wn0 = w0 + h * (R[0][0] * ab0 + R[0][1] * ab1 + R[0][2] * ab2);
      const double wn1 = w1 + h * (R[1][0] * ab0 + R[1][1] * ab1 + R[1][2] * ab2);
      const double wn2 = w2 + h * (R[2][0] * ab0 + R[2][1] * ab1 + R[2][2] * ab2);
      {
        const int jl = (lane >= 6 && lane < 18) ? (lane - 6) / 3 : 0, jk = (lane >= 6 && lane < 18) ? (lane - 6) % 3 : 0;
        const double a0 = __shfl_sync(full, av[0], 8 * jl), a1 = __shfl_sync(full, av[1], 8 * jl),
                     a2 = __shfl_sync(full, av[2], 8 * jl);
        const double aj = qq_sel3(jk, a0, a1, a2);
        double vnew;   // velocity entry `lane` of [omega (3) | v_lin (3) | joint rates (12)]
        if (lane < 3) vnew = qq_sel3(lane, wn0, wn1, wn2);
        else if (lane < 6) {
          double a = qq_sel3(lane - 3, f[0], f[1], f[2]) * p[20];
          if (lane == 5) a -= grav;
          vnew = xin[19 + lane] + h * a;
        } else vnew = xin[19 + (lane < 18 ? lane : 0)] + h * aj;
        const double e0 = q0 + (0.5 * h) * (-(wn0 * q1) - wn1 * q2 - wn2 * q3);
        const double e1 = q1 + (0.5 * h) * (wn0 * q0 + wn1 * q3 - wn2 * q2);
        const double e2 = q2 + (0.5 * h) * (wn1 * q0 + wn2 * q1 - wn0 * q3);
        const double e3 = q3 + (0.5 * h) * (wn2 * q0 + wn0 * q2 - wn1 * q1);
        if (lane < 18) {
          xout[19 + lane] = vnew;
          if (lane >= 3) xout[lane + 1] = xin[lane + 1] + h * vnew;   // position / joint entry of velocity entry `lane`
        }
        if (lane < 4) xout[lane] = qq_sel4(lane, e0, e1, e2, e3);
      }
      __syncwarp();
      // gyroscopic terms of the body angular rows (d/d omega_body of the omega x I omega term): entry
      // (gy_off) += gy_coef * omega_body[gy_w], constants per lane (branch-free)
      {
        const double wsel = qq_sel3(gy_w, vloc[3], vloc[4], vloc[5]);
        if (lane < 6) Z[gy_off] += gy_coef * wsel;
      }
      __syncwarp();
      if (sub == 0) {
        if (nok) nxt = fetch_pt(nb, nt);   // next point's state and controls: consumed a substep later
      }

      // ---- intermediate coordinates -> state coordinates, for this lane's two columns ------------
      // out[c] = alpha zr[src] + sum_k ck[k] zr[3 + k] + sum_j dk[j] zr[21 + j]   (zr: a row of Z)
      const double G[3][4] = {{-2.0 * inn * qb, 2.0 * inn * qa, 2.0 * inn * qd, -2.0 * inn * qc},
                              {-2.0 * inn * qc, -2.0 * inn * qd, 2.0 * inn * qa, 2.0 * inn * qb},
                              {-2.0 * inn * qd, 2.0 * inn * qc, -2.0 * inn * qb, 2.0 * inn * qa}};
      // [omega_body]x
      const double ob0 = vloc[3], ob1 = vloc[4], ob2 = vloc[5];
      const double OX[3][3] = {{0.0, -ob2, ob1}, {ob2, 0.0, -ob0}, {-ob1, ob0, 0.0}};
      // -R [ab]x
      double MR[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        MR[i][0] = -(R[i][1] * ab2 - R[i][2] * ab1);
        MR[i][1] = -(-R[i][0] * ab2 + R[i][2] * ab0);
        MR[i][2] = -(R[i][0] * ab1 - R[i][1] * ab0);
      }
      double alpha[2], ck[2][3], dk[2][3], wx[2][3];   // wx: h (-R [ab]x G)[i][c] + (c == 19 + i)
      int src[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = q ? c1 : c0;
        alpha[q] = 0.0;
        src[q] = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) ck[q][k] = dk[q][k] = wx[q][k] = 0.0;
        if (c < 4) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double gk = qq_sel4(c, G[k][0], G[k][1], G[k][2], G[k][3]);
            ck[q][k] = gk;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) dk[q][j] = OX[j][0] * ck[q][0] + OX[j][1] * ck[q][1] + OX[j][2] * ck[q][2];
#pragma unroll
          for (int i = 0; i < 3; ++i) wx[q][i] = h * (MR[i][0] * ck[q][0] + MR[i][1] * ck[q][1] + MR[i][2] * ck[q][2]);
        } else if (c >= 19 && c < 22) {
          const int i = c - 19;
#pragma unroll
          for (int j = 0; j < 3; ++j) dk[q][j] = qq_sel3(i, R[0][j], R[1][j], R[2][j]);
#pragma unroll
          for (int r = 0; r < 3; ++r) wx[q][r] = (r == i) ? 1.0 : 0.0;
        } else if (c == 6) {
          alpha[q] = 1.0; src[q] = 2;
        } else if (c >= 7 && c < 19) {
          alpha[q] = 1.0; src[q] = c - 1;
        } else if (c >= 22 && c < 25) {
          alpha[q] = 1.0; src[q] = c - 4;
        } else if (c >= 25 && c < NC) {
          alpha[q] = 1.0; src[q] = c - 1;
        }
      }
      auto tz = [&](int row, int q) {   // transformed entry of Z row `row` at this lane's column q
        const double* zr = Z + row * LD;
        double a = alpha[q] * zr[src[q]];
        a = fma(ck[q][0], zr[3], a); a = fma(ck[q][1], zr[4], a); a = fma(ck[q][2], zr[5], a);
        a = fma(dk[q][0], zr[21], a); a = fma(dk[q][1], zr[22], a); a = fma(dk[q][2], zr[23], a);
        return a;
      };
      // velocity rows of the substep Jacobian in the state's order [omega+ (3) | v_lin+ (3) | joint rates+ (12)]
      double vr[18][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double t3 = tz(3, q), t4 = tz(4, q), t5 = tz(5, q);
#pragma unroll
        for (int i = 0; i < 3; ++i) vr[i][q] = wx[q][i] + (R[i][0] * t3 + R[i][1] * t4 + R[i][2] * t5);
#pragma unroll
        for (int i = 0; i < 3; ++i) vr[3 + i][q] = tz(i, q);
#pragma unroll
        for (int j = 0; j < 12; ++j) vr[6 + j][q] = tz(6 + j, q);
      }
      __syncwarp();   // every read of Z done (substep 2 transforms it in place)
      if (sub == 0) {
        // J1: velocity rows 19..36, then the position rows 0..18 from them
        const double hq = 0.5 * h;
        // d q+ = (I + h/2 A(omega+)) dq + h/2 Bq(q) d omega+
        const double A[4][4] = {{0.0, -wn0, -wn1, -wn2}, {wn0, 0.0, -wn2, wn1}, {wn1, wn2, 0.0, -wn0}, {wn2, -wn1, wn0, 0.0}};
        const double Bq[4][3] = {{-q1, -q2, -q3}, {q0, q3, -q2}, {-q3, q0, q1}, {q2, -q1, q0}};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = q ? c1 : c0;
          if (q == 1 && !has1) break;
#pragma unroll
          for (int r = 0; r < 18; ++r) s.J1[(19 + r) * LD + c] = vr[r][q];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            double v = hq * (Bq[a][0] * vr[0][q] + Bq[a][1] * vr[1][q] + Bq[a][2] * vr[2][q]);
            if (c < 4) {
              const double ac = qq_sel4(c, A[a][0], A[a][1], A[a][2], A[a][3]);
              v += ((a == c) ? 1.0 : 0.0) + hq * ac;
            }
            s.J1[a * LD + c] = v;
          }
#pragma unroll
          for (int r = 3; r < 18; ++r) s.J1[(r + 1) * LD + c] = ((c == r + 1) ? 1.0 : 0.0) + h * vr[r][q];
        }
      } else {
        // J2's velocity rows, in place of Z (dense now: Z is re-zeroed after the chain)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = q ? c1 : c0;
          if (q == 1 && !has1) break;
#pragma unroll
          for (int r = 0; r < 18; ++r) Z[r * LD + c] = vr[r][q];
        }
      }
      __syncwarp();
    }

    // ---- chain: V2 = J2v[:, 0..36] J1 + [0 | J2v[:, 37..48]] -------------------------------------
    double acc[3][7][2];
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    {
      const double* pa[3];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) pa[mt] = s.Z + min(8 * mt + g, 17) * LD + tg;
      const double* pb[7];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) pb[nt] = s.J1 + tg * LD + min(8 * nt + g, NC - 1);
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        // k = 4 kk + tg < 40: rows 37..39 of J1 are zero, the matching columns of J2v are finite
        double a[3], bb[7];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) a[mt] = pa[mt][4 * kk];
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) bb[nt] = pb[nt][4 * kk * LD];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int nt = 0; nt < 7; ++nt) dmma(acc[mt][nt], a[mt], bb[nt]);
      }
    }
    double* fx = d.fx + bt * NX * NX;
    double* fu = d.fu + bt * NX * 12;
    auto store1 = [&](int r, int c, double v) {   // entry (r, c) of [fx | fu]
      if (c < NX) fx[r * NX + c] = v;
      else fu[r * 12 + (c - NX)] = v;
    };
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      const int r = 8 * mt + g;
      if (r < 18) {
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * nt + 2 * tg + e;
            if (c < NC) {
              double v = acc[mt][nt][e];
              if (c >= NX) v += s.Z[r * LD + c];
              store1(19 + r, c, v);
              if (r < 3) s.w2s[r * LD + c] = v;                           // omega+ rows: staged for the quaternion rows
              else store1(r + 1, c, s.J1[(r + 1) * LD + c] + h * v);      // pos / joints: x2 = x1 + h v2
            }
          }
        }
      }
    }
    __syncwarp();
    {
      // quaternion rows: q2 = q1 + h/2 Omega(omega2) q1
      const double hq = 0.5 * h;
      const double u0 = s.st[2][19], u1 = s.st[2][20], u2 = s.st[2][21];      // omega after substep 2
      const double r0 = s.st[1][0], r1 = s.st[1][1], r2 = s.st[1][2], r3 = s.st[1][3];   // quaternion after substep 1
      const double A[4][4] = {{0.0, -u0, -u1, -u2}, {u0, 0.0, -u2, u1}, {u1, u2, 0.0, -u0}, {u2, -u1, u0, 0.0}};
      const double Bq[4][3] = {{-r1, -r2, -r3}, {r0, r3, -r2}, {-r3, r0, r1}, {r2, -r1, r0}};
      for (int c = lane; c < NC; c += 32) {
        const double j0 = s.J1[c], j1 = s.J1[LD + c], j2 = s.J1[2 * LD + c], j3 = s.J1[3 * LD + c];
        const double e0 = s.w2s[c], e1 = s.w2s[LD + c], e2 = s.w2s[2 * LD + c];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double ja = (a == 0) ? j0 : (a == 1) ? j1 : (a == 2) ? j2 : j3;
          const double v = ja + hq * (A[a][0] * j0 + A[a][1] * j1 + A[a][2] * j2 + A[a][3] * j3) +
                           hq * (Bq[a][0] * e0 + Bq[a][1] * e1 + Bq[a][2] * e2);
          store1(a, c, v);
        }
      }
    }
    __syncwarp();
    // Z is dense after the in-place transform: back to its zero pattern for the next point
    for (int i = lane; i < 18 * LD; i += 32) s.Z[i] = 0.0;
    __syncwarp();
    }
    ok = nok;
    b = nb;
    t = nt;
    cur = nxt;
    item += stride;
  }
}

}  // namespace ddp
