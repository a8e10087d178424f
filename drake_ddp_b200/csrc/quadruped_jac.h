// Structured linearization of the quadruped model (models.h Quadruped::step): instead of
// pushing all n+m = 48 seed directions through both substeps, each substep's Jacobian
//   D_s = d(q+, v+) / d(q, v, u)     (36 x 48)
// is assembled from the four legs' closed-form local Jacobians (quadruped_legjac.h, 9 x 16
// each) plus the analytic base/integrator terms, and the substeps are chained:
//   J <- D_s[:, :36] J + [0 | D_s[:, 36:]].
// The functions are written for `nl` cooperating lanes (lane, nl) so the same code runs
// serially on the host (lane 0 of 1: used by the CPU tests to check it against forward-mode
// AD) and warp-parallel in the CUDA kernel.
#pragma once
#include "models.h"
#include "quadruped_legjac.h"

namespace ddp {

struct QuadJac {
  // local input j of leg l -> global column of D (0..47)
  DDP_HD static int gcol(int l, int j) {
    if (j == 0) return 2;                 // pz
    if (j < 4) return 2 + j;              // roll pitch yaw -> 3 4 5
    if (j < 10) return 18 + (j - 4);      // v0..v5
    if (j < 13) return 6 + 3 * l + (j - 10);
    return 24 + 3 * l + (j - 13);
  }

  // local inputs of leg l and its contact state (Fn, dFn) at (q, v)
  DDP_HD static void leg_inputs(int l, const double* q, const double* v, const double* p, double* xi,
                                double* Fn, double* dFn) {
    xi[0] = q[2]; xi[1] = q[3]; xi[2] = q[4]; xi[3] = q[5];
    for (int i = 0; i < 6; ++i) xi[4 + i] = v[i];
    for (int k = 0; k < 3; ++k) {
      xi[10 + k] = q[6 + 3 * l + k];
      xi[13 + k] = v[6 + 3 * l + k];
    }
    // foot height -> depth -> phi, phi'
    const double sx = (l < 2) ? 1.0 : -1.0, sd = (l & 1) ? 1.0 : -1.0;
    const double l1 = p[10], l2 = p[11], l3 = p[12], hx = p[13], hy = p[14], rf = p[15], E = p[16];
    double sr, cr, sp, cp, sa, ca, sh, ch, sk, ck;
    sincos_(q[3], &sr, &cr);
    sincos_(q[4], &sp, &cp);
    sincos_(xi[10], &sa, &ca);
    sincos_(xi[11], &sh, &ch);
    sincos_(xi[11] + xi[12], &sk, &ck);
    const double ly = sd * l1, lx = -(l2 * sh) - l3 * sk, lz = -(l2 * ch) - l3 * ck;
    const double rx = hx * sx + lx, ry = hy * sd + (ly * ca - lz * sa), rz = ly * sa + lz * ca;
    const double cz = q[2] + (-sp) * rx + (cp * sr) * ry + (cp * cr) * rz;
    const double depth = rf - cz;
    const double piE = 3.14159265358979323846 * E;
    if (depth <= 0.0) {
      *Fn = 0.0;
      *dFn = 0.0;
    } else if (depth >= rf) {
      *Fn = piE * rf * rf / 3.0;
      *dFn = 0.0;
    } else {
      *Fn = piE * depth * depth * (1.0 - depth * (2.0 / (3.0 * rf)));
      *dFn = piE * (2.0 * depth - 2.0 * depth * depth / rf);
    }
  }

  // D (36 x 48, row-major, leading dimension ld) of one substep of length h at (q, v) from the
  // four legs' local Jacobians G[l] (9 x 16 each).  vplus = v after the substep (needed by the
  // Euler-angle kinematics).  Lanes split the rows.
  DDP_HD static void assemble(const double* G, const double* q, const double* v, const double* vplus,
                              const double* p, double h, double* D, int ld, int lane, int nl) {
    const double mass = p[2], Ix = p[3], Iy = p[4], Iz = p[5];
    // ---- velocity rows 18..35: Dv = E_v + h * dacc/dz --------------------------------------
    for (int r = lane; r < 18; r += nl) {
      double* row = D + (size_t)(18 + r) * ld;
      for (int c = 0; c < 48; ++c) row[c] = 0.0;
      if (r < 6) {
        const double inv = (r < 3) ? 1.0 / mass : (r == 3 ? 1.0 / Ix : (r == 4 ? 1.0 / Iy : 1.0 / Iz));
        for (int l = 0; l < 4; ++l) {
          const double* g = G + (size_t)l * 144 + (size_t)r * 16;
          for (int j = 0; j < 16; ++j) row[gcol(l, j)] += (h * inv) * g[j];
        }
        if (r == 3) {
          row[22] += -h * (Iz - Iy) * v[5] / Ix;
          row[23] += -h * (Iz - Iy) * v[4] / Ix;
        } else if (r == 4) {
          row[23] += -h * (Ix - Iz) * v[3] / Iy;
          row[21] += -h * (Ix - Iz) * v[5] / Iy;
        } else if (r == 5) {
          row[21] += -h * (Iy - Ix) * v[4] / Iz;
          row[22] += -h * (Iy - Ix) * v[3] / Iz;
        }
      } else {
        const int l = (r - 6) / 3, k = (r - 6) % 3;
        const double* g = G + (size_t)l * 144 + (size_t)(6 + k) * 16;
        for (int j = 0; j < 16; ++j) row[gcol(l, j)] += h * g[j];
        row[36 + 3 * l + k] += h / p[6 + k];
      }
      row[18 + r] += 1.0;
    }
  }
  // position rows 0..17 from the finished velocity rows (call after a sync over the lanes)
  DDP_HD static void assemble_q(const double* q, const double* vplus, double h, double* D, int ld, int lane,
                                int nl) {
    double sr, cr, sp, cp;
    sincos_(q[3], &sr, &cr);
    sincos_(q[4], &sp, &cp);
    const double tp = sp / cp;
    const double wy = vplus[4], wz = vplus[5];
    const double wyz = sr * wy + cr * wz, wr = cr * wy - sr * wz;
    for (int r = lane; r < 18; r += nl) {
      double* row = D + (size_t)r * ld;
      const double* dv = D + (size_t)(18 + r) * ld;
      if (r < 3 || r >= 6) {
        for (int c = 0; c < 48; ++c) row[c] = h * dv[c];
        row[r] += 1.0;
      } else {
        const double* d3 = D + (size_t)21 * ld;
        const double* d4 = D + (size_t)22 * ld;
        const double* d5 = D + (size_t)23 * ld;
        if (r == 3) {
          for (int c = 0; c < 48; ++c) row[c] = h * (d3[c] + tp * (sr * d4[c] + cr * d5[c]));
          row[3] += 1.0 + h * tp * wr;
          row[4] += h * wyz / (cp * cp);
        } else if (r == 4) {
          for (int c = 0; c < 48; ++c) row[c] = h * (cr * d4[c] - sr * d5[c]);
          row[3] += -h * wyz;
          row[4] += 1.0;
        } else {
          for (int c = 0; c < 48; ++c) row[c] = h * ((sr * d4[c] + cr * d5[c]) / cp);
          row[3] += h * wr / cp;
          row[4] += h * wyz * sp / (cp * cp);
          row[5] += 1.0;
        }
      }
    }
  }

  // ---- element-wise (gather) forms of the same assembly, for the CUDA kernel ---------------
  // velocity row r (0..17), column c (0..47) of D:  d v+_r / d z_c
  DDP_HD static double dv_elem(const double* G, const double* v, const double* p, double h, int r, int c) {
    double val = (c == 18 + r) ? 1.0 : 0.0;
    int l = -1, j = -1;
    if (c >= 2 && c <= 5) j = c - 2;
    else if (c >= 18 && c < 24) j = c - 14;
    else if (c >= 6 && c < 18) { l = (c - 6) / 3; j = 10 + (c - 6) % 3; }
    else if (c >= 24 && c < 36) { l = (c - 24) / 3; j = 13 + (c - 24) % 3; }
    if (r < 6) {
      const double mass = p[2], Ix = p[3], Iy = p[4], Iz = p[5];
      const double inv = (r < 3) ? 1.0 / mass : (r == 3 ? 1.0 / Ix : (r == 4 ? 1.0 / Iy : 1.0 / Iz));
      if (j >= 0) {
        const double* g = G + (size_t)r * 16 + j;
        const double gs = (l >= 0) ? g[(size_t)l * 144] : ((g[0] + g[144]) + (g[288] + g[432]));
        val += (h * inv) * gs;
      }
      if (r == 3) {
        if (c == 22) val += -h * (Iz - Iy) * v[5] / Ix;
        if (c == 23) val += -h * (Iz - Iy) * v[4] / Ix;
      } else if (r == 4) {
        if (c == 23) val += -h * (Ix - Iz) * v[3] / Iy;
        if (c == 21) val += -h * (Ix - Iz) * v[5] / Iy;
      } else if (r == 5) {
        if (c == 21) val += -h * (Iy - Ix) * v[4] / Iz;
        if (c == 22) val += -h * (Iy - Ix) * v[3] / Iz;
      }
    } else {
      const int lr = (r - 6) / 3, k = (r - 6) % 3;
      if (j >= 0 && (l < 0 || l == lr)) val += h * G[(size_t)lr * 144 + (size_t)(6 + k) * 16 + j];
      if (c == 36 + 3 * lr + k) val += h / p[6 + k];
    }
    return val;
  }
  // position row r (0..17), column c of D from the finished velocity rows Dv (18 x 48, ld)
  DDP_HD static double dq_elem(const double* Dv, int ld, double sr, double cr, double sp, double cp,
                               const double* vplus, double h, int r, int c) {
    if (r < 3 || r >= 6) return ((c == r) ? 1.0 : 0.0) + h * Dv[(size_t)r * ld + c];
    const double d3 = Dv[(size_t)3 * ld + c], d4 = Dv[(size_t)4 * ld + c], d5 = Dv[(size_t)5 * ld + c];
    const double tp = sp / cp, wy = vplus[4], wz = vplus[5];
    const double wyz = sr * wy + cr * wz, wr = cr * wy - sr * wz;
    if (r == 3) {
      double val = h * (d3 + tp * (sr * d4 + cr * d5));
      if (c == 3) val += 1.0 + h * tp * wr;
      if (c == 4) val += h * wyz / (cp * cp);
      return val;
    }
    if (r == 4) {
      double val = h * (cr * d4 - sr * d5);
      if (c == 3) val += -h * wyz;
      if (c == 4) val += 1.0;
      return val;
    }
    double val = h * ((sr * d4 + cr * d5) / cp);
    if (c == 3) val += h * wr / cp;
    if (c == 4) val += h * wyz * sp / (cp * cp);
    if (c == 5) val += 1.0;
    return val;
  }

  // Serial reference of the whole thing (host tests): fx (36 x 36), fu (36 x 12) row-major.
  static inline void step_jac_serial(const double* x, const double* u, const double* p, double* fx,
                                     double* fu) {
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    double q[18], v[18];
    for (int i = 0; i < 18; ++i) {
      q[i] = x[i];
      v[i] = x[18 + i];
    }
    double J[36 * 48], D[36 * 48], T[36 * 48], G[4 * 144];
    for (int s = 0; s < sub; ++s) {
      for (int l = 0; l < 4; ++l) {
        double xi[16], Fn, dFn;
        leg_inputs(l, q, v, p, xi, &Fn, &dFn);
        quadruped_leg_jac((l < 2) ? 1.0 : -1.0, (l & 1) ? 1.0 : -1.0, xi, p, Fn, dFn, G + l * 144);
      }
      double q0[18], v0[18];
      for (int i = 0; i < 18; ++i) {
        q0[i] = q[i];
        v0[i] = v[i];
      }
      Quadruped::substep<double>(q, v, u, p, h);
      assemble(G, q0, v0, v, p, h, D, 48, 0, 1);
      assemble_q(q0, v, h, D, 48, 0, 1);
      if (s == 0) {
        for (int i = 0; i < 36 * 48; ++i) J[i] = D[i];
      } else {
        for (int r = 0; r < 36; ++r)
          for (int c = 0; c < 48; ++c) {
            double a = (c >= 36) ? D[r * 48 + c] : 0.0;
            for (int k = 0; k < 36; ++k) a += D[r * 48 + k] * J[k * 48 + c];
            T[r * 48 + c] = a;
          }
        for (int i = 0; i < 36 * 48; ++i) J[i] = T[i];
      }
    }
    for (int r = 0; r < 36; ++r) {
      for (int c = 0; c < 36; ++c) fx[r * 36 + c] = J[r * 48 + c];
      for (int c = 0; c < 12; ++c) fu[r * 12 + c] = J[r * 48 + 36 + c];
    }
  }
};

}  // namespace ddp
