// K4 alternative for the quadruped model: structured linearization (quadruped_jac.h).
//
// OPT-IN (DDP_QUAD_STRUCTURED=1).  Measured on B200 at C4 it is correct to 1e-12 against the AD
// kernel but not faster yet (5.1 ms vs 4.5 ms per 1024 x 199 points: the assembly + chain kernel
// executes 1.6 G warp instructions and the generated leg code spills), so the generic
// forward-mode-AD linearize_kernel stays the default.  Two kernels:
//   quad_legjac_kernel  4 lanes per (trajectory, keypoint): lane l evaluates the closed-form
//                       9 x 16 local Jacobian of leg l at every substep (generated code,
//                       quadruped_legjac.h) and the group advances the state between substeps
//                       with the same lane-cooperative substep the rollout uses;
//   quad_chain_kernel   one warp per (trajectory, keypoint): assembles the substep Jacobians
//                       D_s (36 x 48) in shared memory and chains them on the fp64 tensor pipe
//                       (mma.sync.m8n8k4.f64): [fx | fu] = D_2[:, :36] D_1 + [0 | D_2[:, 36:]].
// Both produce exactly the derivative of Quadruped::step (checked against the AD kernel and the
// host AD to 1e-12 in the tests).  Supports 1 or 2 substeps; other settings use the AD kernel.
#pragma once
#include "backward_mma.cuh"
#include "quadruped_jac.h"

namespace ddp {

// One cooperative substep for a 4-lane group (same arithmetic as Quadruped::step_coop's body).
__device__ __forceinline__ void quad_substep_coop(int lane, unsigned mask, int gbase, double* q, double* v,
                                                  const double* u, const double* p, double h) {
  typedef Quadruped Qd;
  const double ua = Qd::pick4(lane, u[0], u[3], u[6], u[9]);
  const double uh = Qd::pick4(lane, u[1], u[4], u[7], u[10]);
  const double uk = Qd::pick4(lane, u[2], u[5], u[8], u[11]);
  const double sx = (lane < 2) ? 1.0 : -1.0, sd = (lane & 1) ? 1.0 : -1.0;
  Qd::BasePose<double> B;
  Qd::base_pose(q, B);
  Qd::LegOut<double> o;
  Qd::leg(sx, sd, Qd::pick4(lane, q[6], q[9], q[12], q[15]), Qd::pick4(lane, q[7], q[10], q[13], q[16]),
          Qd::pick4(lane, q[8], q[11], q[14], q[17]), Qd::pick4(lane, v[6], v[9], v[12], v[15]),
          Qd::pick4(lane, v[7], v[10], v[13], v[16]), Qd::pick4(lane, v[8], v[11], v[14], v[17]), ua, uh, uk,
          q[2], v, B, p, o);
  double f[6] = {o.Fx, o.Fy, o.Fz, o.Tx, o.Ty, o.Tz};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    f[k] += __shfl_xor_sync(mask, f[k], 1);
    f[k] += __shfl_xor_sync(mask, f[k], 2);
  }
  double acc[18];
  Qd::base_acc(f[0], f[1], f[2], f[3], f[4], f[5], v, p, acc);
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    acc[6 + 3 * l] = __shfl_sync(mask, o.a0, gbase + l);
    acc[7 + 3 * l] = __shfl_sync(mask, o.a1, gbase + l);
    acc[8 + 3 * l] = __shfl_sync(mask, o.a2, gbase + l);
  }
  Qd::integrate(q, v, acc, B, h);
}

// G    [ceil(B*T/8)][sub][144][32]  local leg Jacobians, the 32 lanes of a warp (8 items x 4 legs)
//                                interleaved innermost so stores are fully coalesced
// xmid [B*T][sub][36]            state after each substep
__global__ void __launch_bounds__(128, 3) quad_legjac_kernel(Dev d, const int* list, const int* count,
                                                             double* G, double* xmid, int sub) {
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gtid >> 2;
  const int lane = (int)(gtid & 3);
  if (item >= (size_t)d.B * d.T) return;
  const int b = (int)(item / d.T), i = (int)(item % d.T);
  if (!d.active[b] || i >= count[b]) return;
  const int t = list[(size_t)b * d.T + i];
  const unsigned mask = group_mask(4);
  const int gbase = (threadIdx.x & 31) / 4 * 4;
  const double* p = d.params;
  const double h = p[0] / sub;
  const double* xp = d.x_bar + ((size_t)b * d.N + t) * 36;
  const double* up = d.u_bar + ((size_t)b * d.T + t) * 12;
  double q[18], v[18], u[12];
#pragma unroll
  for (int k = 0; k < 18; ++k) {
    q[k] = xp[k];
    v[k] = xp[18 + k];
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) u[k] = up[k];
  const size_t bt = (size_t)b * d.T + t;
  for (int s = 0; s < sub; ++s) {
    double xi[16], Fn, dFn;
    // local inputs of this lane's leg (static indexing via pick4)
    xi[0] = q[2]; xi[1] = q[3]; xi[2] = q[4]; xi[3] = q[5];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi[4 + k] = v[k];
    xi[10] = Quadruped::pick4(lane, q[6], q[9], q[12], q[15]);
    xi[11] = Quadruped::pick4(lane, q[7], q[10], q[13], q[16]);
    xi[12] = Quadruped::pick4(lane, q[8], q[11], q[14], q[17]);
    xi[13] = Quadruped::pick4(lane, v[6], v[9], v[12], v[15]);
    xi[14] = Quadruped::pick4(lane, v[7], v[10], v[13], v[16]);
    xi[15] = Quadruped::pick4(lane, v[8], v[11], v[14], v[17]);
    {
      // contact state of this lane's leg (same expressions as QuadJac::leg_inputs)
      const double sx = (lane < 2) ? 1.0 : -1.0, sd = (lane & 1) ? 1.0 : -1.0;
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
        Fn = 0.0;
        dFn = 0.0;
      } else if (depth >= rf) {
        Fn = piE * rf * rf / 3.0;
        dFn = 0.0;
      } else {
        Fn = piE * depth * depth * (1.0 - depth * (2.0 / (3.0 * rf)));
        dFn = piE * (2.0 * depth - 2.0 * depth * depth / rf);
      }
      quadruped_leg_jac(sx, sd, xi, p, Fn, dFn,
                        G + (((gtid >> 5) * sub + s) * 144) * 32 + (threadIdx.x & 31), 32);
    }
    {
      quad_substep_coop(lane, mask, gbase, q, v, u, p, h);
      {
        double* xm = xmid + (bt * sub + s) * 36;
#pragma unroll
        for (int k = 0; k < 18; ++k) {
          if ((k & 3) == lane) {
            xm[k] = q[k];
            xm[18 + k] = v[k];
          }
        }
      }
    }
  }
}

constexpr int kQuadLd = 52;            // leading dimension of D in shared memory (conflict-free DMMA loads)
constexpr int kQuadThreads = 192;      // one CTA (6 warps = 6 column strips) per (trajectory, keypoint)
constexpr int kQuadTabSize = 18 * 48;  // one gather descriptor per velocity-row element of D

// Gather descriptor of D[18 + r][c] (velocity row r, column c): up to four offsets into the
// substep's 4 x 144 block of local leg Jacobians (10 bits each, 0x3FF = none), the scale
// selector (h/mass, h/Ix, h/Iy, h/Iz, h) and the identity flag.  Built once on the host from
// the same column map as QuadJac::dv_elem.
inline void quad_build_table(unsigned long long* tab) {
  for (int r = 0; r < 18; ++r)
    for (int c = 0; c < 48; ++c) {
      int l = -1, j = -1;
      if (c >= 2 && c <= 5) j = c - 2;
      else if (c >= 18 && c < 24) j = c - 14;
      else if (c >= 6 && c < 18) { l = (c - 6) / 3; j = 10 + (c - 6) % 3; }
      else if (c >= 24 && c < 36) { l = (c - 24) / 3; j = 13 + (c - 24) % 3; }
      unsigned o[4] = {0x3FF, 0x3FF, 0x3FF, 0x3FF};
      unsigned sel;
      if (r < 6) {
        sel = (r < 3) ? 0 : (unsigned)(r - 2);
        if (j >= 0) {
          if (l >= 0) o[0] = l * 144 + r * 16 + j;
          else for (int q = 0; q < 4; ++q) o[q] = q * 144 + r * 16 + j;
        }
      } else {
        sel = 4;
        const int lr = (r - 6) / 3, k = (r - 6) % 3;
        if (j >= 0 && (l < 0 || l == lr)) o[0] = lr * 144 + (6 + k) * 16 + j;
      }
      unsigned long long e = 0;
      for (int q = 0; q < 4; ++q) e |= (unsigned long long)o[q] << (10 * q);
      e |= (unsigned long long)sel << 40;
      e |= (unsigned long long)(c == 18 + r ? 1 : 0) << 43;
      tab[r * 48 + c] = e;
    }
}

struct QuadChainSmem {
  double D[2][36 * kQuadLd];           // substep Jacobians, rows 0..17 positions, 18..35 velocities
  double Gs[2][4 * 144];               // local leg Jacobians of both substeps
  double st[3][36];                    // x_t and the state after each substep
  double trig[2][4];                   // sin/cos of roll, pitch at the start of each substep
};

__global__ void __launch_bounds__(kQuadThreads) quad_chain_kernel(Dev d, const int* list, const int* count,
                                                                   const double* __restrict__ G,
                                                                   const double* __restrict__ xmid,
                                                                   const unsigned long long* __restrict__ tab,
                                                                   int sub) {
  __shared__ QuadChainSmem s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const size_t item = blockIdx.x;
  const int b = (int)(item / d.T), i = (int)(item % d.T);
  if (!d.active[b] || i >= count[b]) return;
  const int t = list[(size_t)b * d.T + i];
  const size_t bt = (size_t)b * d.T + t;
  const double* p = d.params;
  const double h = p[0] / sub;
  // stage states and the leg Jacobians (coalesced, independent loads)
  if (tid < 36) {
    s.st[0][tid] = d.x_bar[((size_t)b * d.N + t) * 36 + tid];
    for (int k = 0; k < sub; ++k) s.st[1 + k][tid] = xmid[(bt * sub + k) * 36 + tid];
  }
  {
    // item's four legs are four consecutive lanes of the legjac kernel's warp (item / 8)
    const size_t witem = item >> 3;
    const int lbase = (int)(item & 7) * 4;
    for (int k = tid; k < sub * 576; k += kQuadThreads) {
      const int sidx = k / 576, r = k - sidx * 576, l = r / 144, e = r - l * 144;
      s.Gs[sidx][r] = G[((witem * sub + sidx) * 144 + e) * 32 + lbase + l];
    }
  }
  __syncthreads();
  if (tid < sub) {
    const double* x0 = s.st[tid];
    double sr, cr, sp, cp;
    sincos_(x0[3], &sr, &cr);
    sincos_(x0[4], &sp, &cp);
    s.trig[tid][0] = sr; s.trig[tid][1] = cr; s.trig[tid][2] = sp; s.trig[tid][3] = cp;
  }
  // ---- velocity rows: table-driven gather, one element per thread and pass ------------------
  {
    const double coef[5] = {h / p[2], h / p[3], h / p[4], h / p[5], h};
    for (int e = tid; e < kQuadTabSize; e += kQuadThreads) {
      const unsigned long long ent = tab[e];
      const int r = e / 48, c = e - 48 * r;
      const unsigned o0 = (unsigned)(ent & 0x3FF), o1 = (unsigned)((ent >> 10) & 0x3FF);
      const unsigned o2 = (unsigned)((ent >> 20) & 0x3FF), o3 = (unsigned)((ent >> 30) & 0x3FF);
      const double cf = coef[(ent >> 40) & 7];
      const double id = ((ent >> 43) & 1) ? 1.0 : 0.0;
      for (int sidx = 0; sidx < sub; ++sidx) {
        const double* Gs = s.Gs[sidx];
        double gs = 0.0;
        if (o1 != 0x3FF) gs = (Gs[o0] + Gs[o1]) + (Gs[o2] + Gs[o3]);
        else if (o0 != 0x3FF) gs = Gs[o0];
        s.D[sidx][(18 + r) * kQuadLd + c] = id + cf * gs;
      }
    }
  }
  __syncthreads();
  // gyroscopic terms and the direct u -> joint acceleration terms (18 entries per substep)
  if (tid < 18 * sub) {
    const int sidx = tid / 18, k = tid % 18;
    const double* v = s.st[sidx] + 18;
    double* Dv = &s.D[sidx][18 * kQuadLd];
    const double Ix = p[3], Iy = p[4], Iz = p[5];
    if (k < 6) {
      const int rr[6] = {3, 3, 4, 4, 5, 5}, cc[6] = {22, 23, 23, 21, 21, 22};
      const double val[6] = {-h * (Iz - Iy) * v[5] / Ix, -h * (Iz - Iy) * v[4] / Ix, -h * (Ix - Iz) * v[3] / Iy,
                             -h * (Ix - Iz) * v[5] / Iy, -h * (Iy - Ix) * v[4] / Iz, -h * (Iy - Ix) * v[3] / Iz};
      Dv[rr[k] * kQuadLd + cc[k]] += val[k];
    } else {
      const int jj = k - 6;   // joint 0..11
      Dv[(6 + jj) * kQuadLd + 36 + jj] += h / p[6 + jj % 3];
    }
  }
  __syncthreads();
  // ---- position rows from the velocity rows ---------------------------------------------------
  for (int e = tid; e < sub * 18 * 48; e += kQuadThreads) {
    const int sidx = e / (18 * 48), e2 = e - sidx * (18 * 48), r = e2 / 48, c = e2 - 48 * r;
    const double* x1 = s.st[sidx + 1];
    s.D[sidx][r * kQuadLd + c] =
        QuadJac::dq_elem(&s.D[sidx][18 * kQuadLd], kQuadLd, s.trig[sidx][0], s.trig[sidx][1], s.trig[sidx][2],
                         s.trig[sidx][3], x1 + 18, h, r, c);
  }
  __syncthreads();
  double* fx = d.fx + bt * 36 * 36;
  double* fu = d.fu + bt * 36 * 12;
  if (sub == 1) {
    for (int idx = tid; idx < 36 * 48; idx += kQuadThreads) {
      const int r = idx / 48, c = idx % 48;
      if (c < 36) fx[r * 36 + c] = s.D[0][r * kQuadLd + c];
      else fu[r * 12 + (c - 36)] = s.D[0][r * kQuadLd + c];
    }
    return;
  }
  // ---- [fx | fu] = D2[:, :36] D1 + [0 | D2[:, 36:]]: warp w owns the 8-column strip w ---------
  const double* D1 = s.D[0];
  const double* D2 = s.D[1];
  double acc[5][2];
#pragma unroll
  for (int mt = 0; mt < 5; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) {
    const int k = 4 * kk + tg;
    const double bf = D1[k * kQuadLd + 8 * warp + g];
#pragma unroll
    for (int mt = 0; mt < 5; ++mt) {
      const int r = 8 * mt + g;
      dmma(acc[mt], (r < 36) ? D2[r * kQuadLd + k] : 0.0, bf);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 5; ++mt) {
    const int r = 8 * mt + g, c = 8 * warp + 2 * tg;
    if (r < 36) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = c + e;
        if (cc < 36) fx[r * 36 + cc] = acc[mt][e];
        else fu[r * 12 + (cc - 36)] = acc[mt][e] + D2[r * kQuadLd + cc];
      }
    }
  }
}

}  // namespace ddp
