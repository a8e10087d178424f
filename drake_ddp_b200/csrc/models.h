// Fixed analytic discrete-time models x+ = f(x, u), shared by the CUDA kernels
// (device) and by the host model library the CPU oracle's dynamics shim calls.
//
// They stand in for the Drake MultibodyPlant discrete update the reference
// calls at /root/reference/ilqr.py:223-229 (forward) and :259-268 (AutoDiff).
// Drake is not available offline, so these are this repo's own models: same
// state sizes and physical constants as the reference's example scripts, a
// semi-implicit Euler discrete map (v+ = v + h a(q,v,u); q+ = q + h N(q) v+),
// and closed-form compliant contact (pressure field p = E(1 - r/R) integrated
// over the sphere/plane contact disk).  See DESIGN.md "Models".
//
// Every model is   template <class S> static void step(x, u, xn, p)
// with S = double (rollouts) or S = ddp::Dual<K> (linearization).
#pragma once
#include "dual.h"

namespace ddp {

enum ModelId {
  MODEL_PENDULUM = 0,       // n=2  m=1   pendulum.py
  MODEL_ACROBOT = 1,        // n=4  m=1   acrobot.py
  MODEL_CARTPOLE = 2,       // n=4  m=1   cart_pole.py
  MODEL_CARTPOLE_WALL = 3,  // n=4  m=1   cart_pole_with_wall.py
  MODEL_QUADRUPED = 4,      // n=36 m=12  mini_cheetah-scale (Euler-angle base)
  MODEL_ARM_BALL = 5,       // n=27 m=7   kinova/panda-scale arm pushing a ball
  MODEL_QUADRUPED_QUAT = 6, // n=37 m=12  the quadruped in the reference's quaternion layout
  MODEL_AFFINE_SIN_4_1 = 10,   // x+ = A x + B u + 0.01 sin x (test stub, any A,B)
  MODEL_AFFINE_SIN_6_2 = 11,
  MODEL_AFFINE_SIN_27_7 = 12,
  MODEL_AFFINE_SIN_36_12 = 13,
  MODEL_AFFINE_SIN_37_12 = 14,
};

// ------------------------------------------------------------------------------
// Pendulum.  x = [theta, thetadot], theta = 0 hanging down.
// p = [dt, mass, length, damping, g]
struct Pendulum {
  static constexpr int n = 2, m = 1, np = 5;
  static constexpr int COOP = 1;
  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const double h = p[0], ml2 = p[1] * p[2] * p[2], mgl = p[1] * p[4] * p[2];
    S s, c;
    sincos_(x[0], &s, &c);
    S a = (u[0] - p[3] * x[1] - mgl * s) / ml2;
    S v1 = x[1] + h * a;
    xn[0] = x[0] + h * v1;
    xn[1] = v1;
  }
};

// ------------------------------------------------------------------------------
// Acrobot (elbow actuated).  x = [q1, q2, q1dot, q2dot].
// p = [dt, m1, m2, l1, lc1, lc2, Ic1, Ic2, b1, b2, g]
struct Acrobot {
  static constexpr int n = 4, m = 1, np = 11;
  static constexpr int COOP = 1;
  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const double h = p[0], m1 = p[1], m2 = p[2], l1 = p[3], lc1 = p[4], lc2 = p[5];
    const double I1 = p[6] + m1 * lc1 * lc1, I2 = p[7] + m2 * lc2 * lc2;
    const double b1 = p[8], b2 = p[9], g = p[10];
    const double k = m2 * l1 * lc2;
    S s1, c1, s2, c2, s12, c12;
    sincos_(x[0], &s1, &c1);
    sincos_(x[1], &s2, &c2);
    sincos_(x[0] + x[1], &s12, &c12);
    (void)c1;
    (void)c12;
    S M11 = (I1 + I2 + m2 * l1 * l1) + (2.0 * k) * c2;
    S M12 = I2 + k * c2;
    const double M22 = I2;
    // M qdd = tau_g + B u - C qd - b qd
    S hq = k * s2;
    S r1 = hq * x[3] * (2.0 * x[2] + x[3]) - (m1 * g * lc1 + m2 * g * l1) * s1 -
           (m2 * g * lc2) * s12 - b1 * x[2];
    S r2 = u[0] - hq * x[2] * x[2] - (m2 * g * lc2) * s12 - b2 * x[3];
    S det = M11 * M22 - M12 * M12;
    S a1 = (M22 * r1 - M12 * r2) / det;
    S a2 = (M11 * r2 - M12 * r1) / det;
    S v1 = x[2] + h * a1, v2 = x[3] + h * a2;
    xn[0] = x[0] + h * v1;
    xn[1] = x[1] + h * v2;
    xn[2] = v1;
    xn[3] = v2;
  }
};

// ------------------------------------------------------------------------------
// Cart-pole, optionally with a compliant ball on the pole tip hitting a rigid
// wall.  x = [cart x, theta, xdot, thetadot], theta = 0 hanging, pi upright.
// p = [dt, mc, mp, l, g, ball_radius, modulus E, wall_face_x, substeps]
template <bool WALL>
struct CartPoleT {
  static constexpr int n = 4, m = 1, np = 9;
  static constexpr int COOP = 1;
  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const double mc = p[1], mp = p[2], l = p[3], g = p[4];
    const int sub = WALL ? (int)p[8] : 1;
    const double h = p[0] / sub;
    S q0 = x[0], q1 = x[1], v0 = x[2], v1 = x[3];
    for (int it = 0; it < sub; ++it) {
      S s, c;
      sincos_(q1, &s, &c);
      S r0 = u[0] + (mp * l) * v1 * v1 * s;
      S r1 = -(mp * g * l) * s;
      if (WALL) {
        // tip ball: centre at x + l sin(theta); wall occupies x <= wall_face_x.
        const double R = p[5], E = p[6], xw = p[7];
        S depth = (xw + R) - (q0 + l * s);
        if (val(depth) > 0.0) {
          S F;
          if (val(depth) < R) {
            F = (3.14159265358979323846 * E) * depth * depth * (1.0 - depth * (2.0 / (3.0 * R)));
          } else {
            F = S((3.14159265358979323846 * E) * R * R / 3.0) + 0.0 * depth;
          }
          r0 = r0 + F;
          r1 = r1 + F * (l * c);
        }
      }
      const double M00 = mc + mp, M11 = mp * l * l;
      S M01 = (mp * l) * c;
      S det = M00 * M11 - M01 * M01;
      S a0 = (M11 * r0 - M01 * r1) / det;
      S a1 = (M00 * r1 - M01 * r0) / det;
      v0 = v0 + h * a0;
      v1 = v1 + h * a1;
      q0 = q0 + h * v0;
      q1 = q1 + h * v1;
    }
    xn[0] = q0;
    xn[1] = q1;
    xn[2] = v0;
    xn[3] = v1;
  }
};
typedef CartPoleT<false> CartPole;
typedef CartPoleT<true> CartPoleWall;

// ------------------------------------------------------------------------------
// Compliant sphere / rigid plane normal force, depth >= 0, saturating at R.
// k23R = 2 / (3 R): a division is ~25 dependent instructions and sits on the serial path of a
// contact step, so callers that evaluate the force every substep pass the quotient in (the
// quadruped keeps it in its parameter vector, rounded once on the host: same value, same result).
template <class S>
DDP_HD S sphere_plane_force(const S& depth, double R, double E, double k23R) {
  const double piE = 3.14159265358979323846 * E;
  if (val(depth) <= 0.0) return 0.0 * depth;
  if (val(depth) >= R) return S(piE * R * R / 3.0) + 0.0 * depth;
  return piE * depth * depth * (1.0 - depth * k23R);
}
template <class S>
DDP_HD S sphere_plane_force(const S& depth, double R, double E) {
  return sphere_plane_force(depth, R, E, 2.0 / (3.0 * R));
}

// ------------------------------------------------------------------------------
// Quadruped, mini_cheetah scale: one rigid body (all link masses lumped) on four
// 3-joint legs whose joint dynamics are rotor-inertia dominated; spherical feet
// on compliant ground with regularised Coulomb friction.
//   q = [px py pz | roll pitch yaw | (abad hip knee) x {FR, FL, HR, HL}]   (18)
//   v = [world linear velocity | body angular velocity | joint rates]        (18)
// p = [dt, substeps, mass, Ixx, Iyy, Izz, Ij_abad, Ij_hip, Ij_knee, joint_damping,
//      l_abad, l_thigh, l_shank, hip_x, hip_y, foot_radius, E, mu, v_stiction, g,
//      1/mass, 1/Ixx, 1/Iyy, 1/Izz, 1/Ij_abad, 1/Ij_hip, 1/Ij_knee, 2/(3 foot_radius)]
// The model divides by nothing but cos(pitch) and the slip speed: masses and inertias enter through
// their reciprocals (p[20..26], rounded once on the host), because an fp64 division is ~25 dependent
// instructions on the serial path of every rollout step.
struct Quadruped {
  static constexpr int n = 36, m = 12, np = 28;
  static constexpr int COOP = 4;  // step_coop: one leg per lane of a 4-lane group

  template <class S>
  struct LegOut {
    S Fx, Fy, Fz;  // world force on the body
    S Tx, Ty, Tz;  // body-frame torque on the body
    S a0, a1, a2;  // joint accelerations (abad, hip, knee)
  };
  template <class S>
  struct BasePose {
    S R00, R01, R02, R10, R11, R12, R20, R21, R22, sr, cr, sp, cp;
  };
  template <class S>
  DDP_HD static void base_pose(const S* q, BasePose<S>& B) {
    S sy, cy;
    sincos_(q[3], &B.sr, &B.cr);
    sincos_(q[4], &B.sp, &B.cp);
    sincos_(q[5], &sy, &cy);
    base_pose_trig(sy, cy, B);
  }
  // rotation matrix from the sines / cosines (B.sr, B.cr, B.sp, B.cp already set)
  template <class S>
  DDP_HD static void base_pose_trig(const S& sy, const S& cy, BasePose<S>& B) {
    // R = Rz(yaw) Ry(pitch) Rx(roll), body -> world
    B.R00 = cy * B.cp; B.R01 = cy * B.sp * B.sr - sy * B.cr; B.R02 = cy * B.sp * B.cr + sy * B.sr;
    B.R10 = sy * B.cp; B.R11 = sy * B.sp * B.sr + cy * B.cr; B.R12 = sy * B.sp * B.cr - cy * B.sr;
    B.R20 = -B.sp; B.R21 = B.cp * B.sr; B.R22 = B.cp * B.cr;
  }
  // One leg: foot kinematics, compliant ground contact, joint accelerations.
  // sx = +1 front / -1 hind, sd = +1 left / -1 right; (qa,qh,qk) joint angles, (va,vh,vk) rates.
  template <class S>
  DDP_HD static void leg(double sx, double sd, const S& qa, const S& qh, const S& qk, const S& va,
                         const S& vh, const S& vk, const S& ua, const S& uh, const S& uk, const S& pz,
                         const S* v, const BasePose<S>& B, const double* p, LegOut<S>& o) {
    S sa, ca, sh, ch, sk, ck;
    sincos_(qa, &sa, &ca);
    sincos_(qh, &sh, &ch);
    sincos_(qh + qk, &sk, &ck);
    leg_trig(sx, sd, sa, ca, sh, ch, sk, ck, va, vh, vk, ua, uh, uk, pz, v, B, p, o);
  }
  // the same with the sines / cosines of abad, hip and hip + knee supplied by the caller (the
  // fused linearization computes each of them once per warp)
  template <class S>
  DDP_HD static void leg_trig(double sx, double sd, const S& sa, const S& ca, const S& sh, const S& ch, const S& sk,
                              const S& ck, const S& va, const S& vh, const S& vk, const S& ua, const S& uh,
                              const S& uk, const S& pz, const S* v, const BasePose<S>& B, const double* p,
                              LegOut<S>& o) {
    const double bj = p[9], l1 = p[10], l2 = p[11], l3 = p[12];
    const double hx = p[13], hy = p[14], rf = p[15], E = p[16], mu = p[17], vs = p[18];
    const double ly = sd * l1;
    S lx = -(l2 * sh) - l3 * sk;
    S lz = -(l2 * ch) - l3 * ck;
    // foot in body frame and its joint Jacobian (columns: abad, hip, knee)
    S rx = hx * sx + lx;
    S ry = hy * sd + (ly * ca - lz * sa);
    S rz = ly * sa + lz * ca;
    S J01 = lz, J02 = -(l3 * ck);
    S J10 = -(ly * sa) - lz * ca, J11 = lx * sa, J12 = -(l3 * sk) * sa;
    S J20 = ly * ca - lz * sa, J21 = -(lx * ca), J22 = (l3 * sk) * ca;
    S cz = pz + B.R20 * rx + B.R21 * ry + B.R22 * rz;
    S depth = rf - cz;
    S tau0 = S(0.0), tau1 = S(0.0), tau2 = S(0.0);
    o.Fx = S(0.0); o.Fy = S(0.0); o.Fz = S(0.0);
    o.Tx = S(0.0); o.Ty = S(0.0); o.Tz = S(0.0);
    if (val(depth) > 0.0) {
      // body-frame foot velocity relative to the body origin: w x r + J qd
      S bx = v[4] * rz - v[5] * ry + J01 * vh + J02 * vk;
      S by = v[5] * rx - v[3] * rz + J10 * va + J11 * vh + J12 * vk;
      S bz = v[3] * ry - v[4] * rx + J20 * va + J21 * vh + J22 * vk;
      S wx = v[0] + B.R00 * bx + B.R01 * by + B.R02 * bz;
      S wy = v[1] + B.R10 * bx + B.R11 * by + B.R12 * bz;
      S Fn = sphere_plane_force(depth, rf, E, p[27]);
      S isl = inv_sqrt_(wx * wx + wy * wy + vs * vs);
      S ftn = (mu * Fn) * isl;
      S ftx = -(ftn * wx);
      S fty = -(ftn * wy);
      o.Fx = ftx; o.Fy = fty; o.Fz = Fn;
      S fbx = B.R00 * ftx + B.R10 * fty + B.R20 * Fn;   // force in body frame
      S fby = B.R01 * ftx + B.R11 * fty + B.R21 * Fn;
      S fbz = B.R02 * ftx + B.R12 * fty + B.R22 * Fn;
      o.Tx = ry * fbz - rz * fby;
      o.Ty = rz * fbx - rx * fbz;
      o.Tz = rx * fby - ry * fbx;
      tau0 = J10 * fby + J20 * fbz;
      tau1 = J01 * fbx + J11 * fby + J21 * fbz;
      tau2 = J02 * fbx + J12 * fby + J22 * fbz;
    }
    o.a0 = (ua + tau0 - bj * va) * p[24];
    o.a1 = (uh + tau1 - bj * vh) * p[25];
    o.a2 = (uk + tau2 - bj * vk) * p[26];
  }
  // base accelerations from the summed leg loads, then the semi-implicit Euler update
  template <class S>
  DDP_HD static void integrate(S* q, S* v, const S* acc, const BasePose<S>& B, double h) {
#pragma unroll
    for (int i = 0; i < 18; ++i) v[i] = v[i] + h * acc[i];
    // q+ = q + h N(q) v+   (ZYX Euler rates from body angular velocity)
    q[0] = q[0] + h * v[0];
    q[1] = q[1] + h * v[1];
    q[2] = q[2] + h * v[2];
    S icp = 1.0 / B.cp;
    S tp = B.sp * icp;
    S wyz = B.sr * v[4] + B.cr * v[5];
    q[3] = q[3] + h * (v[3] + tp * wyz);
    q[4] = q[4] + h * (B.cr * v[4] - B.sr * v[5]);
    q[5] = q[5] + h * (wyz * icp);
#pragma unroll
    for (int i = 6; i < 18; ++i) q[i] = q[i] + h * v[i];
  }
  template <class S>
  DDP_HD static void base_acc(const S& Fx, const S& Fy, const S& Fz, const S& Tx, const S& Ty, const S& Tz,
                              const S* v, const double* p, S* acc) {
    const double Ix = p[3], Iy = p[4], Iz = p[5], g = p[19];
    acc[0] = Fx * p[20];
    acc[1] = Fy * p[20];
    acc[2] = Fz * p[20] - g;
    acc[3] = (Tx - (Iz - Iy) * v[4] * v[5]) * p[21];
    acc[4] = (Ty - (Ix - Iz) * v[5] * v[3]) * p[22];
    acc[5] = (Tz - (Iy - Ix) * v[3] * v[4]) * p[23];
  }

  // one semi-implicit Euler substep of length h, in place.  Leg loads are summed pairwise,
  // (leg0 + leg1) + (leg2 + leg3), the association of the 4-lane butterfly in step_coop; only two
  // legs' outputs are alive at a time to keep the dual-number instantiation in registers.
  template <class S>
  DDP_HD static void substep(S* q, S* v, const S* u, const double* p, double h) {
    BasePose<S> B;
    base_pose(q, B);
    S acc[18];
    S sum[2][6];
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {
      LegOut<S> o[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int l = 2 * pair + k;
        leg((l < 2) ? 1.0 : -1.0, (l & 1) ? 1.0 : -1.0, q[6 + 3 * l], q[7 + 3 * l], q[8 + 3 * l],
            v[6 + 3 * l], v[7 + 3 * l], v[8 + 3 * l], u[3 * l], u[3 * l + 1], u[3 * l + 2], q[2], v, B, p,
            o[k]);
        acc[6 + 3 * l] = o[k].a0;
        acc[7 + 3 * l] = o[k].a1;
        acc[8 + 3 * l] = o[k].a2;
      }
      sum[pair][0] = o[0].Fx + o[1].Fx;
      sum[pair][1] = o[0].Fy + o[1].Fy;
      sum[pair][2] = o[0].Fz + o[1].Fz;
      sum[pair][3] = o[0].Tx + o[1].Tx;
      sum[pair][4] = o[0].Ty + o[1].Ty;
      sum[pair][5] = o[0].Tz + o[1].Tz;
    }
    base_acc(sum[0][0] + sum[1][0], sum[0][1] + sum[1][1], sum[0][2] + sum[1][2], sum[0][3] + sum[1][3],
             sum[0][4] + sum[1][4], sum[0][5] + sum[1][5], v, p, acc);
    integrate(q, v, acc, B, h);
  }

  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    S q[18], v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      q[i] = x[i];
      v[i] = x[18 + i];
    }
    for (int it = 0; it < sub; ++it) substep(q, v, u, p, h);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      xn[i] = q[i];
      xn[18 + i] = v[i];
    }
  }

#if defined(__CUDACC__)
  // Same map evaluated by a 4-lane group: lane l computes leg l, loads are combined with a
  // two-stage butterfly, joint accelerations are exchanged with shuffles.  Bit-identical to
  // step<double>() (same operations, same association of the sums).
  __device__ __forceinline__ static double pick4(int l, double a, double b, double c, double d) {
    return l == 0 ? a : (l == 1 ? b : (l == 2 ? c : d));
  }
  // base_pose() with the three sincos pairs dealt to lanes 0..2 of the group and exchanged by
  // shuffle: same values as the scalar version (same inputs, same function), a third of the trig.
  __device__ __forceinline__ static void base_pose_coop(int lane, unsigned mask, int gbase, const double* q,
                                                        BasePose<double>& B) {
    double sn, cs;
    sincos_(pick4(lane, q[3], q[4], q[5], q[5]), &sn, &cs);
    B.sr = __shfl_sync(mask, sn, gbase + 0);
    B.cr = __shfl_sync(mask, cs, gbase + 0);
    B.sp = __shfl_sync(mask, sn, gbase + 1);
    B.cp = __shfl_sync(mask, cs, gbase + 1);
    const double sy = __shfl_sync(mask, sn, gbase + 2), cy = __shfl_sync(mask, cs, gbase + 2);
    B.R00 = cy * B.cp; B.R01 = cy * B.sp * B.sr - sy * B.cr; B.R02 = cy * B.sp * B.cr + sy * B.sr;
    B.R10 = sy * B.cp; B.R11 = sy * B.sp * B.sr + cy * B.cr; B.R12 = sy * B.sp * B.cr - cy * B.sr;
    B.R20 = -B.sp; B.R21 = B.cp * B.sr; B.R22 = B.cp * B.cr;
  }
  __device__ __forceinline__ static void step_coop(int lane, unsigned mask, int gbase, const double* x,
                                                   const double* u, double* xn, const double* p) {
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    double q[18], v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      q[i] = x[i];
      v[i] = x[18 + i];
    }
    const double ua = pick4(lane, u[0], u[3], u[6], u[9]);
    const double uh = pick4(lane, u[1], u[4], u[7], u[10]);
    const double uk = pick4(lane, u[2], u[5], u[8], u[11]);
    const double sx = (lane < 2) ? 1.0 : -1.0, sd = (lane & 1) ? 1.0 : -1.0;
    for (int it = 0; it < sub; ++it) {
      BasePose<double> B;
      base_pose_coop(lane, mask, gbase, q, B);
      LegOut<double> o;
      leg(sx, sd, pick4(lane, q[6], q[9], q[12], q[15]), pick4(lane, q[7], q[10], q[13], q[16]),
          pick4(lane, q[8], q[11], q[14], q[17]), pick4(lane, v[6], v[9], v[12], v[15]),
          pick4(lane, v[7], v[10], v[13], v[16]), pick4(lane, v[8], v[11], v[14], v[17]), ua, uh, uk, q[2],
          v, B, p, o);
      double f[6] = {o.Fx, o.Fy, o.Fz, o.Tx, o.Ty, o.Tz};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(mask, f[k], 1);
        f[k] += __shfl_xor_sync(mask, f[k], 2);
      }
      double acc[18];
      base_acc(f[0], f[1], f[2], f[3], f[4], f[5], v, p, acc);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        acc[6 + 3 * l] = __shfl_sync(mask, o.a0, gbase + l);
        acc[7 + 3 * l] = __shfl_sync(mask, o.a1, gbase + l);
        acc[8 + 3 * l] = __shfl_sync(mask, o.a2, gbase + l);
      }
      integrate(q, v, acc, B, h);
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      xn[i] = q[i];
      xn[18 + i] = v[i];
    }
  }
#endif
};

// ------------------------------------------------------------------------------
// The same quadruped with a quaternion floating base, in the reference's state layout
// (mini_cheetah.py:41-57: n = 37):
//   q = [qw qx qy qz | px py pz | (abad hip knee) x {FR, FL, HR, HL}]            (19)
//   v = [world angular velocity | world linear velocity | joint rates]           (18)
// so x[4] is the base x position and x[22] the base x velocity, as the script assumes.
// The rotation uses the normalised quaternion; the state itself is not renormalised.
// Same parameter vector as Quadruped.
struct QuadrupedQuat {
  static constexpr int n = 37, m = 12, np = 28;
  static constexpr int COOP = 4;  // step_coop: one leg per lane of a 4-lane group
  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    typedef Quadruped Qd;
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    const double Ix = p[3], Iy = p[4], Iz = p[5];
    S qt[4], pos[3], qj[12], w[3], vl[3], vj[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) qt[i] = x[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pos[i] = x[4 + i];
      w[i] = x[19 + i];
      vl[i] = x[22 + i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      qj[i] = x[7 + i];
      vj[i] = x[25 + i];
    }
    for (int it = 0; it < sub; ++it) {
      // rotation matrix of the normalised quaternion (body -> world)
      S inn = inv_sqrt_(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]);
      S a = qt[0] * inn, b = qt[1] * inn, c = qt[2] * inn, d = qt[3] * inn;
      Qd::BasePose<S> B;
      B.R00 = 1.0 - 2.0 * (c * c + d * d); B.R01 = 2.0 * (b * c - a * d); B.R02 = 2.0 * (b * d + a * c);
      B.R10 = 2.0 * (b * c + a * d); B.R11 = 1.0 - 2.0 * (b * b + d * d); B.R12 = 2.0 * (c * d - a * b);
      B.R20 = 2.0 * (b * d - a * c); B.R21 = 2.0 * (c * d + a * b); B.R22 = 1.0 - 2.0 * (b * b + c * c);
      B.sr = S(0.0); B.cr = S(1.0); B.sp = S(0.0); B.cp = S(1.0);
      // leg() wants v = [world linear velocity | BODY angular velocity]
      S vloc[6];
      vloc[0] = vl[0]; vloc[1] = vl[1]; vloc[2] = vl[2];
      vloc[3] = B.R00 * w[0] + B.R10 * w[1] + B.R20 * w[2];
      vloc[4] = B.R01 * w[0] + B.R11 * w[1] + B.R21 * w[2];
      vloc[5] = B.R02 * w[0] + B.R12 * w[1] + B.R22 * w[2];
      S aj[12], sum[2][6];
#pragma unroll
      for (int pair = 0; pair < 2; ++pair) {
        Qd::LegOut<S> o[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int l = 2 * pair + k;
          Qd::leg((l < 2) ? 1.0 : -1.0, (l & 1) ? 1.0 : -1.0, qj[3 * l], qj[3 * l + 1], qj[3 * l + 2], vj[3 * l],
                  vj[3 * l + 1], vj[3 * l + 2], u[3 * l], u[3 * l + 1], u[3 * l + 2], pos[2], vloc, B, p, o[k]);
          aj[3 * l] = o[k].a0;
          aj[3 * l + 1] = o[k].a1;
          aj[3 * l + 2] = o[k].a2;
        }
        sum[pair][0] = o[0].Fx + o[1].Fx; sum[pair][1] = o[0].Fy + o[1].Fy; sum[pair][2] = o[0].Fz + o[1].Fz;
        sum[pair][3] = o[0].Tx + o[1].Tx; sum[pair][4] = o[0].Ty + o[1].Ty; sum[pair][5] = o[0].Tz + o[1].Tz;
      }
      // body-frame Euler equations, then back to the world frame: wdot_W = R wdot_B
      S Tx = sum[0][3] + sum[1][3], Ty = sum[0][4] + sum[1][4], Tz = sum[0][5] + sum[1][5];
      S ab0 = (Tx - (Iz - Iy) * vloc[4] * vloc[5]) * p[21];
      S ab1 = (Ty - (Ix - Iz) * vloc[5] * vloc[3]) * p[22];
      S ab2 = (Tz - (Iy - Ix) * vloc[3] * vloc[4]) * p[23];
      w[0] = w[0] + h * (B.R00 * ab0 + B.R01 * ab1 + B.R02 * ab2);
      w[1] = w[1] + h * (B.R10 * ab0 + B.R11 * ab1 + B.R12 * ab2);
      w[2] = w[2] + h * (B.R20 * ab0 + B.R21 * ab1 + B.R22 * ab2);
      vl[0] = vl[0] + h * ((sum[0][0] + sum[1][0]) * p[20]);
      vl[1] = vl[1] + h * ((sum[0][1] + sum[1][1]) * p[20]);
      vl[2] = vl[2] + h * ((sum[0][2] + sum[1][2]) * p[20] - p[19]);
#pragma unroll
      for (int i = 0; i < 12; ++i) vj[i] = vj[i] + h * aj[i];
      // qdot = 0.5 * (0, w_W) (x) q
      S q0 = qt[0], q1 = qt[1], q2 = qt[2], q3 = qt[3];
      qt[0] = q0 + (0.5 * h) * (-(w[0] * q1) - w[1] * q2 - w[2] * q3);
      qt[1] = q1 + (0.5 * h) * (w[0] * q0 + w[1] * q3 - w[2] * q2);
      qt[2] = q2 + (0.5 * h) * (w[1] * q0 + w[2] * q1 - w[0] * q3);
      qt[3] = q3 + (0.5 * h) * (w[2] * q0 + w[0] * q2 - w[1] * q1);
#pragma unroll
      for (int i = 0; i < 3; ++i) pos[i] = pos[i] + h * vl[i];
#pragma unroll
      for (int i = 0; i < 12; ++i) qj[i] = qj[i] + h * vj[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = qt[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xn[4 + i] = pos[i];
      xn[19 + i] = w[i];
      xn[22 + i] = vl[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      xn[7 + i] = qj[i];
      xn[25 + i] = vj[i];
    }
  }

#if defined(__CUDACC__)
  // The same map evaluated by a 4-lane group: lane l computes leg l, loads are combined with the
  // butterfly (leg0 + leg1) + (leg2 + leg3), joint accelerations are exchanged with shuffles;
  // everything else is evaluated redundantly.  Bit-identical to step<double>().
  __device__ __forceinline__ static void step_coop(int lane, unsigned mask, int gbase, const double* x,
                                                   const double* u, double* xn, const double* p) {
    typedef Quadruped Qd;
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    const double Ix = p[3], Iy = p[4], Iz = p[5];
    double qt[4], pos[3], qj[12], w[3], vl[3], vj[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) qt[i] = x[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pos[i] = x[4 + i];
      w[i] = x[19 + i];
      vl[i] = x[22 + i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      qj[i] = x[7 + i];
      vj[i] = x[25 + i];
    }
    const double ua = Qd::pick4(lane, u[0], u[3], u[6], u[9]);
    const double uh = Qd::pick4(lane, u[1], u[4], u[7], u[10]);
    const double uk = Qd::pick4(lane, u[2], u[5], u[8], u[11]);
    const double sx = (lane < 2) ? 1.0 : -1.0, sd = (lane & 1) ? 1.0 : -1.0;
    for (int it = 0; it < sub; ++it) {
      double inn = inv_sqrt_(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]);
      double a = qt[0] * inn, b = qt[1] * inn, c = qt[2] * inn, d = qt[3] * inn;
      Qd::BasePose<double> B;
      B.R00 = 1.0 - 2.0 * (c * c + d * d); B.R01 = 2.0 * (b * c - a * d); B.R02 = 2.0 * (b * d + a * c);
      B.R10 = 2.0 * (b * c + a * d); B.R11 = 1.0 - 2.0 * (b * b + d * d); B.R12 = 2.0 * (c * d - a * b);
      B.R20 = 2.0 * (b * d - a * c); B.R21 = 2.0 * (c * d + a * b); B.R22 = 1.0 - 2.0 * (b * b + c * c);
      B.sr = 0.0; B.cr = 1.0; B.sp = 0.0; B.cp = 1.0;
      double vloc[6];
      vloc[0] = vl[0]; vloc[1] = vl[1]; vloc[2] = vl[2];
      vloc[3] = B.R00 * w[0] + B.R10 * w[1] + B.R20 * w[2];
      vloc[4] = B.R01 * w[0] + B.R11 * w[1] + B.R21 * w[2];
      vloc[5] = B.R02 * w[0] + B.R12 * w[1] + B.R22 * w[2];
      Qd::LegOut<double> o;
      Qd::leg(sx, sd, Qd::pick4(lane, qj[0], qj[3], qj[6], qj[9]), Qd::pick4(lane, qj[1], qj[4], qj[7], qj[10]),
              Qd::pick4(lane, qj[2], qj[5], qj[8], qj[11]), Qd::pick4(lane, vj[0], vj[3], vj[6], vj[9]),
              Qd::pick4(lane, vj[1], vj[4], vj[7], vj[10]), Qd::pick4(lane, vj[2], vj[5], vj[8], vj[11]), ua, uh, uk,
              pos[2], vloc, B, p, o);
      double f[6] = {o.Fx, o.Fy, o.Fz, o.Tx, o.Ty, o.Tz};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        f[k] += __shfl_xor_sync(mask, f[k], 1);
        f[k] += __shfl_xor_sync(mask, f[k], 2);
      }
      double aj[12];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        aj[3 * l] = __shfl_sync(mask, o.a0, gbase + l);
        aj[3 * l + 1] = __shfl_sync(mask, o.a1, gbase + l);
        aj[3 * l + 2] = __shfl_sync(mask, o.a2, gbase + l);
      }
      double ab0 = (f[3] - (Iz - Iy) * vloc[4] * vloc[5]) * p[21];
      double ab1 = (f[4] - (Ix - Iz) * vloc[5] * vloc[3]) * p[22];
      double ab2 = (f[5] - (Iy - Ix) * vloc[3] * vloc[4]) * p[23];
      w[0] = w[0] + h * (B.R00 * ab0 + B.R01 * ab1 + B.R02 * ab2);
      w[1] = w[1] + h * (B.R10 * ab0 + B.R11 * ab1 + B.R12 * ab2);
      w[2] = w[2] + h * (B.R20 * ab0 + B.R21 * ab1 + B.R22 * ab2);
      vl[0] = vl[0] + h * (f[0] * p[20]);
      vl[1] = vl[1] + h * (f[1] * p[20]);
      vl[2] = vl[2] + h * (f[2] * p[20] - p[19]);
#pragma unroll
      for (int i = 0; i < 12; ++i) vj[i] = vj[i] + h * aj[i];
      double q0 = qt[0], q1 = qt[1], q2 = qt[2], q3 = qt[3];
      qt[0] = q0 + (0.5 * h) * (-(w[0] * q1) - w[1] * q2 - w[2] * q3);
      qt[1] = q1 + (0.5 * h) * (w[0] * q0 + w[1] * q3 - w[2] * q2);
      qt[2] = q2 + (0.5 * h) * (w[1] * q0 + w[2] * q1 - w[0] * q3);
      qt[3] = q3 + (0.5 * h) * (w[2] * q0 + w[0] * q2 - w[1] * q1);
#pragma unroll
      for (int i = 0; i < 3; ++i) pos[i] = pos[i] + h * vl[i];
#pragma unroll
      for (int i = 0; i < 12; ++i) qj[i] = qj[i] + h * vj[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = qt[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xn[4 + i] = pos[i];
      xn[19 + i] = w[i];
      xn[22 + i] = vl[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      xn[7 + i] = qj[i];
      xn[25 + i] = vj[i];
    }
  }
#endif
};

// ------------------------------------------------------------------------------
// Arm pushing a ball on a table, kinova_gen3 / panda_fr3 scale (7 + 7 + 13):
//   q = [7 joint angles | ball quaternion (w x y z) | ball position]   (14)
//   v = [7 joint rates  | ball angular velocity (world) | ball linear velocity] (13)
// The arm is a 7R chain (alternating z / y joint axes, link lengths d[0..6]) whose
// joint dynamics are rotor-inertia dominated with gravity compensated; a compliant
// sphere at the tool tip pushes a free ball (sphere/sphere contact) that rests on a
// compliant table (sphere/plane contact) with regularised friction at both contacts.
// Normal forces carry Hunt-Crossley dissipation: Fn = Fe(depth) * max(0, 1 + d * depth_rate).
// p = [dt, substeps, Ij, joint_damping, d0..d6 (7), tip_radius, ball_radius, ball_mass,
//      E, mu, v_stiction, g, base_z, dissipation,
//      Re = rt rb / (rt + rb), 2 / (3 Re), 2 / (3 rb)]
struct ArmBall {
  static constexpr int n = 27, m = 7, np = 23;
  static constexpr int COOP = 1;

  // tool frame while walking up the chain: rotation (rows x, y, z of the world axes) and origin
  template <class S>
  struct Frame {
    S Rxx, Rxy, Rxz, Ryx, Ryy, Ryz, Rzx, Rzy, Rzz, px, py, pz;
  };
  template <class S>
  DDP_HD static void fk_init(const S& q0, double bz, Frame<S>& F) {
    F.Rxx = 1.0 + 0.0 * q0; F.Rxy = 0.0 * q0; F.Rxz = F.Rxy;
    F.Ryx = F.Rxy; F.Ryy = F.Rxx; F.Ryz = F.Rxy;
    F.Rzx = F.Rxy; F.Rzy = F.Rxy; F.Rzz = F.Rxx;
    F.px = F.Rxy; F.py = F.Rxy; F.pz = F.Rxy + bz;
  }
  // joint i: rotate about axis a_i (local z for even i, local y for odd i) by the angle with sine s
  // and cosine c, then go di along the rotated local z.  Returns the joint axis in the world frame;
  // the joint origin is (F.px, F.py, F.pz) BEFORE the call.
  template <class S>
  DDP_HD static void fk_joint(int i, const S& s, const S& c, double di, Frame<S>& F, S& ax, S& ay, S& az) {
    if ((i & 1) == 0) {  // about local z
      ax = F.Rxz; ay = F.Ryz; az = F.Rzz;
      S nxx = F.Rxx * c + F.Rxy * s, nxy = F.Rxy * c - F.Rxx * s;
      S nyx = F.Ryx * c + F.Ryy * s, nyy = F.Ryy * c - F.Ryx * s;
      S nzx = F.Rzx * c + F.Rzy * s, nzy = F.Rzy * c - F.Rzx * s;
      F.Rxx = nxx; F.Rxy = nxy; F.Ryx = nyx; F.Ryy = nyy; F.Rzx = nzx; F.Rzy = nzy;
    } else {  // about local y
      ax = F.Rxy; ay = F.Ryy; az = F.Rzy;
      S nxx = F.Rxx * c - F.Rxz * s, nxz = F.Rxz * c + F.Rxx * s;
      S nyx = F.Ryx * c - F.Ryz * s, nyz = F.Ryz * c + F.Ryx * s;
      S nzx = F.Rzx * c - F.Rzz * s, nzz = F.Rzz * c + F.Rzx * s;
      F.Rxx = nxx; F.Rxz = nxz; F.Ryx = nyx; F.Ryz = nyz; F.Rzx = nzx; F.Rzz = nzz;
    }
    F.px = F.px + di * F.Rxz;
    F.py = F.py + di * F.Ryz;
    F.pz = F.pz + di * F.Rzz;
  }
  // contact loads: force on the tool tip, force and world-frame torque on the ball
  template <class S>
  struct Loads {
    S ftx, fty, ftz, fbx, fby, fbz, tbx, tby, tbz;
  };
  // (px, py, pz) tool tip, (tvx, tvy, tvz) its velocity, (bx_, by_, bz_) ball centre, w / bv ball
  // angular / linear velocity
  template <class S>
  DDP_HD static void contacts(const S& px, const S& py, const S& pz, const S& tvx, const S& tvy, const S& tvz,
                              const S& bx_, const S& by_, const S& bz_, const S* w, const S* bv, const double* p,
                              Loads<S>& o) {
    const double rt = p[11], rb = p[12], mb = p[13], E = p[14], mu = p[15], vs = p[16];
    const double g = p[17], diss = p[19];
    // functions of the parameters alone, rounded once on the host (systems.arm_ball): three divisions
    // per substep otherwise (the compiler cannot move loads of p[] out of a loop that stores)
    const double Re = p[20], k23Re = p[21], k23rb = p[22];   // rt rb / (rt + rb), 2 / (3 Re), 2 / (3 rb)
    S fbx = 0.0 * px, fby = fbx, fbz = fbx - mb * g;  // force on ball
    S tbx = 0.0 * px, tby = tbx, tbz = tbx;           // torque on ball (world)
    S ftx = 0.0 * px, fty = ftx, ftz = ftx;           // force on tool tip
    // tip sphere vs ball
    {
      S nx = bx_ - px, ny = by_ - py, nz = bz_ - pz;
      S idist;
      S dist = sqrt_pair_(nx * nx + ny * ny + nz * nz + 1e-12, &idist);
      S depth = (rt + rb) - dist;
      if (val(depth) > 0.0) {
        nx = div_root_(nx, dist, idist); ny = div_root_(ny, dist, idist); nz = div_root_(nz, dist, idist);  // tip -> ball
        // relative velocity of ball surface point w.r.t. tip at the contact
        S cxr = -(rb)*nx, cyr = -(rb)*ny, czr = -(rb)*nz;  // contact point rel. ball centre
        S rvx = bv[0] + (w[1] * czr - w[2] * cyr) - tvx;
        S rvy = bv[1] + (w[2] * cxr - w[0] * czr) - tvy;
        S rvz = bv[2] + (w[0] * cyr - w[1] * cxr) - tvz;
        S vn = rvx * nx + rvy * ny + rvz * nz;   // separation rate = -depth rate
        S Fn = sphere_plane_force(depth, Re, E, k23Re);
        S hc = 1.0 - diss * vn;
        if (val(hc) < 0.0) hc = S(0.0);
        Fn = Fn * hc;
        S tx = rvx - vn * nx, ty = rvy - vn * ny, tz = rvz - vn * nz;
        S isl;
        S sl = sqrt_pair_(tx * tx + ty * ty + tz * tz + vs * vs, &isl);
        S cfx = Fn * nx - div_root_((mu * Fn) * tx, sl, isl);
        S cfy = Fn * ny - div_root_((mu * Fn) * ty, sl, isl);
        S cfz = Fn * nz - div_root_((mu * Fn) * tz, sl, isl);
        fbx = fbx + cfx; fby = fby + cfy; fbz = fbz + cfz;
        tbx = tbx + (cyr * cfz - czr * cfy);
        tby = tby + (czr * cfx - cxr * cfz);
        tbz = tbz + (cxr * cfy - cyr * cfx);
        ftx = ftx - cfx; fty = fty - cfy; ftz = ftz - cfz;
      }
    }
    // ball vs table (z = 0)
    {
      S depth = rb - bz_;
      if (val(depth) > 0.0) {
        S Fn = sphere_plane_force(depth, rb, E, k23rb);
        S hc = 1.0 - diss * bv[2];
        if (val(hc) < 0.0) hc = S(0.0);
        Fn = Fn * hc;
        // contact point velocity: v + w x (0,0,-rb)
        S cvx = bv[0] - w[1] * rb;
        S cvy = bv[1] + w[0] * rb;
        S isl;
        S sl = sqrt_pair_(cvx * cvx + cvy * cvy + vs * vs, &isl);
        S fx_ = div_root_(-(mu * Fn) * cvx, sl, isl), fy_ = div_root_(-(mu * Fn) * cvy, sl, isl);
        fbx = fbx + fx_; fby = fby + fy_; fbz = fbz + Fn;
        // torque = (0,0,-rb) x (fx, fy, Fn)
        tbx = tbx + rb * fy_;
        tby = tby - rb * fx_;
      }
    }
    o.ftx = ftx; o.fty = fty; o.ftz = ftz;
    o.fbx = fbx; o.fby = fby; o.fbz = fbz;
    o.tbx = tbx; o.tby = tby; o.tbz = tbz;
  }
  // semi-implicit Euler of the free ball: vb = [w (3) | v (3)], qb = [quaternion (4) | position (3)]
  template <class S>
  DDP_HD static void ball_integrate(S* qb, S* vb, const Loads<S>& L, double h, double mb, double Ib) {
    vb[0] = vb[0] + (h / Ib) * L.tbx;
    vb[1] = vb[1] + (h / Ib) * L.tby;
    vb[2] = vb[2] + (h / Ib) * L.tbz;
    vb[3] = vb[3] + (h / mb) * L.fbx;
    vb[4] = vb[4] + (h / mb) * L.fby;
    vb[5] = vb[5] + (h / mb) * L.fbz;
    // quaternion rate for world-frame angular velocity: qdot = 0.5 * (0,w) (x) q
    S qw = qb[0], qx = qb[1], qy = qb[2], qz = qb[3];
    qb[0] = qw + (0.5 * h) * (-(vb[0] * qx) - vb[1] * qy - vb[2] * qz);
    qb[1] = qx + (0.5 * h) * (vb[0] * qw + vb[1] * qz - vb[2] * qy);
    qb[2] = qy + (0.5 * h) * (vb[1] * qw + vb[2] * qx - vb[0] * qz);
    qb[3] = qz + (0.5 * h) * (vb[2] * qw + vb[0] * qy - vb[1] * qx);
    qb[4] = qb[4] + h * vb[3];
    qb[5] = qb[5] + h * vb[4];
    qb[6] = qb[6] + h * vb[5];
  }

  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const int sub = (int)p[1];
    const double h = p[0] / sub;
    const double Ij = p[2], bj = p[3];
    const double* d = p + 4;
    const double rb = p[12], mb = p[13], bz = p[18];
    const double Ib = 0.4 * mb * rb * rb;
    S q[14], v[13];
#pragma unroll
    for (int i = 0; i < 14; ++i) q[i] = x[i];
#pragma unroll
    for (int i = 0; i < 13; ++i) v[i] = x[14 + i];
    for (int it = 0; it < sub; ++it) {
      // ---- forward kinematics of the tool tip with its position Jacobian ----
      Frame<S> F;
      fk_init(q[0], bz, F);
      S ox[7], oy[7], oz[7], ax[7], ay[7], az[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        S s, c;
        sincos_(q[i], &s, &c);
        ox[i] = F.px;
        oy[i] = F.py;
        oz[i] = F.pz;
        fk_joint(i, s, c, d[i], F, ax[i], ay[i], az[i]);
      }
      // tip velocity = sum_i (a_i x (p - o_i)) qd_i ; keep the Jacobian columns
      S Jx[7], Jy[7], Jz[7];
      S tvx = 0.0 * F.px, tvy = tvx, tvz = tvx;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        S dx = F.px - ox[i], dy = F.py - oy[i], dz = F.pz - oz[i];
        Jx[i] = ay[i] * dz - az[i] * dy;
        Jy[i] = az[i] * dx - ax[i] * dz;
        Jz[i] = ax[i] * dy - ay[i] * dx;
        tvx = tvx + Jx[i] * v[i];
        tvy = tvy + Jy[i] * v[i];
        tvz = tvz + Jz[i] * v[i];
      }
      // ---- contacts: tip sphere vs ball, ball vs table ----
      Loads<S> L;
      contacts(F.px, F.py, F.pz, tvx, tvy, tvz, q[11], q[12], q[13], v + 7, v + 10, p, L);
      // ---- accelerations, semi-implicit Euler ----
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        S tau = u[i] + Jx[i] * L.ftx + Jy[i] * L.fty + Jz[i] * L.ftz - bj * v[i];
        v[i] = v[i] + (h / Ij) * tau;
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) q[i] = q[i] + h * v[i];
      ball_integrate(q + 7, v + 7, L, h, mb, Ib);
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) xn[i] = q[i];
#pragma unroll
    for (int i = 0; i < 13; ++i) xn[14 + i] = v[i];
  }
};

// ------------------------------------------------------------------------------
// Test stub used by the survey probe: x+ = A x + B u + 0.01 sin(x), any (n, m).
// p = [dt (unused), A row-major (n*n), B row-major (n*m)]
template <int N_, int M_>
struct AffineSin {
  static constexpr int n = N_, m = M_, np = 1 + N_ * N_ + N_ * M_;
  static constexpr int COOP = 1;
  template <class S>
  DDP_HD static void step(const S* x, const S* u, S* xn, const double* p) {
    const double* A = p + 1;
    const double* B = p + 1 + N_ * N_;
    for (int i = 0; i < N_; ++i) {
      S s, c;
      sincos_(x[i], &s, &c);
      (void)c;
      S acc = 0.01 * s;
      for (int j = 0; j < N_; ++j) acc = acc + A[i * N_ + j] * x[j];
      for (int j = 0; j < M_; ++j) acc = acc + B[i * M_ + j] * u[j];
      xn[i] = acc;
    }
  }
};

// Dispatch a functor templated on the model type.
#define DDP_MODEL_SWITCH(id, CALL)                                   \
  switch (id) {                                                      \
    case ::ddp::MODEL_PENDULUM: { typedef ::ddp::Pendulum Model; CALL; } break;           \
    case ::ddp::MODEL_ACROBOT: { typedef ::ddp::Acrobot Model; CALL; } break;             \
    case ::ddp::MODEL_CARTPOLE: { typedef ::ddp::CartPole Model; CALL; } break;           \
    case ::ddp::MODEL_CARTPOLE_WALL: { typedef ::ddp::CartPoleWall Model; CALL; } break;  \
    case ::ddp::MODEL_QUADRUPED: { typedef ::ddp::Quadruped Model; CALL; } break;         \
    case ::ddp::MODEL_ARM_BALL: { typedef ::ddp::ArmBall Model; CALL; } break;            \
    case ::ddp::MODEL_QUADRUPED_QUAT: { typedef ::ddp::QuadrupedQuat Model; CALL; } break; \
    case ::ddp::MODEL_AFFINE_SIN_4_1: { typedef ::ddp::AffineSin<4, 1> Model; CALL; } break;     \
    case ::ddp::MODEL_AFFINE_SIN_6_2: { typedef ::ddp::AffineSin<6, 2> Model; CALL; } break;     \
    case ::ddp::MODEL_AFFINE_SIN_27_7: { typedef ::ddp::AffineSin<27, 7> Model; CALL; } break;   \
    case ::ddp::MODEL_AFFINE_SIN_36_12: { typedef ::ddp::AffineSin<36, 12> Model; CALL; } break; \
    case ::ddp::MODEL_AFFINE_SIN_37_12: { typedef ::ddp::AffineSin<37, 12> Model; CALL; } break; \
    default: return -1;                                              \
  }

}  // namespace ddp
