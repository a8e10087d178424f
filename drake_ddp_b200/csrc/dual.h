// Forward-mode dual numbers used to linearize the discrete dynamics map.
//
// The reference obtains fx, fu by pushing n+m AutoDiffXd seeds through Drake's
// discrete update (/root/reference/ilqr.py:253-270).  Here the same thing is
// done in-kernel: the model's step() is a template on the scalar type, and the
// Jacobian pass instantiates it with Dual<K>.  On the device the n+m seed
// directions are spread over the G lanes of a lane-group (each lane carries K
// of them; the value part is recomputed redundantly by every lane, which is
// free in SIMT).  On the host the same template runs with K = n+m in one
// thread, which is what the CPU oracle's dynamics shim calls.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DDP_HD __host__ __device__ __forceinline__
#else
#define DDP_HD inline
#endif

namespace ddp {

template <int K>
struct Dual {
  double v;
  double d[K];
  DDP_HD Dual() {}
  DDP_HD Dual(double c) : v(c) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = 0.0;
  }
};

// ---- value extraction (for branches on contact state etc.) -----------------
DDP_HD double val(double a) { return a; }
template <int K>
DDP_HD double val(const Dual<K>& a) { return a.v; }

// ---- Dual (op) Dual ----------------------------------------------------------
template <int K>
DDP_HD Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <int K>
DDP_HD Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <int K>
DDP_HD Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
template <int K>
DDP_HD Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r;
  const double ib = 1.0 / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
template <int K>
DDP_HD Dual<K> operator-(const Dual<K>& a) {
  Dual<K> r;
  r.v = -a.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
  return r;
}

// ---- Dual (op) double ---------------------------------------------------------
template <int K>
DDP_HD Dual<K> operator+(const Dual<K>& a, double b) {
  Dual<K> r = a;
  r.v += b;
  return r;
}
template <int K>
DDP_HD Dual<K> operator+(double b, const Dual<K>& a) { return a + b; }
template <int K>
DDP_HD Dual<K> operator-(const Dual<K>& a, double b) {
  Dual<K> r = a;
  r.v -= b;
  return r;
}
template <int K>
DDP_HD Dual<K> operator-(double b, const Dual<K>& a) {
  Dual<K> r;
  r.v = b - a.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
  return r;
}
template <int K>
DDP_HD Dual<K> operator*(const Dual<K>& a, double b) {
  Dual<K> r;
  r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <int K>
DDP_HD Dual<K> operator*(double b, const Dual<K>& a) { return a * b; }
template <int K>
DDP_HD Dual<K> operator/(const Dual<K>& a, double b) { return a * (1.0 / b); }
template <int K>
DDP_HD Dual<K> operator/(double a, const Dual<K>& b) {
  Dual<K> r;
  const double ib = 1.0 / b.v;
  r.v = a * ib;
  const double s = -r.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * b.d[k];
  return r;
}
template <int K>
DDP_HD Dual<K>& operator+=(Dual<K>& a, const Dual<K>& b) { a = a + b; return a; }
template <int K>
DDP_HD Dual<K>& operator-=(Dual<K>& a, const Dual<K>& b) { a = a - b; return a; }
template <int K>
DDP_HD Dual<K>& operator+=(Dual<K>& a, double b) { a.v += b; return a; }

// ---- elementary functions -----------------------------------------------------
DDP_HD void sincos_(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  ::sincos(a, s, c);
#else
  *s = ::sin(a);
  *c = ::cos(a);
#endif
}
template <int K>
DDP_HD void sincos_(const Dual<K>& a, Dual<K>* s, Dual<K>* c) {
  double sv, cv;
  sincos_(a.v, &sv, &cv);
  s->v = sv;
  c->v = cv;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    s->d[k] = cv * a.d[k];
    c->d[k] = -sv * a.d[k];
  }
}
DDP_HD double sqrt_(double a) { return ::sqrt(a); }
template <int K>
DDP_HD Dual<K> sqrt_(const Dual<K>& a) {
  Dual<K> r;
  r.v = ::sqrt(a.v);
  const double h = 0.5 / r.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = h * a.d[k];
  return r;
}
// 1 / sqrt(a).  Plain double on the device: ONE rsqrt (max error 1 ulp) instead of a square root
// followed by a division, both of which sit on the serial path of a contact step (the same class of
// host / device difference as sincos: the host keeps 1 / sqrt).  The dual version takes its value
// from the same function and forms the derivative of 1.0 / sqrt_(a) with one division less: sqrt_
// divides 0.5 by the root and the quotient divides 1 by it again, and 0.5 / r == 0.5 * (1 / r)
// exactly (a power of two commutes with rounding), so on the host it is 1.0 / sqrt_(a) bit for bit.
DDP_HD double inv_sqrt_(double a) {
#if defined(__CUDA_ARCH__)
  return ::rsqrt(a);
#else
  return 1.0 / ::sqrt(a);
#endif
}
template <int K>
DDP_HD Dual<K> inv_sqrt_(const Dual<K>& a) {
  Dual<K> r;
  const double ib = inv_sqrt_(a.v);   // device: rsqrt, like the plain-double path
  const double h = 0.5 * ib;          // sqrt_: d sqrt = (0.5 / sqrt) da
  const double s = -ib * ib;          // 1 / b: d = -(1 / b^2) db
  r.v = ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * (h * a.d[k]);
  return r;
}

// r = sqrt(q) when 1 / r is wanted too (a unit normal, a friction direction), and x / r through
// it.  Plain double on the device: one rsqrt, r = q * (1 / r), x / r = x * (1 / r), instead of a
// square root and a division per use on the serial path of a contact step (plain double and dual
// alike).  The host keeps sqrt and x / r: the arithmetic of the oracle and of the golden vectors is
// what it was.
DDP_HD double sqrt_pair_(double q, double* inv) {
#if defined(__CUDA_ARCH__)
  const double i = ::rsqrt(q);
  *inv = i;
  return q * i;
#else
  *inv = 0.0;
  return ::sqrt(q);
#endif
}
DDP_HD double div_root_(double x, double root, double inv) {
#if defined(__CUDA_ARCH__)
  (void)root;
  return x * inv;
#else
  (void)inv;
  return x / root;
#endif
}
template <int K>
DDP_HD Dual<K> sqrt_pair_(const Dual<K>& q, Dual<K>* inv) {
#if defined(__CUDA_ARCH__)
  *inv = inv_sqrt_(q);
  return q * (*inv);
#else
  *inv = q;   // not used on the host
  return sqrt_(q);
#endif
}
template <int K>
DDP_HD Dual<K> div_root_(const Dual<K>& x, const Dual<K>& root, const Dual<K>& inv) {
#if defined(__CUDA_ARCH__)
  (void)root;
  return x * inv;
#else
  (void)inv;
  return x / root;
#endif
}

}  // namespace ddp
