// K6 (large-n path): backward Riccati sweep on the fp64 tensor pipe.
//
// Replaces _backward_pass (/root/reference/ilqr.py:623-667) for models with n >= 16.  One CTA
// per trajectory, sequential over t = N-2 .. 0.  Per step the dense contractions run as
// mma.sync.m8n8k4 f64 (DMMA) tiles out of shared memory; Vxx / Vx never leave the SM.  With
// S = [fx | fu] (n x (n+m), the two tiles stacked so that 8-wide strips are not padded twice):
//   A1  [W(:, tail) | Wu] = Vxx S(:, strips that contain fu)   first, so that ...
//   A2  Quu = luu + fu' Wu                                   ... Quu exists early: the vector warp
//                                                            inverts it (ilqr.py:655) under A3 and B
//   A3  W(:, head) = Vxx fx(:, the other strips)
//   B   [Qxx Qx ; Qux Qu] = [lxx lx ; 0 lu] + S' [W | Vx]     Vx rides along as column n of W, so
//                                                            Qx and Qu cost no extra DMMA; Qxx overwrites Vxx
//   C1  K = Quu^-1 Qux      C2  Vxx <- Qxx - Qux' K           vector warp: kappa, dV, Vx
// The fx / x_bar / u_bar tiles of step t-1 are prefetched into the other half of a double
// buffer by 1-D bulk TMA (cp.async.bulk + mbarrier) while step t computes, fu is refilled as
// soon as it is dead; when the tile sizes are not 16-byte multiples (odd n) the same schedule
// runs on 8-byte cp.async copies issued by all threads.
// Warp roles (warp id % 4 selects the SM sub-partition, and DMMA and DFMA share the fp64 pipe of
// their sub-partition): a CTA is 4 warps, three DMMA warps that each own a third of the row
// tiles (or column strips) of every product, so an operand fragment fetched from shared memory
// feeds 2-6 DMMAs, and one vector warp for the serial fp64 work (cost gradients, the inverse of
// Quu -- Newton-Schulz on the tensor pipe seeded with the previous step's inverse, Gauss-Jordan
// as fallback --, kappa, dV, Vx).  The vector role goes to warp `slot`, the CTA's residency slot
// on its SM (per-SM bitmask in global memory), so the four CTAs of an SM put their vector warps
// on four different sub-partitions and every sub-partition hosts three DMMA warps.
// DESIGN.md section 3 has the measured phase times and what was tried without gain.
#pragma once
#include "kernels.cuh"

namespace ddp {

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
// element (r, c) of a row-major matrix with leading dimension ld, zero outside R x C
__device__ __forceinline__ double ldz(const double* P, int ld, int r, int c, int R, int C) {
  return (r < R && c < C) ? P[r * ld + c] : 0.0;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// predicated DMMA: p is warp-uniform (mma.sync needs the whole warp or none of it)
__device__ __forceinline__ void dmma_p(double (&c)[2], double a, double b, int p) {
  asm volatile(
      "{\n"
      ".reg .pred pp;\n"
      "setp.ne.s32 pp, %4, 0;\n"
      "@pp mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      "}\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b), "r"(p));
}

template <int n, int m>
struct BwdMmaCfg {
  static constexpr int TN = (n + 7) / 8, TM = (m + 7) / 8, KN = (n + 3) / 4, KM = (m + 3) / 4;
  static constexpr int NS = n + m, TS = (NS + 7) / 8;  // stacked [fx | fu] columns / 8-wide tiles
  static constexpr int TF = n / 8;                     // stacked strips [0, TF) hold fx columns only
  static constexpr int NMW = 3, NW = NMW + 1, NT = NW * 32;  // 3 DMMA warps + the vector warp
  static_assert(NW == 4, "the role rotation assumes one warp per SM sub-partition");
  // leading dimension of W: room for Vx as column n, and conflict-free 4 x 8 operand fetches
  static constexpr int ldw() {
    int v = n + 1;
    while (!(v % 8 == 4 || v % 16 == 8)) ++v;
    return v;
  }
  static constexpr int LDW = ldw(), TW = (n + 1 + 7) / 8;
  // bulk TMA needs 16-byte sizes and 16-byte aligned tile starts in global memory
  static constexpr bool TMA = ((n * n) % 2 == 0) && ((n * m) % 2 == 0) && (n % 2 == 0) && (m % 2 == 0);
  static constexpr bool EVEN = (n % 2 == 0) && (m % 2 == 0);  // C-fragment pairs never straddle rows
  static constexpr int even(int v) { return (v + 1) & ~1; }
#ifndef DDP_BWD_MINB
#define DDP_BWD_MINB 4
#endif
  static constexpr int MINB = (n <= 36) ? DDP_BWD_MINB : 3;  // CTAs per SM the register budget is sized for
  static_assert(2 * m <= n, "the Newton-Schulz scratch (2 m^2) lives in the Wu buffer (n m)");
  static_assert(n <= 64 && m <= 32, "vector warp keeps lx in two registers per lane and lu in one");
};

template <int n, int m>
struct BwdMmaSmem {
  typedef BwdMmaCfg<n, m> C;
  alignas(16) double Fx[2][C::even(n * n)];   // double-buffered bulk-TMA destination
  alignas(16) double Fu[C::even(n * m)];      // single buffer: refilled after phase B (dead by then)
  alignas(16) double xb[2][C::even(n)];
  alignas(16) double ub[2][C::even(m)];
  alignas(16) double Vxx[C::even(n * n)];
  alignas(16) double W[n * C::LDW];           // W = Vxx fx in columns 0..n-1, Vx in column n
  alignas(16) double WuKt[C::even(n * m)];    // Wu = Vxx fu (phase A), then K_t (phase C)
  alignas(16) double Qux[C::even(m * n)];
  alignas(16) double Quu[C::even(m * m)];
  alignas(16) double QuuInv[C::even(m * m)];  // Quu^-1; kept across steps: it seeds the next inversion
  double QxM[n], QuM[m];                      // fx' Vx, fu' Vx (column n of product B)
  double Qu[m], g[m], Qd2[n], Rd2[m];         // Qd2, Rd2 = diagonals of lxx = 2 Q, luu = 2 R
  alignas(8) uint64_t bar[2];
  alignas(8) uint64_t barFu;
  int slot;                                   // CTA slot on this SM (deals the warp roles)
};

// DMMA operand slots of one warp: slot i covers 8 rows (A) or 8 columns (B) of the operand; p[i]
// is this lane's element at k = tg, consecutive k-steps are 4*KS (OpS, compile-time stride) or
// s4[i] (OpD, per-lane stride: the stacked [fx | fu] operand changes matrix inside a tile) apart.
template <int MT, int KS>
struct OpS {
  const double* p[MT];
  __device__ __forceinline__ double ld(int i, int kk) const { return p[i][kk * 4 * KS]; }
};
template <int MT>
struct OpD {
  const double* p[MT];
  int s4[MT];
  __device__ __forceinline__ double ld(int i, int kk) const { return p[i][kk * s4[i]]; }
};

// group w (0..2) of X items dealt to the three DMMA warps: sizes differ by at most one
template <int X>
__device__ __forceinline__ void split3(int w, int& start, int& cnt) {
  constexpr int base = X / 3, rem = X % 3;
  cnt = base + (w < rem ? 1 : 0);
  start = w * base + min(w, rem);
}

// acc[i][j] += A_i B_j over KT k-steps for an MR x MC block of 8 x 8 tiles; slots i >= nr or
// j >= nc are skipped (RP / CP say whether that can happen at all).  KDIM is the true inner
// dimension: lanes whose k falls into the padding of the last k-step feed zeros.  Each operand
// fragment loaded from shared memory feeds MC (A) or MR (B) DMMAs.
template <int MR, int MC, int KT, int KDIM, bool RP, bool CP, class OA, class OB>
__device__ __forceinline__ void mma_rect(double (&acc)[MR][MC][2], const OA& A, const OB& Bo, int nr, int nc,
                                         int tg) {
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    const bool kin = (4 * kk + 3 < KDIM) || (4 * kk + tg < KDIM);
    double a[MR], b[MC];
#pragma unroll
    for (int i = 0; i < MR; ++i) a[i] = kin ? A.ld(i, kk) : 0.0;
#pragma unroll
    for (int j = 0; j < MC; ++j) b[j] = kin ? Bo.ld(j, kk) : 0.0;
#pragma unroll
    for (int i = 0; i < MR; ++i)
#pragma unroll
      for (int j = 0; j < MC; ++j) {
        if (RP && CP) dmma_p(acc[i][j], a[i], b[j], (i < nr) & (j < nc));
        else if (RP) dmma_p(acc[i][j], a[i], b[j], i < nr);
        else if (CP) dmma_p(acc[i][j], a[i], b[j], j < nc);
        else dmma(acc[i][j], a[i], b[j]);
      }
  }
}

// One product block: rows [r0, r0 + nr) x columns [c0, c0 + nc) in tiles, nr <= MR, nc <= MC.
// mkA(op, slot, r) / mkB(op, slot, c) fill operand slots for row r / column c (lane offsets
// included); store(r, c, v0, v1) receives the C-fragment pairs (row r, columns c and c + 1).
template <int MR, int MC, int KT, int KDIM, bool RP, bool CP, class OA, class OB, class MKA, class MKB, class FS>
__device__ __forceinline__ void mma_product(int r0, int nr, int c0, int nc, const int g, const int tg, MKA mkA,
                                            MKB mkB, FS store) {
  OA A;
  OB Bo;
#pragma unroll
  for (int i = 0; i < MR; ++i) mkA(A, i, 8 * (r0 + ((i < nr) ? i : 0)) + g);
#pragma unroll
  for (int j = 0; j < MC; ++j) mkB(Bo, j, 8 * (c0 + ((j < nc) ? j : 0)) + g);
  double acc[MR][MC][2];
#pragma unroll
  for (int i = 0; i < MR; ++i)
#pragma unroll
    for (int j = 0; j < MC; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  mma_rect<MR, MC, KT, KDIM, RP, CP>(acc, A, Bo, nr, nc, tg);
#pragma unroll
  for (int i = 0; i < MR; ++i)
#pragma unroll
    for (int j = 0; j < MC; ++j)
      if (i < nr && j < nc) store(8 * (r0 + i) + g, 8 * (c0 + j) + 2 * tg, acc[i][j][0], acc[i][j][1]);
}

// single 8 x 8 tile with three interleaved accumulation chains (k-steps 0,3,6.. / 1,4,7.. /
// 2,5,8..): a third of the dependent-DMMA latency of one chain
template <int KT, int KDIM, class FA, class FB>
__device__ __forceinline__ void mma_tile3(double (&acc)[2], int tg, FA A, FB Bm) {
  double c[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    const bool kin = (4 * kk + 3 < KDIM) || (4 * kk + tg < KDIM);
    dmma(c[kk % 3], kin ? A(kk) : 0.0, kin ? Bm(kk) : 0.0);
  }
  acc[0] = c[0][0] + (c[1][0] + c[2][0]);
  acc[1] = c[0][1] + (c[1][1] + c[2][1]);
}

// store the pair (v0, v1) at P[idx], P[idx + 1]; one 16-byte store when pairs are aligned
template <bool VEC>
__device__ __forceinline__ void st_pair(double* P, int idx, bool ok0, bool ok1, double v0, double v1) {
  if (VEC) {
    if (ok0) *reinterpret_cast<double2*>(P + idx) = make_double2(v0, v1);
  } else {
    if (ok0) P[idx] = v0;
    if (ok1) P[idx + 1] = v1;
  }
}

#ifdef DDP_BWD_PROFILE
__device__ int g_bwd_fallbacks;  // Gauss-Jordan inversions (first step of every trajectory included)
__device__ int g_bwd_passes;     // Newton-Schulz passes, all inversions
#endif

// Inverse of the m x m matrix A by Newton-Schulz iteration on the fp64 tensor pipe, one warp:
//   R = I - A X ;  X <- X + X R      (error squares every pass)
// started from the inverse of the previous backward step, which is still in X: Quu moves little
// between neighbouring timesteps, so three or four passes (24 DMMAs each at m = 12) reach full
// precision, against roughly a thousand dependent-latency-bound instructions of Gauss-Jordan.
// The inverse sits on the serial path of a backward step, and on this machine every dependent
// DFMA / shuffle / vote costs 25-75 cycles even uncontended, so that difference is the step time.
// Returns false when the start is not contracting (max |R| >= 1/2, or NaN) or the iteration
// does not reach max |R| < 2^-24 within 8 passes (after which one more pass squares it below an
// ulp): the caller then runs the Gauss-Jordan inverse with partial pivoting (invert_warp).
// Xs, Rs: m*m doubles of scratch each.  Stands in for np.linalg.inv(Quu) (ilqr.py:655).
template <int m>
__device__ __forceinline__ bool invert_newton_warp(const double* A, double* X, double* Xs, double* Rs) {
  constexpr int TM = (m + 7) / 8, KM = (m + 3) / 4;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const unsigned full = 0xffffffffu;
  double xc[TM][TM][2];
#pragma unroll
  for (int mt = 0; mt < TM; ++mt)
#pragma unroll
    for (int nt = 0; nt < TM; ++nt) {
      const int r = 8 * mt + g, c = 8 * nt + 2 * tg;
      xc[mt][nt][0] = ldz(X, m, r, c, m, m);
      xc[mt][nt][1] = ldz(X, m, r, c + 1, m, m);
    }
  for (int i = lane; i < m * m; i += 32) Xs[i] = X[i];
  __syncwarp();
  bool ok = false;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    double y[TM][TM][2];
#pragma unroll
    for (int mt = 0; mt < TM; ++mt)
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) y[mt][nt][0] = y[mt][nt][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < KM; ++kk) {
      double a[TM], bb[TM];
#pragma unroll
      for (int mt = 0; mt < TM; ++mt) a[mt] = ldz(A, m, 8 * mt + g, 4 * kk + tg, m, m);
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) bb[nt] = ldz(Xs, m, 4 * kk + tg, 8 * nt + g, m, m);
#pragma unroll
      for (int mt = 0; mt < TM; ++mt)
#pragma unroll
        for (int nt = 0; nt < TM; ++nt) dmma(y[mt][nt], a[mt], bb[nt]);
    }
    unsigned hmax = 0;
#pragma unroll
    for (int mt = 0; mt < TM; ++mt)
#pragma unroll
      for (int nt = 0; nt < TM; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = 8 * mt + g, c = 8 * nt + 2 * tg + e;
          double rv = ((r == c) ? 1.0 : 0.0) - y[mt][nt][e];
          if (r >= m || c >= m) rv = 0.0;
          y[mt][nt][e] = rv;
          hmax = max(hmax, (unsigned)(__double_as_longlong(fabs(rv)) >> 32));
        }
    hmax = __reduce_max_sync(full, hmax);
    if (hmax >= 0x3FE00000u) break;  // |r| >= 1/2, inf or NaN: not contracting
#pragma unroll
    for (int mt = 0; mt < TM; ++mt)
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) {
        const int r = 8 * mt + g, c = 8 * nt + 2 * tg;
        if (r < m && c < m) Rs[r * m + c] = y[mt][nt][0];
        if (r < m && c + 1 < m) Rs[r * m + c + 1] = y[mt][nt][1];
      }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < KM; ++kk) {
      double a[TM], bb[TM];
#pragma unroll
      for (int mt = 0; mt < TM; ++mt) a[mt] = ldz(Xs, m, 8 * mt + g, 4 * kk + tg, m, m);
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) bb[nt] = ldz(Rs, m, 4 * kk + tg, 8 * nt + g, m, m);
#pragma unroll
      for (int mt = 0; mt < TM; ++mt)
#pragma unroll
        for (int nt = 0; nt < TM; ++nt) dmma(xc[mt][nt], a[mt], bb[nt]);
    }
    __syncwarp();  // every read of Xs done
#pragma unroll
    for (int mt = 0; mt < TM; ++mt)
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) {
        const int r = 8 * mt + g, c = 8 * nt + 2 * tg;
        if (r < m && c < m) Xs[r * m + c] = xc[mt][nt][0];
        if (r < m && c + 1 < m) Xs[r * m + c + 1] = xc[mt][nt][1];
      }
    __syncwarp();
#ifdef DDP_BWD_PROFILE
    if (lane == 0) atomicAdd(&g_bwd_passes, 1);
#endif
    if (hmax < 0x3E700000u) {  // max |R| < 2^-24 before this pass: error now below an ulp
      ok = true;
      break;
    }
  }
  if (ok) {
    for (int i = lane; i < m * m; i += 32) X[i] = Xs[i];
    __syncwarp();
  }
  return ok;
}

// four CTAs per SM at the headline shape: 4 x (this + 1 KB reserved) must fit in 227 KB
static_assert(sizeof(BwdMmaSmem<36, 12>) <= 57088, "backward_mma_kernel: shared memory budget for 4 CTAs/SM");

#if !defined(DDP_BWD_PROFILE) && !defined(DDP_BWD_FENCE)
#define DDP_BWD_FENCE 1
#endif
#ifdef DDP_BWD_PROFILE
// per-phase cycle totals of one DMMA warp and the vector warp of two CTAs (first / second wave)
__device__ long long g_bwd_prof[2][4][16];   // [cta][role][phase]

// the clock is read after a shared-memory load so that it cannot run ahead of a barrier the
// warp has arrived at but not yet passed (BAR.SYNC.DEFER_BLOCKING)
#define BWD_TICK(i)                                                     \
  do {                                                                  \
    if (prof_on) {                                                      \
      const int dep_ = *reinterpret_cast<volatile int*>(&s.slot);       \
      long long now_;                                                   \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_) : "r"(dep_) : "memory"); \
      prof_acc[i] += now_ - prof_last;                                  \
      prof_last = now_;                                                 \
    }                                                                   \
  } while (0)
#elif DDP_BWD_FENCE
// Phase fences: a never-taken branch at every phase boundary ends the basic block there, so
// ptxas schedules every phase on its own.  Measured at C4 on B200: 4.02 ms per sweep without,
// 3.65 ms with (-DDDP_BWD_FENCE=0 builds the unfenced kernel).
#define BWD_TICK(i)                                                     \
  do {                                                                  \
    if (fence_never) d.dV[i] = (double)clock64();                       \
  } while (0)
#else
#define BWD_TICK(i)
#endif

template <class Model>
__global__ void __launch_bounds__(BwdMmaCfg<Model::n, Model::m>::NT, BwdMmaCfg<Model::n, Model::m>::MINB)
backward_mma_kernel(Dev d) {
  constexpr int n = Model::n, m = Model::m;
  typedef BwdMmaCfg<n, m> C;
  constexpr int TN = C::TN, TM = C::TM, KN = C::KN, KM = C::KM, NT = C::NT, TS = C::TS, TF = C::TF, TW = C::TW;
  constexpr int NMW = C::NMW, LDW = C::LDW;
  constexpr bool EVEN = C::EVEN;
  // largest group when X tiles are dealt to the three DMMA warps
  constexpr int GS = (TS + 2) / 3, GN = (TN + 2) / 3;
  // K = Quu^-1 Qux: strips dealt (C1A, C1A, rest) so that the warp with the smallest share of the
  // value update (the last group of split3) takes most of it
  constexpr int C1A = (TN >= 5) ? TN / 5 : ((TN >= 3) ? 1 : 0), C1R = TN - 2 * C1A;
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdMmaSmem<n, m>& s = *reinterpret_cast<BwdMmaSmem<n, m>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int N = d.N, T = d.T;
  const double* Q = d.Q;
  const double* R = d.R;
  const double* Qf = d.Qf;
  const bool diag = d.diag_cost != 0;
  const double* xnom = d.x_nom + (size_t)b * n;
  const double* gfx = d.fx + (size_t)b * T * n * n;
  const double* gfu = d.fu + (size_t)b * T * n * m;
  const double* gxb = d.x_bar + (size_t)b * N * n;
  const double* gub = d.u_bar + (size_t)b * T * m;
  constexpr uint32_t kStageBytes = (n * n + n + m) * 8;

  auto issue_tile = [&](int t, int buf) {   // one thread
    mbar_expect_tx(&s.bar[buf], kStageBytes);
    tma_load_1d(s.Fx[buf], gfx + (size_t)t * n * n, n * n * 8, &s.bar[buf]);
    tma_load_1d(s.xb[buf], gxb + (size_t)t * n, n * 8, &s.bar[buf]);
    tma_load_1d(s.ub[buf], gub + (size_t)t * m, m * 8, &s.bar[buf]);
  };
  auto issue_fu = [&](int t) {              // one thread
    mbar_expect_tx(&s.barFu, n * m * 8);
    tma_load_1d(s.Fu, gfu + (size_t)t * n * m, n * m * 8, &s.barFu);
  };
  // odd n or m: tile starts are only 8-byte aligned, so the tiles travel as 8-byte cp.async
  // copies issued by all threads (same double-buffer schedule as the bulk-TMA path)
  auto cp8 = [&](double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
  };
  auto copy_tile_async = [&](int t, int buf) {   // all threads
    for (int i = tid; i < n * n; i += NT) cp8(&s.Fx[buf][i], gfx + (size_t)t * n * n + i);
    for (int i = tid; i < n; i += NT) cp8(&s.xb[buf][i], gxb + (size_t)t * n + i);
    for (int i = tid; i < m; i += NT) cp8(&s.ub[buf][i], gub + (size_t)t * m + i);
  };
  auto copy_fu_async = [&](int t) {              // all threads
    for (int i = tid; i < n * m; i += NT) cp8(&s.Fu[i], gfu + (size_t)t * n * m + i);
  };

  // Deal the warp roles by CTA slot: the CTAs resident on one SM take distinct slots (per-SM
  // bitmask in global memory), and the vector role goes to warp `slot`, so every sub-partition
  // hosts one vector warp and three DMMA warps instead of all vector warps queueing on one.
  int* slot_word = nullptr;
  if (tid == 0) {
    int sl = 0;
    if (d.sm_slots) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      slot_word = d.sm_slots + (smid & 1023);
      for (sl = 0; sl < 4; ++sl)
        if (!(atomicOr(slot_word, 1 << sl) & (1 << sl))) break;
      if (sl == 4) {  // more than 4 resident CTAs (another solver on this GPU): share slot 0
        sl = 0;
        slot_word = nullptr;
      }
    }
    s.slot = sl;
    if (C::TMA) {
      mbar_init(&s.bar[0], 1);
      mbar_init(&s.bar[1], 1);
      mbar_init(&s.barFu, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const int slot = s.slot;
  const int role = (warp - slot - 1) & 3;   // 0..2: DMMA roles; 3 (warp == slot): the vector warp
#ifdef DDP_BWD_STAGGER
  // CTAs started together run their phases in lock-step and then all want the tensor pipe in
  // the same phases: offset them by a fraction of a step
  {
    const long long until = clock64() + (long long)slot * DDP_BWD_STAGGER;
    while (clock64() < until) {
    }
  }
#endif
  const bool issuer = (role == NMW) && (lane == 0);  // drives the TMA queue
  if (C::TMA) {
    if (issuer) {
      issue_tile(T - 1, 0);
      issue_fu(T - 1);
    }
  } else {
    copy_tile_async(T - 1, 0);
    copy_fu_async(T - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // Vx, Vxx <- terminal cost partials at x_bar[:, -1]            (ilqr.py:638, 203-204)
  {
    const double* xl = gxb + (size_t)(N - 1) * n;
    for (int i = tid; i < n * n; i += NT) s.Vxx[i] = 2.0 * Qf[i];
    for (int i = tid; i < n * LDW; i += NT) s.W[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      double a = 0.0, c = 0.0;
      for (int j = 0; j < n; ++j) {
        a = fma(2.0 * Qf[i * n + j], xl[j], a);
        c = fma(2.0 * xnom[j], Qf[j * n + i], c);
      }
      s.W[i * LDW + n] = a - c;
      s.Qd2[i] = 2.0 * Q[i * n + i];
    }
    for (int i = tid; i < m; i += NT) s.Rd2[i] = 2.0 * R[i * m + i];
  }
  // vector warp: the step-independent half of lx = 2 Q x - 2 x_nom' Q       (ilqr.py:180)
  double lxc0 = 0.0, lxc1 = 0.0;
  if (role == NMW) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = lane + 32 * h;
      double c = 0.0;
      if (k < n) {
        if (diag) c = 2.0 * xnom[k] * Q[k * n + k];
        else
          for (int j = 0; j < n; ++j) c = fma(2.0 * xnom[j], Q[j * n + k], c);
      }
      if (h == 0) lxc0 = c;
      else lxc1 = c;
    }
  }
  __syncthreads();

  uint32_t parity[2] = {0, 0};
  uint32_t parityFu = 0;
  int buf = 0;
#if DDP_BWD_FENCE
  const bool fence_never = (DDP_BWD_FENCE == 2) ? (tid == d.N + 100000) : (d.N < 0);
#endif
#ifdef DDP_BWD_PROFILE
  const int prof_cta = (b == 5) ? 0 : ((b == 100) ? 1 : -1);
  const bool prof_on = prof_cta >= 0 && lane == 0;
  long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_last = clock64();
#endif
  for (int t = T - 1; t >= 0; --t, buf ^= 1) {
    BWD_TICK(0);
    if (C::TMA) {
      if (issuer && t > 0) issue_tile(t - 1, buf ^ 1);
      mbar_wait(&s.bar[buf], parity[buf]);
      parity[buf] ^= 1;
      mbar_wait(&s.barFu, parityFu);
      parityFu ^= 1;
    } else {
      // this step's tiles (and fu, issued after the previous step's phase B) have landed ...
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      // ... and the next step's start travelling into the other half
      if (t > 0) {
        copy_tile_async(t - 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    BWD_TICK(1);
    const double* Fx = s.Fx[buf];
    const double* Fu = s.Fu;

    double lx0 = 0.0, lx1 = 0.0, lu = 0.0;   // vector warp: cost gradients of this step
    if (role < NMW) {
      // operand slot of the stacked S = [fx | fu] for column (or S' row) c: lane element S[tg][c]
      auto stacked_slot = [&](auto& o, int i, int c) {
        c = min(c, n + m - 1);
        const bool in_fx = c < n;
        o.p[i] = in_fx ? (Fx + tg * n + c) : (Fu + tg * m + (c - n));
        o.s4[i] = in_fx ? 4 * n : 4 * m;
      };
      auto vxx_rows = [&](OpS<GN, 1>& o, int i, int r) { o.p[i] = s.Vxx + min(r, n - 1) * n + tg; };
      // (the epilogues are written branch-free -- selected pointers, predicated stores: a
      // divergent branch per tile costs more than the DMMAs of the tile)
      auto store_w = [&](int r, int c, double v0, double v1) {   // columns of [W | Wu]
        const bool rok = r < n;
        if (EVEN) {  // c even, n even: the pair is inside W or inside Wu
          const bool in_w = c < n;
          double* dst = in_w ? (s.W + r * LDW + c) : (s.WuKt + r * m + (c - n));
          if (rok && c < n + m) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
          if (rok) {
            if (c < n) s.W[r * LDW + c] = v0;
            else if (c < n + m) s.WuKt[r * m + (c - n)] = v0;
            if (c + 1 < n) s.W[r * LDW + c + 1] = v1;
            else if (c + 1 < n + m) s.WuKt[r * m + (c + 1 - n)] = v1;
          }
        }
      };
      auto stacked_tail = [&](OpD<TS - TF>& o, int i, int c) { stacked_slot(o, i, c); };
      int ra0, nra;   // this warp's row tiles of the Vxx products
      split3<TN>(role, ra0, nra);
      // ---- A1: [W(:, tail) | Wu] = Vxx S(:, tail strips): the strips that contain fu first, so
      // that Quu and its inverse (the serial path) start as early as possible -------------------
      mma_product<GN, TS - TF, KN, n, (TN % 3) != 0, false, OpS<GN, 1>, OpD<TS - TF>>(
          ra0, nra, TF, TS - TF, g, tg, vxx_rows, stacked_tail, store_w);
      // Wu complete (DMMA warps only)
      asm volatile("bar.sync 1, %0;" ::"r"(NMW * 32) : "memory");
      BWD_TICK(2);
      // ---- A2: Quu = luu + fu' Wu (ilqr.py:654), handed to the vector warp (barrier 2) --------
      for (int q = role; q < TM * TM; q += NMW) {
        const int r8 = 8 * (q / TM), c8 = 8 * (q % TM);
        const double* pa = Fu + tg * m + min(r8 + g, m - 1);
        const double* pb = s.WuKt + tg * m + min(c8 + g, m - 1);
        double acc[2];
        mma_tile3<KN, n>(
            acc, tg, [&](int kk) { return pa[kk * 4 * m]; }, [&](int kk) { return pb[kk * 4 * m]; });
        const int r = r8 + g, c = c8 + 2 * tg;
        if (r < m) {
          if (c < m) s.Quu[r * m + c] = 2.0 * R[r * m + c] + acc[0] + ((r == c) ? d.quu_reg : 0.0);
          if (c + 1 < m) s.Quu[r * m + c + 1] = 2.0 * R[r * m + c + 1] + acc[1] + ((r == c + 1) ? d.quu_reg : 0.0);
        }
      }
      asm volatile("bar.arrive 2, %0;" ::"r"(NT) : "memory");
      BWD_TICK(3);
      // ---- A3: W(:, head) = Vxx fx(:, head strips) -------------------------------------------
      mma_product<GN, TF, KN, n, (TN % 3) != 0, false, OpS<GN, 1>, OpS<TF, n>>(
          ra0, nra, 0, TF, g, tg, vxx_rows, [&](OpS<TF, n>& o, int j, int c) { o.p[j] = Fx + tg * n + c; }, store_w);
      // W complete and every read of the old Vxx done (DMMA warps only)
      asm volatile("bar.sync 1, %0;" ::"r"(NMW * 32) : "memory");
      BWD_TICK(4);
      // ---- B: [Qxx Qx ; Qux Qu] = [lxx . ; 0 .] + S' [W | Vx]          (ilqr.py:651-653,656) -----
      // warp w owns its group of S' row tiles, all column strips of [W | Vx].  (Claiming single
      // row tiles from a ticket counter instead, to even out the different progress of the three
      // sub-partitions under contention, was measured slower: 3.79 vs 3.66 ms -- less operand reuse.)
      {
        int r0, nr;
        split3<TS>(role, r0, nr);
        mma_product<GS, TW, KN, n, (TS % 3) != 0, false, OpD<GS>, OpS<TW, LDW>>(
            r0, nr, 0, TW, g, tg, stacked_slot,
            [&](OpS<TW, LDW>& o, int j, int c) { o.p[j] = s.W + tg * LDW + min(c, LDW - 1); },
            [&](int r, int c, double v0, double v1) {
              const bool rok = ((n + m) % 8 == 0) || (r < n + m);
              const bool isx = r < n;                  // Qxx row, else Qux row r - n
              // column n: fx' Vx, fu' Vx
              double* vec = isx ? (s.QxM + r) : (s.QuM + (r - n));
              if (rok && c == n) *vec = v0;
              if (!EVEN && rok && c + 1 == n) *vec = v1;
              if (diag) {
                const double qd = s.Qd2[isx ? r : 0];
                v0 += (isx && r == c) ? qd : 0.0;
                v1 += (isx && r == c + 1) ? qd : 0.0;
              } else if (isx) {
                if (c < n) v0 += 2.0 * Q[r * n + c];
                if (c + 1 < n) v1 += 2.0 * Q[r * n + c + 1];
              }
              double* dst = isx ? (s.Vxx + r * n) : (s.Qux + (r - n) * n);
              st_pair<EVEN>(dst, c, rok && c < n, rok && c + 1 < n, v0, v1);
            });
      }
      BWD_TICK(5);
    } else {
      // lx = 2 Q x - 2 x_nom' Q ; lu = 2 R u                       (ilqr.py:180-181)
      const double* xb = s.xb[buf];
      const double* ub = s.ub[buf];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        double a = 0.0;
        if (k < n) {
          if (diag) a = s.Qd2[k] * xb[k];
          else
            for (int j = 0; j < n; ++j) a = fma(2.0 * Q[k * n + j], xb[j], a);
        }
        if (h == 0) lx0 = a - lxc0;
        else lx1 = a - lxc1;
      }
      if (lane < m) {
        if (diag) lu = s.Rd2[lane] * ub[lane];
        else
          for (int j = 0; j < m; ++j) lu = fma(2.0 * R[lane * m + j], ub[j], lu);
      }
      BWD_TICK(2);
      asm volatile("bar.sync 2, %0;" ::"r"(NT) : "memory");
      BWD_TICK(3);
      // Quu^-1 (ilqr.py:655): Newton-Schulz from the previous step's inverse; Wu is dead by
      // now, so its buffer is the scratch
      if (t == T - 1 || (d.bwd_flags & 1) || !invert_newton_warp<m>(s.Quu, s.QuuInv, s.WuKt, s.WuKt + m * m)) {
        invert_warp<m>(s.Quu, s.QuuInv);
#ifdef DDP_BWD_PROFILE
        if (lane == 0) atomicAdd(&g_bwd_fallbacks, 1);
#endif
      }
      BWD_TICK(4);
    }
    __syncthreads();
    BWD_TICK(6);

    // fu of this step is dead now: refill the single Fu buffer with the next step's tile
    if (C::TMA) {
      if (issuer && t > 0) issue_fu(t - 1);
    } else if (t > 0) {
      copy_fu_async(t - 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---------------- phase C: K = Quu^-1 Qux ; Vxx = Qxx - Qux' K ; vector warp: kappa, dV, Vx --
    if (role < NMW) {
      double* gK = d.K + ((size_t)b * T + t) * m * n;
      {
        const int c0 = (role < 2) ? role * C1A : 2 * C1A, nc = (role < 2) ? C1A : C1R;
        mma_product<TM, C1R, KM, m, false, true, OpS<TM, 1>, OpS<C1R, n>>(
            0, TM, c0, nc, g, tg,
            [&](OpS<TM, 1>& o, int i, int r) { o.p[i] = s.QuuInv + min(r, m - 1) * m + tg; },
            [&](OpS<C1R, n>& o, int j, int c) { o.p[j] = s.Qux + tg * n + min(c, n - 1); },
            [&](int r, int c, double v0, double v1) {
              const bool rok = r < m;
              st_pair<EVEN>(s.WuKt, r * n + c, rok && c < n, rok && c + 1 < n, v0, v1);
              st_pair<EVEN>(gK, r * n + c, rok && c < n, rok && c + 1 < n, v0, v1);
            });
      }
      BWD_TICK(7);
      // every strip of K_t is needed by every warp's tiles of the update
      asm volatile("bar.sync 1, %0;" ::"r"(NMW * 32) : "memory");
      BWD_TICK(8);
      {
        int c0, nc;
        split3<TN>(role, c0, nc);
        mma_product<TN, GN, KM, m, false, (TN % 3) != 0, OpS<TN, n>, OpS<GN, n>>(
            0, TN, c0, nc, g, tg,
            [&](OpS<TN, n>& o, int i, int r) { o.p[i] = s.Qux + tg * n + min(r, n - 1); },
            [&](OpS<GN, n>& o, int j, int c) { o.p[j] = s.WuKt + tg * n + min(c, n - 1); },
            [&](int r, int c, double v0, double v1) {
              if (EVEN) {
                const bool ok = r < n && c < n;
                double2* p = reinterpret_cast<double2*>(&s.Vxx[ok ? (r * n + c) : 0]);
                double2 o = *p;
                o.x -= v0;
                o.y -= v1;
                if (ok) *p = o;
              } else if (r < n) {
                if (c < n) s.Vxx[r * n + c] -= v0;
                if (c + 1 < n) s.Vxx[r * n + c + 1] -= v1;
              }
            });
      }
    } else {
      // Qu = lu + fu' Vx ; kappa = Quu^-1 Qu ; g = Qu' Quu^-1 ; dV = g Qu ; Vx = Qx - g Qux
      //                                                        (ilqr.py:652,659,663,666)
      double qu = 0.0;
      if (lane < m) {
        qu = lu + s.QuM[lane];
        s.Qu[lane] = qu;
      }
      __syncwarp();
      double gr = 0.0;
      if (lane < m) {
        double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
        int j = 0;
        for (; j + 1 < m; j += 2) {
          a0 = fma(s.QuuInv[lane * m + j], s.Qu[j], a0);
          a1 = fma(s.QuuInv[lane * m + j + 1], s.Qu[j + 1], a1);
          c0 = fma(s.Qu[j], s.QuuInv[j * m + lane], c0);
          c1 = fma(s.Qu[j + 1], s.QuuInv[(j + 1) * m + lane], c1);
        }
        if (j < m) {
          a0 = fma(s.QuuInv[lane * m + j], s.Qu[j], a0);
          c0 = fma(s.Qu[j], s.QuuInv[j * m + lane], c0);
        }
        gr = c0 + c1;
        s.g[lane] = gr;
        d.kappa[((size_t)b * T + t) * m + lane] = a0 + a1;
      }
      __syncwarp();
      // dV = sum_r g_r Qu_r: butterfly over the warp (lanes >= m hold 0)
      double dv = gr * qu;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
      if (lane == 0) d.dV[(size_t)b * T + t] = dv;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        if (k < n) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0;
          int j = 0;
          for (; j + 2 < m; j += 3) {
            a0 = fma(s.g[j], s.Qux[j * n + k], a0);
            a1 = fma(s.g[j + 1], s.Qux[(j + 1) * n + k], a1);
            a2 = fma(s.g[j + 2], s.Qux[(j + 2) * n + k], a2);
          }
          for (; j < m; ++j) a0 = fma(s.g[j], s.Qux[j * n + k], a0);
          const double qx = (h == 0 ? lx0 : lx1) + s.QxM[k];
          s.W[k * LDW + n] = qx - ((a0 + a1) + a2);
        }
      }
    }
    BWD_TICK(9);
    __syncthreads();
    BWD_TICK(10);
  }
  if (tid == 0 && slot_word) atomicAnd(slot_word, ~(1 << slot));
#ifdef DDP_BWD_PROFILE
  if (prof_on)
    for (int i = 0; i < 12; ++i) g_bwd_prof[prof_cta][role][i] = prof_acc[i];
#endif
}

}  // namespace ddp
