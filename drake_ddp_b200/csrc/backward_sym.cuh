// K6 (large-n path): backward Riccati sweep on the fp64 tensor pipe, symmetric stacked form.
//
// Replaces _backward_pass (/root/reference/ilqr.py:623-667) for models with n >= 16.  One CTA
// (3 DMMA warps + 1 vector warp) per trajectory, sequential over t = N-2 .. 0; Vxx, Vx never
// leave the SM.  With the stacked Jacobian S = [fx | fu] (n x (n+m)) the five Q-terms of
// ilqr.py:651-656 are the blocks of ONE symmetric matrix
//
//        M = S' Vxx S = [ fx'Vxx fx   fx'Vxx fu ]   =  [ Qxx - lxx    Qux'     ]
//                       [ fu'Vxx fx   fu'Vxx fu ]      [ Qux          Quu - luu ]
//
// so only the tiles on and above the diagonal are computed (21 of 36 at n+m = 48) and mirrored
// where the lower block is needed.  The work is dealt by 8-wide COLUMN strips of S: the warp that
// owns strip c computes W_c = Vxx S_c (all rows) into its private slice of shared memory and
// then M(r, c) = S_r' W_c for r <= c from that slice alone -- no CTA barrier between the two
// products; the strips are dealt (5,0) (4,1) (3,2).  The same warp owns column strip c of the
// gains and of the Vxx update: it computes K_c = Quu^-1 Qux_c and the tiles r <= c of
// Vxx = Qxx - Qux_r' K_c, with its own tiles Qxx(r, c) = lxx + M(r, c) -- still in registers -- as
// the accumulators of the products against -K_c.  Per step that leaves two CTA barriers (inputs
// ready / Quu handed to the vector warp) and two mbarrier hand-offs (Q-terms and Quu^-1 complete /
// fx, fu released); past the first nobody reads the old Vxx, so the update may overwrite it.
//   DMMAs per step at (n, m) = (36, 12): 270 (W) + 189 (M) + 30 (K) + 45 (update) = 534, against
//   681 for the unsymmetric schedule of round 1; the inverse adds 24 per Newton-Schulz pass.
// The strips that hold fu columns go first, so Quu exists after a third of the products and the
// vector warp inverts it (ilqr.py:655; Newton-Schulz on the tensor pipe seeded with the previous
// step's inverse, Gauss-Jordan with partial pivoting as fallback) under the rest.  The vector
// warp also owns the vectors: lx, lu, Qx = lx + fx'Vx, Qu = lu + fu'Vx, kappa, dV, Vx.
// Symmetry: the reference never symmetrises Vxx, so its Vxx is symmetric up to rounding; here
// the off-diagonal tiles are exact mirrors.  The difference is of the order of one rounding per
// step (tests/test_gpu_parity.py::test_backward_pass_teacher_forced: 1e-9 on K, kappa, dV).
// fx tiles of step t-1 arrive by 1-D bulk TMA (cp.async.bulk + mbarrier) into the other half of
// a double buffer while step t computes, fu is refilled as soon as it is dead; odd n or m fall
// back to 8-byte cp.async copies issued by all threads.
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
// position of column j (or of the even-aligned column pair starting at j) of row k in a W / K strip
__device__ __forceinline__ int bs_wcol(int k, int j) { return j ^ ((k & 2) << 1); }
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// whole-warp call
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
  // lanes can leave the spin loop in different iterations and the compiler does not see the branch:
  // reconverge before the next warp-collective instruction (DMMA, bar.sync are .aligned)
  __syncwarp();
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Named barriers between warp-specialised roles: the warps of a CTA meet at these from DIFFERENT
// code locations (one copy per role function).  `bar.sync` is `barrier.sync.aligned`: PTX asks the
// threads of ONE warp to execute the same instruction, which holds here.
//
// Two limits of compute-sanitizer shape the test-only build -DDDP_SANITIZER_BUILD
// (libddp_b200_racecheck.so, tests/test_gpu_parity.py::test_compute_sanitizer_...):
//  * synccheck reports an aligned barrier that two warps reach from two code locations as
//    "divergent threads in block" (scratch/ub/sync_named.cu reproduces it in 20 lines) and accepts
//    the non-aligned `barrier.sync`, which costs 5 % of the sweep (2.89 -> 3.05 ms): the product
//    keeps the aligned form, the sanitizer build uses the non-aligned one;
//  * racecheck does not model mbarrier ordering: it flags any mbarrier-ordered hand-off, libcu++'s
//    cuda::barrier included (scratch/ub/race_mbar.cu).  The hand-offs between the warps of a CTA
//    (old Vxx released / Q-terms complete / fx, fu released) are mbarrier arrive + wait pairs -- a
//    warp arrives when ITS part is done and only waits where it needs the others; the sanitizer
//    build does them with bar.sync / bar.arrive at the wait points (stricter: a wait becomes a
//    rendezvous), everything else is the same code.
#ifndef DDP_BWD_BARQ_NAMED
#define DDP_BWD_BARQ_NAMED 0
#endif
constexpr bool kBsBarQNamed = DDP_BWD_BARQ_NAMED != 0;   // experiment: "Q-terms complete" as bar.sync instead of an mbarrier
#ifdef DDP_SANITIZER_BUILD
constexpr bool kBsNamed = true;
#define DDP_BAR_SYNC "barrier.sync"
#define DDP_BAR_ARRIVE "barrier.arrive"
#else
constexpr bool kBsNamed = false;
#define DDP_BAR_SYNC "bar.sync"
#define DDP_BAR_ARRIVE "bar.arrive"
#endif
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile(DDP_BAR_SYNC " %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile(DDP_BAR_ARRIVE " %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// items 0 .. count-1 with cost base + i + 1 dealt to three warps, longest first to the least
// loaded warp; every warp's list comes out in descending order
struct BsDeal {
  int cnt[3];
  int list[3][8];
};
constexpr BsDeal bs_deal(int count, int base) {
  BsDeal d = {};
  int load[3] = {0, 0, 0};
  for (int i = count - 1; i >= 0; --i) {
    int w = 0;
    for (int k = 1; k < 3; ++k)
      if (load[k] < load[w]) w = k;
    d.list[w][d.cnt[w]++] = i;
    load[w] += base + i + 1;
  }
  return d;
}
constexpr int bs_max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
// the entries below `limit` of every warp's list, in the same order
constexpr BsDeal bs_below(BsDeal in, int limit) {
  BsDeal d = {};
  for (int w = 0; w < 3; ++w)
    for (int p = 0; p < in.cnt[w]; ++p)
      if (in.list[w][p] < limit) d.list[w][d.cnt[w]++] = in.list[w][p];
  return d;
}
// for every entry of `sub` its position in the same warp's list of `all`
constexpr BsDeal bs_positions(BsDeal all, BsDeal sub) {
  BsDeal d = {};
  for (int w = 0; w < 3; ++w) {
    d.cnt[w] = sub.cnt[w];
    for (int q = 0; q < sub.cnt[w]; ++q)
      for (int p = 0; p < all.cnt[w]; ++p)
        if (all.list[w][p] == sub.list[w][q]) d.list[w][q] = p;
  }
  return d;
}

template <int n, int m>
struct BsCfg {
  static constexpr int TN = (n + 7) / 8, TM = (m + 7) / 8, KN = (n + 3) / 4, KM = (m + 3) / 4;
  static constexpr int NS = n + m, TS = (NS + 7) / 8;  // stacked [fx | fu] columns / 8-wide strips
  static constexpr int TQ = n / 8;                     // first stacked strip that holds an fu column
  static constexpr int NMW = 3, NW = NMW + 1, NT = NW * 32;  // 3 DMMA warps + the vector warp
  static_assert(TS <= 8 && TN <= 8, "strip lists hold 8 entries");
  // Bulk TMA needs 16-byte sizes and 16-byte aligned addresses on both sides.  A tile with an odd
  // number of doubles (n = 37: fx; n = 27, m = 7: fx and fu) starts 16-byte aligned only at every
  // other step: its copy is rounded out to the enclosing 16-byte window (one double more, before or
  // after the tile; the arrays are 256-byte aligned and carry 16 bytes of slack) and the tile then
  // starts at element 0 or 1 of the buffer.  -DDDP_BWD_NO_TMA: 8-byte cp.async copies by all threads
  // instead (the round-1 path for odd sizes; 17 % slower at n = 36).
#ifdef DDP_BWD_NO_TMA
  static constexpr bool TMA = false;
#else
  static constexpr bool TMA = true;
#endif
  // fx buffers: two (the next tile travels while the step computes) up to n = 36; ONE above, where
  // two would cost the fourth CTA per SM (n = 37: 61.6 KB against 50.6 KB): the next tile is then
  // requested as soon as every DMMA warp is through its products (barS) and lands under the gains /
  // Vxx update and the other CTAs of the SM
  static constexpr int NFX = (TMA && n > 36) ? 1 : 2;
  static constexpr bool ODD_FX = (n * n) % 2 != 0, ODD_FU = (n * m) % 2 != 0;
  static constexpr uint32_t FX_BYTES = (n * n + (ODD_FX ? 1 : 0)) * 8, FU_BYTES = (n * m + (ODD_FU ? 1 : 0)) * 8;
  // tile `idx` (= b T + t) of an array of tiles of `elems` doubles: first double of the 16-byte
  // window that holds its start, and where the tile starts inside the window (0 or 1)
  static __device__ __forceinline__ int tile_pre(size_t idx, int elems) { return (int)((idx * (size_t)elems) & 1); }
  static constexpr bool EVEN = (n % 2 == 0) && (m % 2 == 0);  // C-fragment pairs never straddle n or rows
  static constexpr int even(int v) { return (v + 1) & ~1; }
  // stacked strips dealt to the DMMA warps: W_c, the tiles M(r <= c, c) and, for c < TN, the
  // column strip c of K with its update tiles
  static constexpr BsDeal SD = bs_deal(TS, TN);
  static constexpr int MAXS = bs_max3(SD.cnt[0], SD.cnt[1], SD.cnt[2]);
  // column strips of K (K_c and its update tiles): the warp's own Vxx strips, so that a tile
  // Qxx(r, c) = lxx + M(r, c) never leaves the registers between the products and the update
  static constexpr BsDeal KD = bs_below(SD, TN);
  static constexpr BsDeal KP = bs_positions(SD, KD);   // where a K strip sits among the warp's stacked strips
  static constexpr int MAXK = bs_max3(KD.cnt[0], KD.cnt[1], KD.cnt[2]);
  // leading dimension of Vxx in shared memory: the A-operand fetch of the W product reads rows g,
  // columns 4 kk + tg; a half-warp (g = 0..3) is conflict-free when the row stride is 8 banks mod 32,
  // i.e. LDV = 4 mod 16 doubles: 36 as it is, 37 -> 52, 27 -> 36 (DDP_BWD_LDV_PLAIN: LDV = n)
#ifdef DDP_BWD_LDV_PLAIN
  static constexpr int LDV = n;
#else
  static constexpr int LDV = (n % 16 == 4) ? n : ((n + 11) / 16) * 16 + 4;
#endif
#ifndef DDP_BWD_MINB
#define DDP_BWD_MINB 4
#endif
#ifndef DDP_BWD_MINB_BIG
#define DDP_BWD_MINB_BIG 4
#endif
  static constexpr int MINB = (n <= 36) ? DDP_BWD_MINB : DDP_BWD_MINB_BIG;  // CTAs per SM the register budget is sized for
  static_assert(n <= 64 && m <= 32, "vector warp keeps lx in two registers per lane and lu in one");
  static_assert(n % 8 != 0, "Vx rides along as row n of the last (partly empty) row tile of Vxx");
};

template <int n, int m>
struct BsSmem {
  typedef BsCfg<n, m> C;
  alignas(16) double Fx[C::NFX][C::even(n * n)];   // bulk-TMA destination (double-buffered up to n = 36)
  alignas(16) double Fu[C::even(n * m)];      // single buffer: refilled once the Q-terms exist
  alignas(16) double Vxx[C::even(n * C::LDV)];
  // W_c = Vxx S_c, one private [n][8] slice per stacked strip; once the Q-terms exist the same
  // memory holds the column strips of K, [m][8] each
  // Column j of row k sits at position j ^ (4 * bit1(k)) of the row (bs_wcol): the B-operand fetch
  // of an m8n8k4 DMMA reads rows k = 4 kk + tg, columns g, and a half-warp (g = 0..3, all tg) of
  // 8-byte accesses must cover 32 distinct banks; in the plain [k][8] layout rows tg and tg + 2 of
  // a half-warp share their banks (2-way conflict: 35 M of 481 M wavefronts per launch in r2a)
  alignas(16) double Wf[C::TS * n * 8];
  alignas(16) double Qux[C::even(m * n)];
  alignas(16) double Quu[C::even(m * m)];
  alignas(16) double QuuInv[C::even(m * m)];  // Quu^-1; kept across steps: it seeds the next inversion
  alignas(16) double NsR[C::even(m * m)];     // Newton-Schulz residual
  double Vx[C::even(n)];                      // Vx; rides along as row n of Vxx in the W product, so
  double SVx[8 * C::TS];                      // SVx = S' Vx = [fx' Vx | fu' Vx] costs no extra DMMA
  double QxM[n];                              // Qx
  double Qu[m], g[m], Qd2[n];                 // Qd2 = diagonal of lxx = 2 Q
  alignas(8) uint64_t bar[2];                 // fx tile landed (per buffer)
  alignas(8) uint64_t barFu;
  alignas(8) uint64_t barQ;                   // Qux complete and Quu^-1 ready
  alignas(8) uint64_t barS;                   // every DMMA warp is done reading fx / fu of the step
  int slot;                                   // CTA slot on this SM (deals the warp roles)
};
// four CTAs per SM at the headline shape: 4 x (this + 1 KB reserved) must fit in 227 KB
static_assert(sizeof(BsSmem<36, 12>) <= 57088, "backward_sym_kernel: shared memory budget for 4 CTAs/SM");

// Inverse of the m x m matrix A by Newton-Schulz iteration on the fp64 tensor pipe, one warp:
//   R = I - A X ;  X <- X + X R      (the residual squares every pass)
// started from the inverse of the previous backward step, which is still in X: Quu moves little
// between neighbouring timesteps, so three passes on average (24 DMMAs each at m = 12) reach full
// precision, against roughly a thousand dependent-latency-bound instructions of Gauss-Jordan.
// X is updated in place.  Accepted when max|R_ij| < 2^-26 before a pass, i.e. ||R||_inf < m 2^-26 <=
// 2^-22 for m <= 16: after that pass the residual is bounded by 2^-44 = 5.7e-14, as small as
// LAPACK's getrf/getri leave I - Quu X for the condition numbers seen here (np.linalg.inv,
// ilqr.py:655); not "below an ulp".  The test runs from the third pass on (2.8 passes were needed on
// average when every pass was tested).  Returns false -- X is then garbage -- when the start does
// not contract (row sums of |R| >= 1/2, or NaN) or 8 passes are not enough: the caller falls back
// to Gauss-Jordan with partial pivoting (invert_warp).  Rs: m*m doubles of scratch.
template <int m>
__device__ __forceinline__ bool invert_newton_warp(const double* A, double* X, double* Rs) {
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
  bool ok = false;
  static_assert(m <= 16, "the acceptance bounds below assume m <= 16");
  constexpr unsigned kContract = 0x3FA00000u, kAccept = 0x3E500000u;   // high words of 2^-5, 2^-26
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
      for (int nt = 0; nt < TM; ++nt) bb[nt] = ldz(X, m, 4 * kk + tg, 8 * nt + g, m, m);
#pragma unroll
      for (int mt = 0; mt < TM; ++mt)
#pragma unroll
        for (int nt = 0; nt < TM; ++nt) dmma(y[mt][nt], a[mt], bb[nt]);
    }
    // R = I - A X and the largest |entry| of it (integer max over the high words: NaN and inf
    // come out on top); m * max|R_ij| bounds the infinity norm
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
    // the first two passes run unchecked (a diverging start shows in the third): the reduction
    // sits on the serial path of the step
    const bool check = it >= 2;
    if (check) hmax = __reduce_max_sync(full, hmax);
    if (check && hmax >= kContract) break;  // max|R| >= 2^-5 (||R||_inf may be >= 1/2), inf or NaN
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
      for (int mt = 0; mt < TM; ++mt) a[mt] = ldz(X, m, 8 * mt + g, 4 * kk + tg, m, m);
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) bb[nt] = ldz(Rs, m, 4 * kk + tg, 8 * nt + g, m, m);
#pragma unroll
      for (int mt = 0; mt < TM; ++mt)
#pragma unroll
        for (int nt = 0; nt < TM; ++nt) dmma(xc[mt][nt], a[mt], bb[nt]);
    }
    __syncwarp();  // every read of X done
#pragma unroll
    for (int mt = 0; mt < TM; ++mt)
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) {
        const int r = 8 * mt + g, c = 8 * nt + 2 * tg;
        if (r < m && c < m) X[r * m + c] = xc[mt][nt][0];
        if (r < m && c + 1 < m) X[r * m + c + 1] = xc[mt][nt][1];
      }
    __syncwarp();
    if (check && hmax < kAccept) {  // max|R| < 2^-26 before this pass: ||R||_inf < m 2^-26 <= 2^-22
      ok = true;
      break;
    }
  }
  return ok;
}

#ifdef DDP_BWD_PROFILE
// per-phase cycle totals of the four warps of two CTAs (b = 5: first wave, b = 700: second wave)
__device__ long long g_bwd_prof[2][4][16];
#define BS_PROF_DECL                                                         \
  const int prof_cta = (x.b == 5) ? 0 : ((x.b == 700) ? 1 : -1);            \
  const bool prof_on = prof_cta >= 0 && x.lane == 0;                         \
  long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; \
  long long prof_last = clock64();
#define BS_TICK(i)                                 \
  do {                                             \
    if (prof_on) {                                 \
      const long long now_ = clock64();            \
      prof_acc[i] += now_ - prof_last;             \
      prof_last = now_;                            \
    }                                              \
  } while (0)
#define BS_PROF_STORE(role)                                                        \
  if (prof_on)                                                                     \
    for (int i = 0; i < 16; ++i) g_bwd_prof[prof_cta][role][i] = prof_acc[i];
#else
#define BS_PROF_DECL
#define BS_TICK(i)
#define BS_PROF_STORE(role)
#endif

// Everything a warp needs, shared by the role functions.
template <int n, int m>
struct BsCtx {
  BsSmem<n, m>* s;
  Dev d;
  int b, lane, g, tg, tid;
  size_t tile0;   // index of the trajectory's first (b, t) tile: b T
  const double *Q, *R, *gfx, *gfu, *gxb, *gub, *xnom;
  bool diag;
};

// stacked S = [fx | fu]: pointer to element S[tg][col] of the current tiles and the lane's k-stride
template <int n, int m>
__device__ __forceinline__ void bs_stacked(const double* Fx, const double* Fu, int tg, int col, const double*& p,
                                           int& s4) {
  col = min(col, n + m - 1);
  const bool in_fx = col < n;
  p = in_fx ? (Fx + tg * n + col) : (Fu + tg * m + (col - n));
  s4 = in_fx ? 4 * n : 4 * m;
}

// ---- one DMMA warp: W is the warp's index (0..2), everything about its strips is static ----------
// element routing of a stacked tile M(r, c): (i, j) stacked row / column, v the value: the
// elements that belong to Qux / Quu are stored (VXX = false; the Qxx elements stay in the warp's
// registers and become the accumulators of the Vxx update).
template <int n, int m, bool VXX>
__device__ __forceinline__ void bs_route(BsSmem<n, m>& s, const BsCtx<n, m>& x, int i, int j, double v, bool off) {
  static_assert(!VXX, "Qxx tiles are no longer routed through shared memory");
  constexpr int NS = n + m;
  if (i >= NS || j >= NS) return;
  if (i < n && j < n) {                        // Qxx = lxx + fx' Vxx fx        (ilqr.py:653): see the update
  } else if (!VXX) {
    if (i < n) {                               // Qxu = Qux'                    (ilqr.py:656)
      s.Qux[(j - n) * n + i] = v;
    } else if (j >= n) {                       // Quu = luu + fu' Vxx fu         (ilqr.py:654)
      const int a = i - n, bq = j - n;
      s.Quu[a * m + bq] = 2.0 * x.R[a * m + bq] + v + ((a == bq) ? x.d.quu_reg : 0.0);
      if (off) s.Quu[bq * m + a] = 2.0 * x.R[bq * m + a] + v;   // a != bq in an off-diagonal tile
    }
    // i >= n > j: below the diagonal of the tile that straddles n; the mirror element above the
    // diagonal already delivers this entry of Qux (one writer per entry: bit-reproducible)
  }
}

// W_c = Vxx S_c for one or two stacked strips (one sweep over Vxx feeds both), into the strips'
// private slices of Wf.  Row n of the A operand is Vx', so row n of the result is (S' Vx) over
// the strip: Qx and Qu cost no extra DMMA.
template <int n, int m>
__device__ __forceinline__ void bs_w_strips(BsSmem<n, m>& s, const double* Fx, const double* Fu, int g, int tg,
                                            int c0, int c1, bool two) {
  typedef BsCfg<n, m> C;
  constexpr int TN = C::TN, KN = C::KN, LDV = C::LDV;
  const double *pb0, *pb1;
  int sb0, sb1;
  bs_stacked<n, m>(Fx, Fu, tg, 8 * c0 + g, pb0, sb0);
  bs_stacked<n, m>(Fx, Fu, tg, 8 * c1 + g, pb1, sb1);
  const double* pa[TN];
#pragma unroll
  for (int i = 0; i < TN; ++i)
    pa[i] = (8 * i + g == n) ? (s.Vx + tg) : (s.Vxx + min(8 * i + g, n - 1) * LDV + tg);
  double acc[TN][2][2];
#pragma unroll
  for (int i = 0; i < TN; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < KN; ++kk) {
    const bool kin = (4 * kk + 3 < n) || (4 * kk + tg < n);
    double a[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) a[i] = kin ? pa[i][4 * kk] : 0.0;
    const double b0 = kin ? pb0[kk * sb0] : 0.0;
    const double b1 = (two && kin) ? pb1[kk * sb1] : 0.0;
#pragma unroll
    for (int i = 0; i < TN; ++i) {
      dmma(acc[i][0], a[i], b0);
      if (two) dmma(acc[i][1], a[i], b1);
    }
  }
  double* w0 = s.Wf + c0 * (n * 8);
  double* w1 = s.Wf + c1 * (n * 8);
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    const int r = 8 * i + g;
    if (r < n) {
      *reinterpret_cast<double2*>(w0 + r * 8 + bs_wcol(r, 2 * tg)) = make_double2(acc[i][0][0], acc[i][0][1]);
      if (two) *reinterpret_cast<double2*>(w1 + r * 8 + bs_wcol(r, 2 * tg)) = make_double2(acc[i][1][0], acc[i][1][1]);
    } else if (r == n) {
      *reinterpret_cast<double2*>(s.SVx + 8 * c0 + 2 * tg) = make_double2(acc[i][0][0], acc[i][0][1]);
      if (two) *reinterpret_cast<double2*>(s.SVx + 8 * c1 + 2 * tg) = make_double2(acc[i][1][0], acc[i][1][1]);
    }
  }
}

template <int n, int m, int W>
__device__ __forceinline__ void bs_dmma_warp(const BsCtx<n, m>& x) {
  typedef BsCfg<n, m> C;
  constexpr int TN = C::TN, TM = C::TM, KN = C::KN, KM = C::KM, TS = C::TS, TQ = C::TQ, NT = C::NT;
  constexpr int LDV = C::LDV;
  constexpr bool EVEN = C::EVEN;
  constexpr BsDeal SD = C::SD, KD = C::KD, KP = C::KP;
  constexpr int NSTR = SD.cnt[W], NKST = KD.cnt[W];
  BsSmem<n, m>& s = *x.s;
  const Dev& d = x.d;
  const int g = x.g, tg = x.tg, T = d.T;
  uint32_t parity[2] = {0, 0}, parityFu = 0, parityQ = 0, parityS = 0;
  (void)parityS;
  int buf = 0;
  BS_PROF_DECL
  for (int t = T - 1; t >= 0; --t, buf = (buf ^ 1) & (C::NFX - 1)) {
    BS_TICK(0);
    // ---- inputs of this step: fx (TMA), fu, and the Vxx / Vx the previous step left -------------
    if (C::TMA) {
      mbar_wait(&s.bar[buf], parity[buf]);
      parity[buf] ^= 1;
      mbar_wait(&s.barFu, parityFu);
      parityFu ^= 1;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    BS_TICK(1);
    named_bar_sync(0, NT);
    BS_TICK(2);
    if (!C::TMA && t > 0) {   // the next step's fx starts travelling into the other half
      for (int i = x.tid; i < n * n; i += NT)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fx[(buf ^ 1) & (C::NFX - 1)][i])),
                     "l"(x.gfx + (size_t)(t - 1) * n * n + i)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const double* Fx = s.Fx[buf] + ((C::TMA && C::ODD_FX) ? C::tile_pre(x.tile0 + t, n * n) : 0);
    const double* Fu = s.Fu + ((C::TMA && C::ODD_FU) ? C::tile_pre(x.tile0 + t, n * m) : 0);

    // ---- phase 1, first part: W_c = Vxx S_c for the strips that hold fu columns (c >= TQ): Quu
    // = luu + fu' Vxx fu starts the serial path of the step (Quu -> its inverse -> K -> Vxx) and
    // comes out of these strips alone ---------------------------------------------------------------
    constexpr int NHIW = (NSTR > 0 && SD.list[W][0] >= TQ) ? ((NSTR > 1 && SD.list[W][1] >= TQ) ? 2 : 1) : 0;
    static_assert(NSTR < 3 || SD.list[W][NSTR > 2 ? 2 : 0] < TQ, "at most two strips with fu columns per warp");
    if (NHIW > 0) bs_w_strips<n, m>(s, Fx, Fu, g, tg, SD.list[W][0], SD.list[W][NHIW - 1], NHIW == 2);
    else named_bar_arrive(2, NT);   // nothing of Quu comes from this warp
    __syncwarp();
    // ---- phase 2a: the tiles that hold Quu (rows r >= TQ of the strips c >= TQ) first: Quu is
    // handed to the vector warp, which inverts it under the rest of the products --------------------
    double hacc[C::MAXS][TS - TQ][2];
#pragma unroll
    for (int p0 = 0; p0 < NSTR; ++p0) {
      const int c = SD.list[W][p0];
      if (c < TQ) continue;
      const double* pw = s.Wf + c * (n * 8) + tg * 8 + bs_wcol(tg, g);   // rows 4 kk + tg: bit 1 is tg's
      const double* pa[TS - TQ];
      int sa[TS - TQ];
#pragma unroll
      for (int q = 0; q < TS - TQ; ++q) {
        hacc[p0][q][0] = hacc[p0][q][1] = 0.0;
        bs_stacked<n, m>(Fx, Fu, tg, 8 * min(TQ + q, c) + g, pa[q], sa[q]);
      }
#pragma unroll
      for (int kk = 0; kk < KN; ++kk) {
        const bool kin = (4 * kk + 3 < n) || (4 * kk + tg < n);
        const double bw = kin ? pw[kk * 32] : 0.0;
#pragma unroll
        for (int q = 0; q < TS - TQ; ++q)
          if (TQ + q <= c) dmma(hacc[p0][q], kin ? pa[q][kk * sa[q]] : 0.0, bw);
      }
#pragma unroll
      for (int q = 0; q < TS - TQ; ++q)
        if (TQ + q <= c) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            bs_route<n, m, false>(s, x, 8 * (TQ + q) + g, 8 * c + 2 * tg + e, hacc[p0][q][e], TQ + q < c);
        }
    }
    if (NHIW > 0) named_bar_arrive(2, NT);   // Quu complete as far as this warp is concerned
    BS_TICK(4);
    // ---- phase 1, second part: W_c for the other strips, two per sweep over Vxx ----------------
#pragma unroll
    for (int p0 = NHIW; p0 < NSTR; p0 += 2) {
      const bool two = (p0 + 1 < NSTR);
      bs_w_strips<n, m>(s, Fx, Fu, g, tg, SD.list[W][p0], SD.list[W][two ? p0 + 1 : p0], two);
    }
    __syncwarp();
    BS_TICK(3);

    // ---- phase 2b: the other tiles M(r, c), r <= c, r < TQ, all strips of the warp in one sweep
    // over S (an operand fragment of S feeds every strip that needs the row tile).  The tiles stay
    // in registers: their Qux / Quu elements are stored now, their Qxx elements become the
    // accumulators of the Vxx update below ---------------------------------------------------------
    constexpr int RMAX = (TQ < TS) ? TQ : TS;      // row tiles 0 .. RMAX-1 can be needed
    double acc[C::MAXS][RMAX > 0 ? RMAX : 1][2];
    {
      const double* pw[C::MAXS];
#pragma unroll
      for (int p0 = 0; p0 < C::MAXS; ++p0) {
        const int c = SD.list[W][p0 < NSTR ? p0 : 0];
        pw[p0] = s.Wf + c * (n * 8) + tg * 8 + bs_wcol(tg, g);
#pragma unroll
        for (int r = 0; r < RMAX; ++r) acc[p0][r][0] = acc[p0][r][1] = 0.0;
      }
      const double* pa[RMAX > 0 ? RMAX : 1];
      int sa[RMAX > 0 ? RMAX : 1];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) bs_stacked<n, m>(Fx, Fu, tg, 8 * r + g, pa[r], sa[r]);
#pragma unroll
      for (int kk = 0; kk < KN; ++kk) {
        const bool kin = (4 * kk + 3 < n) || (4 * kk + tg < n);
        double bw[C::MAXS];
#pragma unroll
        for (int p0 = 0; p0 < C::MAXS; ++p0) bw[p0] = (p0 < NSTR && kin) ? pw[p0][kk * 32] : 0.0;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          bool need = false;
#pragma unroll
          for (int p0 = 0; p0 < NSTR; ++p0) need = need || (r <= SD.list[W][p0]);
          if (!need) continue;
          const double a = kin ? pa[r][kk * sa[r]] : 0.0;
#pragma unroll
          for (int p0 = 0; p0 < NSTR; ++p0)
            if (r <= SD.list[W][p0]) dmma(acc[p0][r], a, bw[p0]);
        }
      }
      __syncwarp();
      // this warp no longer reads fx / fu of the step
      if (kBsNamed) {
        if (C::TMA) named_bar_arrive(5, NT);
      } else if (x.lane == 0) mbar_arrive(&s.barS);
      BS_TICK(5);
      BS_TICK(6);
#pragma unroll
      for (int p0 = 0; p0 < NSTR; ++p0) {
        const int c = SD.list[W][p0];
        if (c < TQ) continue;                      // strips below TQ hold Qxx elements only
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (r <= c) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              bs_route<n, m, false>(s, x, 8 * r + g, 8 * c + 2 * tg + e, acc[p0][r][e], r < c);
          }
      }
    }
    // ---- all Q-terms complete and Quu^-1 ready ---------------------------------------------------
    __syncwarp();
    BS_TICK(7);
    if (kBsNamed || kBsBarQNamed) named_bar_sync(4, NT);
    else {
      if (x.lane == 0) mbar_arrive(&s.barQ);
      mbar_wait(&s.barQ, parityQ);
    }
    parityQ ^= 1;
    BS_TICK(8);
    if (!C::TMA) {
      // cp.async path: fu is refilled by all threads once nobody reads it any more
      if (kBsNamed) named_bar_sync(5, NT);
      else mbar_wait(&s.barS, parityS);
      parityS ^= 1;
      if (t > 0) {
        for (int i = x.tid; i < n * m; i += NT)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fu[i])),
                       "l"(x.gfu + (size_t)(t - 1) * n * m + i)
                       : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }

    // ---- phase C: for the column strips of K dealt to this warp: K_c = Quu^-1 Qux_c, then the
    // tiles r <= c of Vxx -= Qux_r' K_c.  The strips go through each stage together: short dependent
    // DMMA chains interleave and the operand fragments of Quu^-1 / Qux feed every strip -----------
    double* gK = d.K + ((size_t)x.b * T + t) * m * n;
    {
      double kacc[C::MAXK][TM][2];
      const double* pb[C::MAXK];
#pragma unroll
      for (int p0 = 0; p0 < C::MAXK; ++p0) {
        const int c = KD.list[W][p0 < NKST ? p0 : 0];
        pb[p0] = s.Qux + tg * n + min(8 * c + g, n - 1);
#pragma unroll
        for (int i = 0; i < TM; ++i) kacc[p0][i][0] = kacc[p0][i][1] = 0.0;
      }
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) {
        const bool kin = (4 * kk + 3 < m) || (4 * kk + tg < m);
        double bq[C::MAXK];
#pragma unroll
        for (int p0 = 0; p0 < C::MAXK; ++p0) bq[p0] = (p0 < NKST && kin) ? pb[p0][kk * 4 * n] : 0.0;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const double a = kin ? s.QuuInv[min(8 * i + g, m - 1) * m + 4 * kk + tg] : 0.0;
#pragma unroll
          for (int p0 = 0; p0 < NKST; ++p0)
            dmma(kacc[p0][i], a, bq[p0]);
        }
      }
#pragma unroll
      for (int p0 = 0; p0 < NKST; ++p0) {
        const int c = KD.list[W][p0];
        double* kt = s.Wf + c * (n * 8);      // -K strip, [m][8] (the W slices are dead by now): the update subtracts
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int r = 8 * i + g, col = 8 * c + 2 * tg;
          if (r < m) {
            *reinterpret_cast<double2*>(kt + r * 8 + bs_wcol(r, 2 * tg)) = make_double2(-kacc[p0][i][0], -kacc[p0][i][1]);
            if (EVEN) {
              if (col < n) *reinterpret_cast<double2*>(gK + r * n + col) = make_double2(kacc[p0][i][0], kacc[p0][i][1]);
            } else {
              if (col < n) gK[r * n + col] = kacc[p0][i][0];
              if (col + 1 < n) gK[r * n + col + 1] = kacc[p0][i][1];
            }
          }
        }
      }
    }
    __syncwarp();
    {
      // Vxx = Qxx - Qux' Quu^-1 Qux (ilqr.py:667), tiles r <= c of the warp's strips: the
      // accumulators start from Qxx = lxx + M (ilqr.py:653; M is still in this warp's registers)
      // and the products run against -K_c.  Every warp is past its W products here (barQ), so the
      // old Vxx is dead; the mirror element is an exact copy.
      double uacc[C::MAXK][TN][2];
      const double* pk[C::MAXK];
#pragma unroll
      for (int p0 = 0; p0 < C::MAXK; ++p0) {
        const int c = KD.list[W][p0 < NKST ? p0 : 0];
        const int sp = KP.list[W][p0 < NKST ? p0 : 0];   // the strip's place among the warp's stacked strips
        pk[p0] = s.Wf + c * (n * 8) + tg * 8 + bs_wcol(tg, g);
#pragma unroll
        for (int r = 0; r < TN; ++r) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = 0.0;
            if (p0 < NKST && r <= c) {
              v = (r < RMAX) ? acc[sp][r < RMAX ? r : 0][e] : hacc[sp][(r >= TQ && r - TQ < TS - TQ) ? r - TQ : 0][e];
              const int i = 8 * r + g, j = 8 * c + 2 * tg + e;
              if (x.diag) {
                if (r == c) v += (i == j && i < n) ? s.Qd2[i] : 0.0;
              } else if (i < n && j < n) {
                v += 2.0 * x.Q[i * n + j];
              }
            }
            uacc[p0][r][e] = v;
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) {
        const bool kin = (4 * kk + 3 < m) || (4 * kk + tg < m);
        double bk[C::MAXK];
#pragma unroll
        for (int p0 = 0; p0 < C::MAXK; ++p0) bk[p0] = (p0 < NKST && kin) ? pk[p0][kk * 32] : 0.0;
#pragma unroll
        for (int r = 0; r < TN; ++r) {
          bool need = false;
#pragma unroll
          for (int p0 = 0; p0 < NKST; ++p0) need = need || (r <= KD.list[W][p0]);
          if (!need) continue;
          const double a = kin ? s.Qux[(4 * kk + tg) * n + min(8 * r + g, n - 1)] : 0.0;
#pragma unroll
          for (int p0 = 0; p0 < NKST; ++p0)
            if (r <= KD.list[W][p0]) dmma(uacc[p0][r], a, bk[p0]);
        }
      }
#pragma unroll
      for (int p0 = 0; p0 < NKST; ++p0) {
        const int c = KD.list[W][p0];
#pragma unroll
        for (int r = 0; r < TN; ++r) {
          if (r > c) continue;
          const bool off = r < c;
          const int i = 8 * r + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 8 * c + 2 * tg + e;
            if (i < n && j < n) {
              const double v = uacc[p0][r][e];
              s.Vxx[i * LDV + j] = v;
              if (off) s.Vxx[j * LDV + i] = v;
            }
          }
        }
      }
    }
    BS_TICK(9);
  }
  BS_PROF_STORE(W)
}

// ---- the vector warp: cost gradients, Qx / Qu, the inverse of Quu, kappa, dV, Vx, the TMA queue ----
template <int n, int m>
__device__ __forceinline__ void bs_vector_warp(const BsCtx<n, m>& x) {
  typedef BsCfg<n, m> C;
  constexpr int NT = C::NT;
  BsSmem<n, m>& s = *x.s;
  const Dev& d = x.d;
  const int lane = x.lane, T = d.T;
  const double* R = x.R;
  const double* Q = x.Q;
  const bool diag = x.diag;
  // step-independent half of lx = 2 Q x - 2 x_nom' Q                     (ilqr.py:180)
  double lxc0 = 0.0, lxc1 = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = lane + 32 * h;
    double c = 0.0;
    if (k < n) {
      if (diag) c = 2.0 * x.xnom[k] * Q[k * n + k];
      else
        for (int j = 0; j < n; ++j) c = fma(2.0 * x.xnom[j], Q[j * n + k], c);
    }
    if (h == 0) lxc0 = c;
    else lxc1 = c;
  }
  const double rd2 = (lane < m) ? 2.0 * R[lane * m + lane] : 0.0;
  // x_bar_t, u_bar_t of the step: fetched one step ahead straight into registers
  double xb0 = (lane < n) ? x.gxb[(size_t)(T - 1) * n + lane] : 0.0;
  double xb1 = (lane + 32 < n) ? x.gxb[(size_t)(T - 1) * n + lane + 32] : 0.0;
  double ubl = (lane < m) ? x.gub[(size_t)(T - 1) * m + lane] : 0.0;
  uint32_t parity[2] = {0, 0}, parityFu = 0, parityQ = 0, parityS = 0;
  int buf = 0;
  BS_PROF_DECL
  for (int t = T - 1; t >= 0; --t, buf = (buf ^ 1) & (C::NFX - 1)) {
    BS_TICK(0);
    if (C::TMA) {
      if (C::NFX == 2 && lane == 0 && t > 0) {   // next step's fx into the other half of the double buffer
        mbar_expect_tx(&s.bar[buf ^ 1], C::FX_BYTES);
        tma_load_1d(s.Fx[(buf ^ 1) & (C::NFX - 1)], x.gfx + (size_t)(t - 1) * n * n - C::tile_pre(x.tile0 + t - 1, n * n),
                    C::FX_BYTES, &s.bar[buf ^ 1]);
      }
      mbar_wait(&s.bar[buf], parity[buf]);
      parity[buf] ^= 1;
      mbar_wait(&s.barFu, parityFu);
      parityFu ^= 1;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    BS_TICK(1);
    named_bar_sync(0, NT);
    BS_TICK(2);
    if (!C::TMA && t > 0) {
      for (int i = x.tid; i < n * n; i += NT)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fx[(buf ^ 1) & (C::NFX - 1)][i])),
                     "l"(x.gfx + (size_t)(t - 1) * n * n + i)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const double cxb0 = xb0, cxb1 = xb1, cub = ubl;
    if (t > 0) {
      xb0 = (lane < n) ? x.gxb[(size_t)(t - 1) * n + lane] : 0.0;
      xb1 = (lane + 32 < n) ? x.gxb[(size_t)(t - 1) * n + lane + 32] : 0.0;
      ubl = (lane < m) ? x.gub[(size_t)(t - 1) * m + lane] : 0.0;
    }
    // lx = 2 Q x - 2 x_nom' Q ; lu = 2 R u                                 (ilqr.py:180-181)
    double lx0 = 0.0, lx1 = 0.0, lu = 0.0;
    if (diag) {
      lx0 = ((lane < n) ? s.Qd2[lane] * cxb0 : 0.0) - lxc0;
      lx1 = ((lane + 32 < n) ? s.Qd2[lane + 32] * cxb1 : 0.0) - lxc1;
      lu = rd2 * cub;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = min(lane + 32 * h, n - 1);
        double a = 0.0;
        for (int j = 0; j < n; ++j) {
          const double xj = __shfl_sync(0xffffffffu, (j < 32) ? cxb0 : cxb1, j & 31);
          a = fma(2.0 * Q[k * n + j], xj, a);
        }
        if (h == 0) lx0 = a - lxc0;
        else lx1 = a - lxc1;
      }
      for (int j = 0; j < m; ++j) {
        const double uj = __shfl_sync(0xffffffffu, cub, j);
        if (lane < m) lu = fma(2.0 * R[lane * m + j], uj, lu);
      }
    }
    // Quu^-1 (ilqr.py:655): Newton-Schulz from the previous step's inverse, else Gauss-Jordan
    BS_TICK(3);
    named_bar_sync(2, NT);
    BS_TICK(4);
    // fu is only read by the strips / row tiles that hold fu columns, and those are done once Quu
    // exists: refill the single Fu buffer now, a whole inversion ahead of its next use
    if (C::TMA && lane == 0 && t > 0) {
      mbar_expect_tx(&s.barFu, C::FU_BYTES);
      tma_load_1d(s.Fu, x.gfu + (size_t)(t - 1) * n * m - C::tile_pre(x.tile0 + t - 1, n * m), C::FU_BYTES, &s.barFu);
    }
    if (t == T - 1 || (d.bwd_flags & 1) || !invert_newton_warp<m>(s.Quu, s.QuuInv, s.NsR))
      invert_warp<m>(s.Quu, s.QuuInv);
    __syncwarp();
    BS_TICK(5);
    // Quu^-1 is what the DMMA warps are waiting for: arrive now, wait (for Qux) only where Qux is
    // needed.  kappa, g and dV need Quu^-1 and Qu = lu + fu' Vx alone, and fu' Vx (S' Vx over the
    // strips that hold fu columns) was complete when Quu was handed over.
    if (!(kBsNamed || kBsBarQNamed) && lane == 0) mbar_arrive(&s.barQ);
    // kappa = Quu^-1 Qu ; g = Qu' Quu^-1 ; dV = g Qu                        (ilqr.py:659,663)
    if (lane < m) s.Qu[lane] = lu + s.SVx[n + lane];
    __syncwarp();
    double qu = 0.0, gr = 0.0;
    if (lane < m) {
      qu = s.Qu[lane];
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
      d.kappa[((size_t)x.b * T + t) * m + lane] = a0 + a1;
    }
    __syncwarp();
    double dv = gr * qu;   // dV = sum_r g_r Qu_r: butterfly over the warp (lanes >= m hold 0)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
    if (lane == 0) d.dV[(size_t)x.b * T + t] = dv;
    BS_TICK(6);
    if (kBsNamed || kBsBarQNamed) named_bar_sync(4, NT);
    else mbar_wait(&s.barQ, parityQ);   // all Q-terms complete
    parityQ ^= 1;
    // Qx = lx + fx' Vx (ilqr.py:651): S' Vx came out of the W products
    if (lane < n) s.QxM[lane] = lx0 + s.SVx[lane];
    if (lane + 32 < n) s.QxM[lane + 32] = lx1 + s.SVx[lane + 32];
    __syncwarp();
    // every DMMA warp is through its products: fx of this step is dead (the vector warp issues the
    // TMA of the step after next into this buffer at the top of the next step)
    if (kBsNamed) named_bar_sync(5, NT);
    else mbar_wait(&s.barS, parityS);
    parityS ^= 1;
    BS_TICK(7);
    if (C::TMA && C::NFX == 1 && lane == 0 && t > 0) {   // single fx buffer: free now, refill it
      mbar_expect_tx(&s.bar[0], C::FX_BYTES);
      tma_load_1d(s.Fx[0], x.gfx + (size_t)(t - 1) * n * n - C::tile_pre(x.tile0 + t - 1, n * n), C::FX_BYTES, &s.bar[0]);
    }
    if (!C::TMA && t > 0) {
      for (int i = x.tid; i < n * m; i += NT)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fu[i])),
                     "l"(x.gfu + (size_t)(t - 1) * n * m + i)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Vx = Qx - g Qux                                                        (ilqr.py:666)
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
        s.Vx[k] = s.QxM[k] - ((a0 + a1) + a2);
      }
    }
    BS_TICK(8);
  }
  BS_PROF_STORE(3)
}

template <class Model>
__global__ void __launch_bounds__(BsCfg<Model::n, Model::m>::NT, BsCfg<Model::n, Model::m>::MINB)
backward_sym_kernel(Dev d) {
  constexpr int n = Model::n, m = Model::m;
  typedef BsCfg<n, m> C;
  constexpr int NT = C::NT, NMW = C::NMW;
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BsSmem<n, m>& s = *reinterpret_cast<BsSmem<n, m>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = d.N, T = d.T;
  BsCtx<n, m> x;
  x.s = &s;
  x.d = d;
  x.b = b;
  x.lane = lane;
  x.g = lane >> 2;
  x.tg = lane & 3;
  x.tid = tid;
  x.Q = d.Q;
  x.R = d.R;
  x.diag = d.diag_cost != 0;
  x.xnom = d.x_nom + (size_t)b * n;
  x.tile0 = (size_t)b * T;
  x.gfx = d.fx + (size_t)b * T * n * n;
  x.gfu = d.fu + (size_t)b * T * n * m;
  x.gxb = d.x_bar + (size_t)b * N * n;
  x.gub = d.u_bar + (size_t)b * T * m;

  // Deal the warp roles by CTA slot: the CTAs resident on one SM take distinct slots (per-SM
  // bitmask in global memory), and the vector role goes to warp `slot`, so every sub-partition
  // (warp id % 4) hosts one vector warp and three DMMA warps instead of all vector warps on one.
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
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    mbar_init(&s.barFu, 1);
    mbar_init(&s.barQ, NMW + 1);
    mbar_init(&s.barS, NMW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int slot = s.slot;
#ifdef DDP_BWD_STAGGER
  {   // experiment: offset the CTAs of an SM by a fraction of a step
    const long long until = clock64() + (long long)slot * DDP_BWD_STAGGER;
    while (clock64() < until) {
    }
  }
#endif
  const int role = (warp - slot - 1) & 3;   // 0..2: DMMA warps; 3 (warp == slot): the vector warp
  if (C::TMA) {
    if (role == NMW && lane == 0) {
      mbar_expect_tx(&s.bar[0], C::FX_BYTES);
      tma_load_1d(s.Fx[0], x.gfx + (size_t)(T - 1) * n * n - C::tile_pre(x.tile0 + T - 1, n * n), C::FX_BYTES, &s.bar[0]);
      mbar_expect_tx(&s.barFu, C::FU_BYTES);
      tma_load_1d(s.Fu, x.gfu + (size_t)(T - 1) * n * m - C::tile_pre(x.tile0 + T - 1, n * m), C::FU_BYTES, &s.barFu);
    }
  } else {
    for (int i = tid; i < n * n; i += NT)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fx[0][i])),
                   "l"(x.gfx + (size_t)(T - 1) * n * n + i)
                   : "memory");
    for (int i = tid; i < n * m; i += NT)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s.Fu[i])),
                   "l"(x.gfu + (size_t)(T - 1) * n * m + i)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // Vx, Vxx <- terminal cost partials at x_bar[:, -1]            (ilqr.py:638, 203-204)
  {
    const double* Qf = d.Qf;
    const double* xl = x.gxb + (size_t)(N - 1) * n;
    for (int i = tid; i < n * n; i += NT) s.Vxx[(i / n) * C::LDV + (i % n)] = 2.0 * Qf[i];
    for (int i = tid; i < n; i += NT) {
      double a = 0.0, c = 0.0;
      for (int j = 0; j < n; ++j) {
        a = fma(2.0 * Qf[i * n + j], xl[j], a);
        c = fma(2.0 * x.xnom[j], Qf[j * n + i], c);
      }
      s.Vx[i] = a - c;
      s.Qd2[i] = 2.0 * x.Q[i * n + i];
    }
  }
  __syncthreads();
  if (role == 0) bs_dmma_warp<n, m, 0>(x);
  else if (role == 1) bs_dmma_warp<n, m, 1>(x);
  else if (role == 2) bs_dmma_warp<n, m, 2>(x);
  else bs_vector_warp<n, m>(x);
  __syncthreads();
  if (tid == 0 && slot_word) atomicAnd(slot_word, ~(1 << slot));
}

}  // namespace ddp
