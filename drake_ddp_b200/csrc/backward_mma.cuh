// K6 (large-n path): backward Riccati sweep on the fp64 tensor pipe.
//
// Replaces _backward_pass (/root/reference/ilqr.py:623-667) for models with n >= 16.  One CTA
// per trajectory, sequential over t = N-2 .. 0.  Per step the dense contractions
//   W = Vxx fx, Wu = Vxx fu, Qxx = lxx + fx' W, Qux = fu' W, Quu = luu + fu' Wu,
//   K = Quu^-1 Qux, Vxx <- Qxx - Qux' K
// run as mma.sync.m8n8k4 f64 (DMMA) tiles out of shared memory; Vxx / Vx never leave the SM.
// The fx / fu / x_bar / u_bar tiles of step t-1 are prefetched into the other half of a
// double buffer by 1-D bulk TMA (cp.async.bulk + mbarrier) while step t computes; when the
// tile sizes are not 16-byte multiples (odd n) the kernel falls back to a cooperative copy.
// Warp roles: warps 0..TN-1 each own one 8-column strip of the n x n products, the last warp
// does the vectors (Qx, Qu), Quu, its inverse (np.linalg.inv stand-in, ilqr.py:655), kappa,
// dV and Vx.
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

template <int n, int m>
struct BwdMmaCfg {
  static constexpr int TN = (n + 7) / 8, TM = (m + 7) / 8, KN = (n + 3) / 4, KM = (m + 3) / 4;
  static constexpr int NW = TN + 1, NT = NW * 32;
  // bulk TMA needs 16-byte sizes and 16-byte aligned tile starts in global memory
  static constexpr bool TMA = ((n * n) % 2 == 0) && ((n * m) % 2 == 0) && (n % 2 == 0) && (m % 2 == 0);
  static constexpr int even(int v) { return (v + 1) & ~1; }
  static constexpr int MINB = (n <= 36) ? 4 : 3;  // CTAs per SM the register budget is sized for
};

template <int n, int m>
struct BwdMmaSmem {
  typedef BwdMmaCfg<n, m> C;
  alignas(16) double Fx[2][C::even(n * n)];   // double-buffered bulk-TMA destination
  alignas(16) double Fu[C::even(n * m)];      // single buffer: refilled after phase 2 (dead by then)
  alignas(16) double xb[2][C::even(n)];
  alignas(16) double ub[2][C::even(m)];
  alignas(16) double Vxx[n * n];
  double W[n * n];
  double WuKt[n * m];   // Wu = Vxx fu (phases 1-2), then K_t (phase 3)
  double Qux[m * n];
  double QuuInv[m * m]; // Quu, overwritten by its inverse
  double Vx[n], Qx[n], Qu[m], g[m];
  alignas(8) uint64_t bar[2];
  alignas(8) uint64_t barFu;
};

template <class Model>
__global__ void __launch_bounds__(BwdMmaCfg<Model::n, Model::m>::NT, BwdMmaCfg<Model::n, Model::m>::MINB)
backward_mma_kernel(Dev d) {
  constexpr int n = Model::n, m = Model::m;
  typedef BwdMmaCfg<n, m> C;
  constexpr int TN = C::TN, TM = C::TM, KN = C::KN, KM = C::KM, NT = C::NT;
  const int b = blockIdx.x;
  if (!d.active[b]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdMmaSmem<n, m>& s = *reinterpret_cast<BwdMmaSmem<n, m>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int N = d.N, T = d.T;
  const double* Q = d.Q;
  const double* R = d.R;
  const double* Qf = d.Qf;
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
  auto copy_tile = [&](int t, int buf) {    // all threads (fallback)
    for (int i = tid; i < n * n; i += NT) s.Fx[buf][i] = gfx[(size_t)t * n * n + i];
    for (int i = tid; i < n * m; i += NT) s.Fu[i] = gfu[(size_t)t * n * m + i];
    for (int i = tid; i < n; i += NT) s.xb[buf][i] = gxb[(size_t)t * n + i];
    for (int i = tid; i < m; i += NT) s.ub[buf][i] = gub[(size_t)t * m + i];
  };

  if (C::TMA) {
    if (tid == 0) {
      mbar_init(&s.bar[0], 1);
      mbar_init(&s.bar[1], 1);
      mbar_init(&s.barFu, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      issue_tile(T - 1, 0);
      issue_fu(T - 1);
    }
  }

  // Vx, Vxx <- terminal cost partials at x_bar[:, -1]            (ilqr.py:638, 203-204)
  {
    const double* xl = gxb + (size_t)(N - 1) * n;
    for (int i = tid; i < n * n; i += NT) s.Vxx[i] = 2.0 * Qf[i];
    for (int i = tid; i < n; i += NT) {
      double a = 0.0, c = 0.0;
      for (int j = 0; j < n; ++j) {
        a = fma(2.0 * Qf[i * n + j], xl[j], a);
        c = fma(2.0 * xnom[j], Qf[j * n + i], c);
      }
      s.Vx[i] = a - c;
    }
  }
  __syncthreads();

  uint32_t parity[2] = {0, 0};
  uint32_t parityFu = 0;
  int buf = 0;
  for (int t = T - 1; t >= 0; --t, buf ^= 1) {
    if (C::TMA) {
      if (tid == 0 && t > 0) issue_tile(t - 1, buf ^ 1);
      mbar_wait(&s.bar[buf], parity[buf]);
      parity[buf] ^= 1;
      mbar_wait(&s.barFu, parityFu);
      parityFu ^= 1;
    } else {
      copy_tile(t, buf);
      __syncthreads();
    }
    const double* Fx = s.Fx[buf];
    const double* Fu = s.Fu;

    // ---------------- phase 1: Wu = Vxx fu and Qx (per strip) ; last warp: Qu -------------------
    if (warp < TN) {
      const int w = warp;
      double accu[TM][2];
#pragma unroll
      for (int i = 0; i < TM; ++i) accu[i][0] = accu[i][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KN; ++kk) {
        const int k = 4 * kk + tg;
        const double af = ldz(s.Vxx, n, 8 * w + g, k, n, n);
#pragma unroll
        for (int nt = 0; nt < TM; ++nt) dmma(accu[nt], af, ldz(Fu, m, k, 8 * nt + g, n, m));
      }
#pragma unroll
      for (int nt = 0; nt < TM; ++nt) {
        const int r = 8 * w + g, c = 8 * nt + 2 * tg;
        if (r < n && c < m) s.WuKt[r * m + c] = accu[nt][0];
        if (r < n && c + 1 < m) s.WuKt[r * m + c + 1] = accu[nt][1];
      }
      // Qx[k] = lx[k] + sum_i fx[i][k] Vx[i] for the 8 columns of this strip   (ilqr.py:651,180)
      // lane = (column g8, quarter q4 of the i range); quarters are combined with two shuffles
      {
        const double* xb = s.xb[buf];
        const int g8 = lane & 7, q4 = lane >> 3, k = 8 * w + g8;
        double q = 0.0;
        if (k < n) {
          for (int i = q4; i < n; i += 4) q = fma(Fx[i * n + k], s.Vx[i], q);
        }
        q += __shfl_xor_sync(0xffffffffu, q, 8);
        q += __shfl_xor_sync(0xffffffffu, q, 16);
        if (q4 == 0 && k < n) {
          double a = 0.0, c = 0.0;
          if (d.diag_cost) {
            a = 2.0 * Q[k * n + k] * xb[k];
            c = 2.0 * xnom[k] * Q[k * n + k];
          } else {
            for (int j = 0; j < n; ++j) {
              a = fma(2.0 * Q[k * n + j], xb[j], a);
              c = fma(2.0 * xnom[j], Q[j * n + k], c);
            }
          }
          s.Qx[k] = (a - c) + q;
        }
      }
    } else {
      // Qu = lu + fu' Vx                                          (ilqr.py:652,181)
      const double* ub = s.ub[buf];
      for (int r = lane; r < m; r += 32) {
        double a = 0.0;
        for (int j = 0; j < m; ++j) a = fma(2.0 * R[r * m + j], ub[j], a);
        double q1 = 0.0;
        int i = 0;
        for (; i + 1 < n; i += 2) {
          a = fma(Fu[i * m + r], s.Vx[i], a);
          q1 = fma(Fu[(i + 1) * m + r], s.Vx[i + 1], q1);
        }
        for (; i < n; ++i) a = fma(Fu[i * m + r], s.Vx[i], a);
        s.Qu[r] = a + q1;
      }
    }
    __syncthreads();

    // ---------------- phase 2 ---------------------------------------------------------------
    // strips: Quu tiles first (hand-off to the last warp through named barrier 2), then
    // W = Vxx fx, then Qxx (into Vxx) and Qux.  Last warp: waits for Quu, inverts it.
    if (warp < TN) {
      const int w = warp;
      // Quu = luu + fu' Wu                                        (ilqr.py:654)
      for (int q = w; q < TM * TM; q += TN) {
        const int mt = q / TM, nt = q % TM;
        double a2[2] = {0.0, 0.0};
#pragma unroll
        for (int kk = 0; kk < KN; ++kk) {
          const int k = 4 * kk + tg;
          dmma(a2, ldz(Fu, m, k, 8 * mt + g, n, m), ldz(s.WuKt, m, k, 8 * nt + g, n, m));
        }
        const int r = 8 * mt + g, c = 8 * nt + 2 * tg;
        if (r < m && c < m) s.QuuInv[r * m + c] = 2.0 * R[r * m + c] + a2[0];
        if (r < m && c + 1 < m) s.QuuInv[r * m + c + 1] = 2.0 * R[r * m + c + 1] + a2[1];
      }
      asm volatile("bar.arrive 2, %0;" ::"r"(NT) : "memory");
      {
        double acc[TN][2];
#pragma unroll
        for (int i = 0; i < TN; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < KN; ++kk) {
          const int k = 4 * kk + tg;
          const double bf = ldz(Fx, n, k, 8 * w + g, n, n);
#pragma unroll
          for (int mt = 0; mt < TN; ++mt) dmma(acc[mt], ldz(s.Vxx, n, 8 * mt + g, k, n, n), bf);
        }
#pragma unroll
        for (int mt = 0; mt < TN; ++mt) {
          const int r = 8 * mt + g, c = 8 * w + 2 * tg;
          if (r < n && c < n) s.W[r * n + c] = acc[mt][0];
          if (r < n && c + 1 < n) s.W[r * n + c + 1] = acc[mt][1];
        }
      }
      // all strips of W written and all reads of the old Vxx done (strip warps only)
      asm volatile("bar.sync 1, %0;" ::"r"(TN * 32) : "memory");
      double aq[TN][2], au[TM][2];
#pragma unroll
      for (int i = 0; i < TN; ++i) aq[i][0] = aq[i][1] = 0.0;
#pragma unroll
      for (int i = 0; i < TM; ++i) au[i][0] = au[i][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KN; ++kk) {
        const int k = 4 * kk + tg;
        const double bf = ldz(s.W, n, k, 8 * w + g, n, n);
#pragma unroll
        for (int mt = 0; mt < TN; ++mt) dmma(aq[mt], ldz(Fx, n, k, 8 * mt + g, n, n), bf);
#pragma unroll
        for (int mt = 0; mt < TM; ++mt) dmma(au[mt], ldz(Fu, m, k, 8 * mt + g, n, m), bf);
      }
#pragma unroll
      for (int mt = 0; mt < TN; ++mt) {
        const int r = 8 * mt + g, c = 8 * w + 2 * tg;
        if (r < n && c < n) s.Vxx[r * n + c] = 2.0 * Q[r * n + c] + aq[mt][0];
        if (r < n && c + 1 < n) s.Vxx[r * n + c + 1] = 2.0 * Q[r * n + c + 1] + aq[mt][1];
      }
#pragma unroll
      for (int mt = 0; mt < TM; ++mt) {
        const int r = 8 * mt + g, c = 8 * w + 2 * tg;
        if (r < m && c < n) s.Qux[r * n + c] = au[mt][0];
        if (r < m && c + 1 < n) s.Qux[r * n + c + 1] = au[mt][1];
      }
    } else {
      asm volatile("bar.sync 2, %0;" ::"r"(NT) : "memory");
      invert_warp<m>(s.QuuInv, s.QuuInv);                               // ilqr.py:655
    }
    __syncthreads();

    // fu of this step is dead now: refill the single Fu buffer with the next step's tile
    if (C::TMA && tid == 0 && t > 0) issue_fu(t - 1);

    // ---------------- phase 3: K = Quu^-1 Qux ; Vxx = Qxx - Qux' K ; last warp: kappa, dV, Vx ---
    if (warp < TN) {
      const int w = warp;
      double ak[TM][2];
#pragma unroll
      for (int i = 0; i < TM; ++i) ak[i][0] = ak[i][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) {
        const int k = 4 * kk + tg;
        const double bf = ldz(s.Qux, n, k, 8 * w + g, m, n);
#pragma unroll
        for (int mt = 0; mt < TM; ++mt) dmma(ak[mt], ldz(s.QuuInv, m, 8 * mt + g, k, m, m), bf);
      }
      double* gK = d.K + ((size_t)b * T + t) * m * n;
#pragma unroll
      for (int mt = 0; mt < TM; ++mt) {
        const int r = 8 * mt + g, c = 8 * w + 2 * tg;
        if (r < m && c < n) {
          s.WuKt[r * n + c] = ak[mt][0];
          gK[r * n + c] = ak[mt][0];
        }
        if (r < m && c + 1 < n) {
          s.WuKt[r * n + c + 1] = ak[mt][1];
          gK[r * n + c + 1] = ak[mt][1];
        }
      }
      __syncwarp();
      double av[TN][2];
#pragma unroll
      for (int i = 0; i < TN; ++i) av[i][0] = av[i][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) {
        const int k = 4 * kk + tg;
        const double bf = ldz(s.WuKt, n, k, 8 * w + g, m, n);
#pragma unroll
        for (int mt = 0; mt < TN; ++mt) dmma(av[mt], ldz(s.Qux, n, k, 8 * mt + g, m, n), bf);
      }
#pragma unroll
      for (int mt = 0; mt < TN; ++mt) {
        const int r = 8 * mt + g, c = 8 * w + 2 * tg;
        if (r < n && c < n) s.Vxx[r * n + c] -= av[mt][0];
        if (r < n && c + 1 < n) s.Vxx[r * n + c + 1] -= av[mt][1];
      }
    } else {
      // kappa = Quu^-1 Qu ; g = Qu' Quu^-1 ; dV = g Qu ; Vx = Qx - g Qux   (ilqr.py:659,663,666)
      for (int r = lane; r < m; r += 32) {
        double a = 0.0, c = 0.0;
        for (int j = 0; j < m; ++j) {
          a = fma(s.QuuInv[r * m + j], s.Qu[j], a);
          c = fma(s.Qu[j], s.QuuInv[j * m + r], c);
        }
        s.g[r] = c;
        d.kappa[((size_t)b * T + t) * m + r] = a;
      }
      __syncwarp();
      if (lane == 0) {
        double a = 0.0;
        for (int j = 0; j < m; ++j) a = fma(s.g[j], s.Qu[j], a);
        d.dV[(size_t)b * T + t] = a;
      }
      for (int k = lane; k < n; k += 32) {
        double a = 0.0;
        for (int j = 0; j < m; ++j) a = fma(s.g[j], s.Qux[j * n + k], a);
        s.Vx[k] = s.Qx[k] - a;
      }
    }
    __syncthreads();
  }
}

}  // namespace ddp
