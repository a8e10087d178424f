// micro-benchmark: DMMA (mma.sync.m8n8k4.f64) issue rate per warp vs warps per SM sub-partition and
// independent accumulators, operands in registers or fetched from shared memory.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int ACC, bool LDS>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ double sm[4 * 36 * 40];
  for (int i = threadIdx.x; i < 4 * 36 * 40; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  const double* pa = sm + (warp & 3) * 36 * 40 + g * 40 + tg;
  const double* pb = sm + (warp & 3) * 36 * 40 + tg * 40 + g;
  double a = lane * 1e-3, b = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {
      if (LDS) {
        double av[2], bv[5];
        av[0] = pa[4 * kk]; av[1] = pa[4 * kk + 8 * 40];
#pragma unroll
        for (int j = 0; j < 5; ++j) bv[j] = pb[4 * kk * 40 + 8 * j];
#pragma unroll
        for (int i = 0; i < ACC; ++i) dmma(c[i], av[i & 1], bv[i % 5]);
      } else {
#pragma unroll
        for (int i = 0; i < ACC; ++i) dmma(c[i], a, b);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ACC, bool LDS>
void run(int warps, double* out, long long* cyc) {
  const int iters = 200;
  k<ACC, LDS><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  k<ACC, LDS><<<148, 32 * warps>>>(out, cyc, iters);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_warp = (double)h / (iters * 9.0 * ACC);
  const double per_smsp = per_warp / ((warps + 3) / 4);
  printf("acc=%2d lds=%d warps/SM=%2d  cycles per DMMA: per warp %.1f, per sub-partition %.1f\n", ACC, (int)LDS, warps,
         per_warp, per_smsp);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  for (int w : {1, 4, 8, 12, 16}) run<10, false>(w, out, cyc);
  for (int w : {4, 8, 12, 16}) run<4, false>(w, out, cyc);
  for (int w : {4, 8}) run<1, false>(w, out, cyc);
  for (int w : {4, 8, 12, 16}) run<10, true>(w, out, cyc);
  return 0;
}
