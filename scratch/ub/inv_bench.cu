// micro-benchmark: latency of one 12x12 warp inversion in isolation and with N co-resident warps
#include <cstdio>
#include <cuda_runtime.h>
#include "../../drake_ddp_b200/csrc/backward_mma.cuh"
using namespace ddp;

template <int VAR>
__global__ void k(const double* A, double* out, long long* cyc, int reps) {
  __shared__ double M[8][144];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 144; i += 32) M[w][i] = A[i];
  __syncwarp();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (VAR == 0) invert_warp<12>(M[w], M[w]);
    else invert_warp_fast<12>(M[w], M[w]);
  }
  long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + w] = (t1 - t0) / reps;
  for (int i = lane; i < 144; i += 32) out[i] = M[w][i];
}
__global__ void lat(double* out, long long* cyc) {
  // dependent-chain latencies: DFMA, SHFL, REDUX
  double x = threadIdx.x * 1e-3 + 1.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) x = fma(x, 1.0000001, 1e-9);
  long long t1 = clock64();
  unsigned u = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) u = __reduce_max_sync(0xffffffffu, u + i);
  long long t2 = clock64();
  double y = x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) y = __shfl_sync(0xffffffffu, y, (i + threadIdx.x) & 31);
  long long t3 = clock64();
  double z = y + 2.0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z = rcp_fast(z) + 1.5;
  long long t4 = clock64();
  unsigned b = u;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) b = __ffs(__ballot_sync(0xffffffffu, (b + threadIdx.x) & 1)) + i;
  long long t5 = clock64();
  if (threadIdx.x == 0) { cyc[0] = (t1 - t0) / 256; cyc[1] = (t2 - t1) / 256; cyc[2] = (t3 - t2) / 256; cyc[3] = (t4 - t3) / 256; cyc[4] = (t5 - t4) / 256; }
  out[threadIdx.x] = x + u + y + z + b;
}
int main() {
  double hA[144];
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) hA[i * 12 + j] = (i == j ? 3.0 : 0.0) + 0.1 * ((i * 7 + j * 3) % 5) + 0.1 * ((j * 7 + i * 3) % 5);
  double *A, *out; long long* cyc;
  cudaMalloc(&A, sizeof(hA)); cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 8 * 1024 * 8);
  cudaMemcpy(A, hA, sizeof(hA), cudaMemcpyHostToDevice);
  long long h[16];
  lat<<<1, 32>>>(out, cyc); cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
  printf("dependent latency: DFMA %lld  REDUX(+iadd) %lld  SHFL64 %lld  rcp_fast+dadd %lld  ballot+ffs %lld\n", h[0], h[1], h[2], h[3], h[4]);
  for (int var = 0; var < 2; ++var)
    for (int warps = 1; warps <= 8; warps *= 2) {
      if (var == 0) k<0><<<1, 32 * warps>>>(A, out, cyc, 2); else k<1><<<1, 32 * warps>>>(A, out, cyc, 2);
      cudaDeviceSynchronize();
      if (var == 0) k<0><<<1, 32 * warps>>>(A, out, cyc, 1); else k<1><<<1, 32 * warps>>>(A, out, cyc, 1);
      cudaMemcpy(h, cyc, 8 * warps, cudaMemcpyDeviceToHost);
      printf("variant %d, %d warps in the CTA: cycles per inversion (warp 0) %lld  err %s\n", var, warps, h[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
