// What does compute-sanitizer synccheck accept?  Two warps meet at a CTA-wide named barrier from
// two different code locations (warp-specialised roles).  V=0: bar.sync 0, 64; V=1: bar.sync 1, 64;
// V=2: barrier.sync 0, 64 (non-aligned); V=3: barrier.sync 1, 64
#include <cstdio>
#include <cstdlib>
template <int V>
__device__ __forceinline__ void meet() {
  if (V == 0) asm volatile("bar.sync 0, 64;" ::: "memory");
  if (V == 1) asm volatile("bar.sync 1, 64;" ::: "memory");
  if (V == 2) asm volatile("barrier.sync 0, 64;" ::: "memory");
  if (V == 3) asm volatile("barrier.sync 1, 64;" ::: "memory");
}
template <int V>
__device__ __noinline__ double role_a(double* buf, int lane) {
  double acc = 0;
  for (int it = 0; it < 4; ++it) {
    buf[lane] = it;
    meet<V>();
    acc += buf[lane + 32];
    meet<V>();
  }
  return acc;
}
template <int V>
__device__ __noinline__ double role_b(double* buf, int lane) {
  double acc = 1;
  for (int it = 0; it < 4; ++it) {
    buf[lane + 32] = 2 * it;
    meet<V>();
    acc *= buf[lane] + 1.0;
    meet<V>();
  }
  return acc;
}
template <int V>
__global__ void k(double* out) {
  __shared__ double buf[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  out[threadIdx.x] = (warp == 0) ? role_a<V>(buf, lane) : role_b<V>(buf, lane);
}
int main(int argc, char** argv) {
  double* o;
  cudaMalloc(&o, 64 * 8);
  int v = argc > 1 ? atoi(argv[1]) : 0;
  if (v == 0) k<0><<<1, 64>>>(o);
  if (v == 1) k<1><<<1, 64>>>(o);
  if (v == 2) k<2><<<1, 64>>>(o);
  if (v == 3) k<3><<<1, 64>>>(o);
  printf("variant %d: %s\n", v, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
