// Does compute-sanitizer racecheck understand mbarrier-ordered shared-memory hand-offs?
// Variants: 0 inline-PTX arrive/try_wait (as in backward_sym.cuh), 1 cuda::barrier (libcu++),
// 2 inline PTX with explicit .release.cta / .acquire.cta, 3 named barrier arrive/sync
#include <cstdio>
#include <cstdint>
#include <cuda/barrier>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int V>
__global__ void k(double* out) {
  __shared__ double buf[64];
  __shared__ alignas(8) uint64_t bar;
  __shared__ cuda::barrier<cuda::thread_scope_block> cb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
    init(&cb, 33);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  buf[threadIdx.x & 63] = 1.0;
  __syncthreads();
  uint32_t par = 0;
  double acc = 0;
  for (int it = 0; it < 4; ++it) {
    if (warp == 0) {
      acc += buf[lane] + buf[lane + 32];   // read
      __syncwarp();
      if (V == 0) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory"); }
      if (V == 2) { if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory"); }
      if (V == 1) { if (lane == 0) (void)cb.arrive(); }
      if (V == 3) asm volatile("bar.arrive 1, 64;" ::: "memory");
    } else {
      if (V == 0) {
        asm volatile("{\n.reg .pred p;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D0;\nbra W0;\nD0:\n}\n" ::"r"(s32(&bar)), "r"(par) : "memory");
      }
      if (V == 2) {
        asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(s32(&bar)), "r"(par) : "memory");
      }
      if (V == 1) cb.arrive_and_wait();
      if (V == 3) asm volatile("bar.sync 1, 64;" ::: "memory");
      par ^= 1;
      buf[lane] = it;        // write after the reader is done
      buf[lane + 32] = it;
    }
    __syncthreads();
  }
  out[threadIdx.x] = acc;
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
