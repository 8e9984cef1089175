// lds_probe2.cu — per-warp issue cost of shared-memory loads with no dependent arithmetic in the loop:
// 32 back-to-back loads per body (volatile asm, results xor-combined once per body).
#include <cuda_runtime.h>
#include <stdio.h>
#define REP 128
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int W>   // W = 1, 2, 4 words per lane
__global__ void k(unsigned* out, long long* cyc, int spread) {
  extern __shared__ __align__(16) unsigned sm[];
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = sa(sm) + (spread ? lane : (lane & 7)) * 4 * W;
  unsigned acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
    unsigned v[32][4];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const unsigned a = base + ((j * 37 + r) & 63) * 128 * W;
      if (W == 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3]) : "r"(a));
      else if (W == 2) { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[j][0]), "=r"(v[j][1]) : "r"(a)); v[j][2] = v[j][3] = 0; }
      else { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[j][0]) : "r"(a)); v[j][1] = v[j][2] = v[j][3] = 0; }
    }
    unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) x ^= v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
    acc += x;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMallocManaged(&cyc, 64);
  for (int nw = 1; nw <= 8; nw *= 2)
    for (int spread = 0; spread < 2; ++spread) {
#define RUN(W) cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); k<W><<<1, 32 * nw, 49152>>>(out, cyc, spread); cudaDeviceSynchronize(); \
      printf("warps %d %s  LDS.%-3d %6.2f cycles per load per warp\n", nw, spread ? "32 distinct addresses" : "8 distinct (broadcast)", 32 * W, (double)cyc[0] / REP / 32);
      RUN(1) RUN(2) RUN(4)
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
