// lds_probe3.cu — what does a shared-memory load cost the SM's pipe, by width and by number of active lanes?
// NW warps of one CTA issue independent volatile loads back to back (8 per loop trip, results folded with LOP3 into
// eight independent words), conflict-free addresses.  Reported: cycles per load as seen by one warp, and SM cycles
// per load instruction (= the pipe's occupancy per instruction once enough warps are issuing).
#include <cuda_runtime.h>
#include <stdio.h>
#define REP 256
template <int WIDTH>
__global__ void k(unsigned* out, long long* cyc, int active) {
  extern __shared__ __align__(16) unsigned sm[];
  for (int i = threadIdx.x; i < 10240; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane * (4 * WIDTH) + (threadIdx.x >> 5) * 2048;
  unsigned s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  long long t0 = clock64();
  if (lane < active) {
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        unsigned x = 0, y = 0, z = 0, w = 0;
        if (WIDTH == 4)
          asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + (r & 3) * 512));
        else if (WIDTH == 2)
          asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(base + (r & 3) * 512));
        else
          asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x) : "r"(base + (r & 3) * 512));
        s[r] ^= x ^ y ^ z ^ w;
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = s[0] ^ s[1] ^ s[2] ^ s[3] ^ s[4] ^ s[5] ^ s[6] ^ s[7];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64);
  const int act[4] = {32, 24, 16, 8};
  for (int w = 0; w < 3; ++w)
    for (int a = 0; a < 4; ++a)
      for (int nw = 1; nw <= 16; nw *= 4) {
        if (w == 0) k<4><<<1, 32 * nw, 40960>>>(out, cyc, act[a]);
        if (w == 1) k<2><<<1, 32 * nw, 40960>>>(out, cyc, act[a]);
        if (w == 2) k<1><<<1, 32 * nw, 40960>>>(out, cyc, act[a]);
        cudaDeviceSynchronize();
        printf("LDS.%-3d active lanes %2d warps %2d: %5.2f cycles per load per warp, %5.2f SM cycles per load\n", w == 0 ? 128 : w == 1 ? 64 : 32, act[a], nw,
               (double)cyc[0] / REP / 8, (double)cyc[0] / REP / 8 / nw);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
