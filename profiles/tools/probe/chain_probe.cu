// chain_probe.cu — the K2r chain-warp unit (16 rows: 3 LDS.128 + 8 FMUL + 8 FADD per row, two chains) in isolation:
// cycles per unit for one warp alone, with NW-1 other warps doing the same on the SM, with LDS-heavy neighbours.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define REP 64
__global__ void k(float* out, long long* cyc, int mode) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 24576 + 4096; i += blockDim.x) sm[i] = 1.0f + (float)(i % 97) * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = lane & 7, kp = min(lane >> 3, 2);
  const float4* sdk = reinterpret_cast<const float4*>(sm) + 2 * kp * 256 + c;
  const float4* pdb = reinterpret_cast<const float4*>(sm + 24576) + c;
  float acc0 = -0.0f, acc1 = -0.0f;
  long long t0 = clock64();
  if (mode == 0 || w == 0) {
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      const float4* sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
      const float4* pd4 = pdb + (u & 1) * 320;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 d[8], a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          d[r] = pd4[(8 * h + r) * 10];
          a[r] = sd4[(8 * h + r) * 8];
          b[r] = sd4[256 + (8 * h + r) * 8];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          acc0 = acc0 + a[r].x * d[r].x; acc1 = acc1 + b[r].x * d[r].x;
          acc0 = acc0 + a[r].y * d[r].y; acc1 = acc1 + b[r].y * d[r].y;
          acc0 = acc0 + a[r].z * d[r].z; acc1 = acc1 + b[r].z * d[r].z;
          acc0 = acc0 + a[r].w * d[r].w; acc1 = acc1 + b[r].w * d[r].w;
        }
      }
    }
  } else {
    // neighbours: producer-like traffic (2 LDS.32 + 1 STS.32 + 8 FP per row)
    float* win = sm + 24576 + 1024 * (w & 1);
    float a0 = win[lane + 1], b0 = win[lane], r = 0.f;
#pragma unroll 1
    for (int u = 0; u < REP * 2; ++u) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a1 = win[((j + 1) & 15) * 40 + lane + 1], b1 = win[((j + 1) & 15) * 40 + lane];
        const float pn = ((0.3f * a1 + 0.2f * b1) + 0.4f * a0) + 0.1f * b0;
        win[j * 40 + 4 * (lane & 7) + (lane >> 3) + 512] = r - pn;
        a0 = a1; b0 = b1; r = pn;
      }
    }
    acc0 = r;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc0 + acc1;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMallocManaged(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000);
  for (int mode = 0; mode < 2; ++mode)
    for (int nw = 1; nw <= 8; nw *= 2) {
      k<<<1, 32 * nw, 116000>>>(out, cyc, mode); cudaDeviceSynchronize();
      printf("mode %d (%s) warps %d: %7.1f cycles per unit (16 rows) on warp 0\n", mode, mode ? "warp 0 chain, others producer-like" : "all warps chain", nw, (double)cyc[0] / REP);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
