// lds_probe.cu — shared-memory load issue/throughput on sm_100a: cycles per LDS for 16 independent loads per loop body.
//   pattern 0: LDS.128, all quarter-warps read the same 128 B (1 wavefront)
//   pattern 1: LDS.128, 32 distinct float4 (512 B, 4 wavefronts)
//   pattern 2: LDS.32, 32 consecutive words (1 wavefront)
//   pattern 3: LDS.64, 32 distinct float2 (256 B)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define REP 128
template <int PAT>
__global__ void k(float* out, long long* cyc) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
    float4 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int row = ((r + j) & 15) + 16 * (w & 3);
      if (PAT == 0) v[j] = reinterpret_cast<const float4*>(sm)[row * 32 + (lane & 7)];
      else if (PAT == 1) v[j] = reinterpret_cast<const float4*>(sm)[row * 32 + lane];
      else if (PAT == 2) { v[j].x = sm[row * 128 + lane]; v[j].y = v[j].z = v[j].w = 0.f; }
      else { const float2 t = reinterpret_cast<const float2*>(sm)[row * 64 + lane]; v[j].x = t.x; v[j].y = t.y; v[j].z = v[j].w = 0.f; }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[PAT] = t1 - t0;
}
int main(int argc, char** argv) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMallocManaged(&cyc, 64);
  const char* names[] = {"LDS.128 1 wavefront (128 B)", "LDS.128 4 wavefronts (512 B)", "LDS.32 1 wavefront", "LDS.64 2 wavefronts"};
  for (int nw = 1; nw <= 16; nw *= 2) {
#define RUN(P) cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); k<P><<<1, 32 * nw, 49152>>>(out, cyc); cudaDeviceSynchronize(); \
    printf("warps %2d  %-32s %6.2f cycles per LDS per warp, %6.2f per LDS on the SM\n", nw, names[P], (double)cyc[P] / REP / 16, (double)cyc[P] / REP / 16 / nw);
    RUN(0) RUN(1) RUN(2) RUN(3)
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
