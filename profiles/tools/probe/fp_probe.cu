// fp_probe.cu — single-warp issue rates on sm_100a, compiled with -fmad=false like the library:
//   dependent FADD chain, N independent FADD chains, FMUL+FADD mix as in the reference-order chain warp, LDS.128 feed.
#include <cuda_runtime.h>
#include <stdio.h>
#define REP 256
template <int MODE>
__global__ void k(float* out, const float* in, long long* cyc) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  float a0 = in[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = in[32 + threadIdx.x], c = in[64 + threadIdx.x];
  const float4* s4 = reinterpret_cast<const float4*>(sm) + (threadIdx.x & 7);
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
    if (MODE == 0) {          // 16 dependent FADD
#pragma unroll
      for (int j = 0; j < 16; ++j) a0 = a0 + b;
    } else if (MODE == 1) {   // 2 chains x 8
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0 = a0 + b; a1 = a1 + c; }
    } else if (MODE == 2) {   // 4 chains x 4
#pragma unroll
      for (int j = 0; j < 4; ++j) { a0 = a0 + b; a1 = a1 + c; a2 = a2 + b; a3 = a3 + c; }
    } else if (MODE == 3) {   // 8 chains x 2
#pragma unroll
      for (int j = 0; j < 2; ++j) { a0 = a0 + b; a1 = a1 + c; a2 = a2 + b; a3 = a3 + c; a4 = a4 + b; a5 = a5 + c; a6 = a6 + b; a7 = a7 + c; }
    } else if (MODE == 4) {   // chain warp: 8 x (FMUL, dependent FADD), operands in registers
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0 = a0 + a1 * b; a1 = a1 + c; }   // a1 update keeps the products distinct
    } else if (MODE == 5) {   // chain warp with LDS.128 feed: 4 rows x (2 LDS.128, 4 FMUL, 4 FADD) = 16 FADD
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 s = s4[((r * 4 + j) & 63) * 8], d = s4[512 + ((r * 4 + j) & 63) * 8];
        a0 = a0 + s.x * d.x; a0 = a0 + s.y * d.y; a0 = a0 + s.z * d.z; a0 = a0 + s.w * d.w;
      }
    } else if (MODE == 6) {   // FADD-only chain with LDS.128 feed: 4 rows x (1 LDS.128, 4 FADD)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 s = s4[((r * 4 + j) & 63) * 8];
        a0 = a0 + s.x; a0 = a0 + s.y; a0 = a0 + s.z; a0 = a0 + s.w;
      }
    } else if (MODE == 7) {   // two chains interleaved, FADD-only, LDS.128 feed: 2 x 8 FADD
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 s = s4[((r * 2 + j) & 63) * 8], d = s4[512 + ((r * 2 + j) & 63) * 8];
        a0 = a0 + s.x; a1 = a1 + d.x; a0 = a0 + s.y; a1 = a1 + d.y; a0 = a0 + s.z; a1 = a1 + d.z; a0 = a0 + s.w; a1 = a1 + d.w;
      }
    } else if (MODE == 8) {   // 16 independent FMUL
#pragma unroll
      for (int j = 0; j < 2; ++j) { a0 = a0 * b; a1 = a1 * c; a2 = a2 * b; a3 = a3 * c; a4 = a4 * b; a5 = a5 * c; a6 = a6 * b; a7 = a7 * c; }
    } else if (MODE == 9) {   // 16 independent FFMA (explicit)
#pragma unroll
      for (int j = 0; j < 2; ++j) { a0 = fmaf(a0, b, c); a1 = fmaf(a1, c, b); a2 = fmaf(a2, b, c); a3 = fmaf(a3, c, b); a4 = fmaf(a4, b, c); a5 = fmaf(a5, c, b); a6 = fmaf(a6, b, c); a7 = fmaf(a7, c, b); }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockDim.x * blockIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main(int argc, char** argv) {
  const int nw = argc > 1 ? atoi(argv[1]) : 1;   // warps per CTA (all on one SM)
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 16 * 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"16 dependent FADD", "2 chains x 8 FADD", "4 chains x 4 FADD", "8 chains x 2 FADD", "8 x (FMUL + dep FADD) + 8 FADD(other chain)",
                         "4 rows x (2 LDS.128 + 4 FMUL + 4 dep FADD)", "4 rows x (1 LDS.128 + 4 dep FADD)", "2 chains x 8 FADD, LDS.128 feed", "16 indep FMUL", "16 indep FFMA"};
#define RUN(M) k<M><<<1, 32 * nw>>>(out, in, cyc); cudaDeviceSynchronize(); printf("warps %d  %-48s %6.2f cycles per loop body (%lld total)\n", nw, names[M], (double)cyc[M] / REP, cyc[M]);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
