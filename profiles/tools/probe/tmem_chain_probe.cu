// tmem_chain_probe.cu — can the steepest-descent operand of a reference-order chain come from TENSOR MEMORY instead
// of shared memory?  One chain per lane: per row 4 dependent FADD, 4 FMUL, the sd quad and the pdiff quad.
//   mode 0: sd by LDS.128, pdiff by LDS.128 (K2r as built)
//   mode 1: sd by tcgen05.ld.32x32b.x4 from the lane's own TMEM row (column = element index), pdiff by LDS.128,
//           the load of row r + 1 issued before the additions of row r, tcgen05.wait::ld after them
//   mode 2: like 1, eight rows of sd per tcgen05.ld (.x32), pdiff by LDS.128
// NW chain warps per CTA (warp w uses TMEM lanes 32 (w % 4) ..), plus optional 4 "producer-like" warps.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define REP 64
#define NCOLS 256
__device__ __forceinline__ void tm_ld4(uint32_t taddr, float4& v) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
  v.x = __uint_as_float(a); v.y = __uint_as_float(b); v.z = __uint_as_float(c); v.w = __uint_as_float(d);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st4(uint32_t taddr, float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
               "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)));
}
__global__ void k(float* out, long long* cyc, int mode, int nchain, int nprod) {
  extern __shared__ __align__(16) float sm[];
  __shared__ uint32_t s_taddr;
  for (int i = threadIdx.x; i < 24576 + 4096; i += blockDim.x) sm[i] = 1.0f + (float)(i % 97) * 1e-3f;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = lane & 7, kk = lane >> 3;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_taddr)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_taddr + ((uint32_t)(32 * (w & 3)) << 16);
  const float4* sdk = reinterpret_cast<const float4*>(sm) + kk * 256 + c;
  const float4* pdb = reinterpret_cast<const float4*>(sm + 24576) + c;
  if (w < nchain && w < 4) {   // fill this warp's TMEM rows: column 4 r' + i <- the quad the LDS form would read for row r'
    for (int r = 0; r < NCOLS / 4; ++r) {
      const int uu = r >> 4;
      tm_st4(tbase + 4 * r, sdk[((uu >> 1) & 3) * 1536 + (uu & 1) * 128 + (r & 15) * 8]);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  float acc = -0.0f;
  long long t0 = clock64();
  if (w < nchain) {
    if (mode == 0) {
#pragma unroll 1
      for (int u = 0; u < REP; ++u) {
        const int uu = u & 3;
        const float4* sd4 = sdk + ((uu >> 1) & 3) * 1536 + (uu & 1) * 128;
        const float4* pd4 = pdb + (u & 1) * 320;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 d[8], a[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { d[r] = pd4[(8 * h + r) * 10]; a[r] = sd4[(8 * h + r) * 8]; }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            acc = acc + a[r].x * d[r].x; acc = acc + a[r].y * d[r].y; acc = acc + a[r].z * d[r].z; acc = acc + a[r].w * d[r].w;
          }
        }
      }
    } else if (mode == 1) {
      float4 a, an, d, dn;
      tm_ld4(tbase, a);
      d = pdb[0];
      tm_wait_ld();
#pragma unroll 1
      for (int u = 0; u < REP; ++u) {
        const float4* pd4 = pdb + (u & 1) * 320;
        const uint32_t tu = tbase + (u & 3) * 64;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int rn = (r + 1) & 15;
          tm_ld4(tu + 4 * rn, an);
          dn = pd4[rn * 10];
          acc = acc + a.x * d.x; acc = acc + a.y * d.y; acc = acc + a.z * d.z; acc = acc + a.w * d.w;
          tm_wait_ld();
          a = an; d = dn;
        }
      }
    } else {
      // eight rows of sd (32 columns) per tcgen05.ld
#pragma unroll 1
      for (int u = 0; u < REP; ++u) {
        const float4* pd4 = pdb + (u & 1) * 320;
        const uint32_t tu = tbase + (u & 3) * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t s[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
              "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]),
                "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(s[16]), "=r"(s[17]), "=r"(s[18]),
                "=r"(s[19]), "=r"(s[20]), "=r"(s[21]), "=r"(s[22]), "=r"(s[23]), "=r"(s[24]), "=r"(s[25]), "=r"(s[26]), "=r"(s[27]),
                "=r"(s[28]), "=r"(s[29]), "=r"(s[30]), "=r"(s[31])
              : "r"(tu + 32 * h));
          float4 d[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) d[r] = pd4[(8 * h + r) * 10];
          tm_wait_ld();
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            acc = acc + __uint_as_float(s[4 * r]) * d[r].x; acc = acc + __uint_as_float(s[4 * r + 1]) * d[r].y;
            acc = acc + __uint_as_float(s[4 * r + 2]) * d[r].z; acc = acc + __uint_as_float(s[4 * r + 3]) * d[r].w;
          }
        }
      }
    }
  } else if (w < nchain + nprod) {
    // producer-like shared-memory traffic (2 LDS.32 + 1 STS.32 + 8 FP per row)
    float* win = sm + 24576 + 1024 * (w & 1);
    float a0 = win[lane + 1], b0 = win[lane], r = 0.f;
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a1 = win[((j + 1) & 15) * 40 + lane + 1], b1 = win[((j + 1) & 15) * 40 + lane];
        const float pn = ((0.3f * a1 + 0.2f * b1) + 0.4f * a0) + 0.1f * b0;
        win[j * 40 + 4 * (lane & 7) + (lane >> 3) + 512] = r - pn;
        a0 = a1; b0 = b1; r = pn;
      }
    }
    acc = r;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_taddr), "n"(NCOLS));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMallocManaged(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000);
  float ref = 0.f;
  for (int mode = 0; mode < 3; ++mode)
    for (int nchain = 1; nchain <= 4; nchain *= 2)
      for (int nprod = 0; nprod <= 4; nprod += 4) {
        k<<<1, 32 * (nchain + nprod), 116000>>>(out, cyc, mode, nchain, nprod);
        cudaError_t e = cudaDeviceSynchronize();
        float h = 0.f; cudaMemcpy(&h, out + 5, 4, cudaMemcpyDeviceToHost);
        if (mode == 0 && nchain == 1 && nprod == 0) ref = h;
        printf("mode %d chain warps %d producer warps %d: %6.1f cycles per row   lane-5 sum %s (%g)  %s\n", mode, nchain, nprod,
               (double)cyc[0] / REP / 16, h == ref ? "equal" : "DIFFERENT", h, cudaGetErrorString(e));
      }
}
