// chain_probe2.cu — how cheap can one row (4 dependent FADD of one chain per lane) get, depending on how its
// operands arrive?  One warp alone, 2 and 4 warps per SM sub-partition mix (blockDim 32..256).
//   mode 0: K2r as built: 8 rows of (sd quad, pdiff quad) loaded, then 32 FMUL + 32 dependent FADD
//   mode 1: the same loads, software-pipelined two rows ahead and interleaved with the additions
//   mode 2: products delivered: ONE LDS.128 per row, 4 dependent FADD (batch of 8 rows)
//   mode 3: products delivered, pipelined two rows ahead
//   mode 5: two chains per lane, three LDS.128 per row, two rows ahead; mode 6: the same, lanes 24..31 retired
//   mode 4: no loads at all: 4 FMUL + 4 dependent FADD per row from registers (the floor)
#include <cuda_runtime.h>
#include <stdio.h>
#define REP 64
__device__ __forceinline__ float4 lds128(const float4* p) {
  float4 v;
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__global__ void k(float* out, long long* cyc, int mode) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 24576 + 4096; i += blockDim.x) sm[i] = 1.0f + (float)(i % 97) * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31, c = lane & 7, kk = lane >> 3;
  const float4* sdk = reinterpret_cast<const float4*>(sm) + kk * 256 + c;
  const float4* pdb = reinterpret_cast<const float4*>(sm + 24576) + c;
  float acc = -0.0f;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      const float4* sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
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
  } else if (mode == 7 || mode == 8 || mode == 9) {
    // mode 0 plus the per-unit hand-offs of the kernel: 7 = __syncwarp + mbarrier.arrive by lane 0; 8 = + ld.acquire peek;
    // 9 = only the ld.acquire peek
    __shared__ unsigned long long bar[8];
    __shared__ unsigned flag[8];
    if (threadIdx.x < 8) { flag[threadIdx.x] = 1u;
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[threadIdx.x])), "r"(1 << 19)); }
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      const float4* sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
      const float4* pd4 = pdb + (u & 1) * 320;
      unsigned nf = 1u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 d[8], a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { d[r] = pd4[(8 * h + r) * 10]; a[r] = sd4[(8 * h + r) * 8]; }
        if (h == 1 && mode >= 8)
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(nf) : "r"((unsigned)__cvta_generic_to_shared(&flag[(u + 1) & 7])) : "memory");
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          acc = acc + a[r].x * d[r].x; acc = acc + a[r].y * d[r].y; acc = acc + a[r].z * d[r].z; acc = acc + a[r].w * d[r].w;
        }
      }
      if (mode <= 8) {
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&bar[u & 7])) : "memory");
      }
      while (nf != 1u)
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(nf) : "r"((unsigned)__cvta_generic_to_shared(&flag[(u + 1) & 7])) : "memory");
    }
  } else if (mode == 1) {
    // two rows ahead
    const float4* sd4 = sdk; const float4* pd4 = pdb;
    float4 a0 = lds128(sd4), d0 = lds128(pd4), a1 = lds128(sd4 + 8), d1 = lds128(pd4 + 10);
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
      pd4 = pdb + (u & 1) * 320;
#pragma unroll
      for (int r = 0; r < 16; r += 2) {
        const int rn = (r + 2) & 15;
        float p0 = a0.x * d0.x, p1 = a0.y * d0.y, p2 = a0.z * d0.z, p3 = a0.w * d0.w;
        acc = acc + p0;
        a0 = lds128(sd4 + rn * 8);
        acc = acc + p1;
        d0 = lds128(pd4 + rn * 10);
        acc = acc + p2; acc = acc + p3;
        p0 = a1.x * d1.x; p1 = a1.y * d1.y; p2 = a1.z * d1.z; p3 = a1.w * d1.w;
        acc = acc + p0;
        a1 = lds128(sd4 + (rn + 1) * 8);
        acc = acc + p1;
        d1 = lds128(pd4 + (rn + 1) * 10);
        acc = acc + p2; acc = acc + p3;
      }
    }
  } else if (mode == 2) {
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      const float4* sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = sd4[(8 * h + r) * 8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc = acc + a[r].x; acc = acc + a[r].y; acc = acc + a[r].z; acc = acc + a[r].w; }
      }
    }
  } else if (mode == 3) {
    const float4* sd4 = sdk;
    float4 a0 = lds128(sd4), a1 = lds128(sd4 + 8);
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      sd4 = sdk + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
#pragma unroll
      for (int r = 0; r < 16; r += 2) {
        const int rn = (r + 2) & 15;
        float4 b = a0;
        acc = acc + b.x;
        a0 = lds128(sd4 + rn * 8);
        acc = acc + b.y; acc = acc + b.z; acc = acc + b.w;
        b = a1;
        acc = acc + b.x;
        a1 = lds128(sd4 + (rn + 1) * 8);
        acc = acc + b.y; acc = acc + b.z; acc = acc + b.w;
      }
    }
  } else if (mode == 5 || mode == 6) {
    // two chains per lane (planes kp, kp + 1 of the same columns): three loads per row, rolling two rows ahead
    const int kp = min(kk, 2);
    const float4* sdp = reinterpret_cast<const float4*>(sm) + 2 * kp * 256 + c;
    float acc1 = -0.0f;
    if (mode == 6 && lane >= 24) { out[threadIdx.x] = 0.f; return; }
    const float4* sd4 = sdp; const float4* pd4 = pdb;
    float4 a0 = lds128(sd4), b0 = lds128(sd4 + 256), d0 = lds128(pd4);
    float4 a1 = lds128(sd4 + 8), b1 = lds128(sd4 + 256 + 8), d1 = lds128(pd4 + 10);
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
      sd4 = sdp + ((u >> 1) & 3) * 1536 + (u & 1) * 128;
      pd4 = pdb + (u & 1) * 320;
#pragma unroll
      for (int r = 0; r < 16; r += 2) {
        const int rn = (r + 2) & 15;
        float p0 = a0.x * d0.x, p1 = a0.y * d0.y, p2 = a0.z * d0.z, p3 = a0.w * d0.w;
        float q0 = b0.x * d0.x, q1 = b0.y * d0.y, q2 = b0.z * d0.z, q3 = b0.w * d0.w;
        acc = acc + p0; acc1 = acc1 + q0;
        a0 = lds128(sd4 + rn * 8);
        acc = acc + p1; acc1 = acc1 + q1;
        b0 = lds128(sd4 + 256 + rn * 8);
        acc = acc + p2; acc1 = acc1 + q2;
        d0 = lds128(pd4 + rn * 10);
        acc = acc + p3; acc1 = acc1 + q3;
        p0 = a1.x * d1.x; p1 = a1.y * d1.y; p2 = a1.z * d1.z; p3 = a1.w * d1.w;
        q0 = b1.x * d1.x; q1 = b1.y * d1.y; q2 = b1.z * d1.z; q3 = b1.w * d1.w;
        acc = acc + p0; acc1 = acc1 + q0;
        a1 = lds128(sd4 + (rn + 1) * 8);
        acc = acc + p1; acc1 = acc1 + q1;
        b1 = lds128(sd4 + 256 + (rn + 1) * 8);
        acc = acc + p2; acc1 = acc1 + q2;
        d1 = lds128(pd4 + (rn + 1) * 10);
        acc = acc + p3; acc1 = acc1 + q3;
      }
    }
    acc += acc1;
  } else {
    float4 a = sdk[0], d = pdb[0];
#pragma unroll 1
    for (int u = 0; u < REP; ++u) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        acc = acc + a.x * d.x; acc = acc + a.y * d.y; acc = acc + a.z * d.z; acc = acc + a.w * d.w;
        a.x += 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMallocManaged(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000);
  for (int mode = 0; mode < 10; ++mode)
    for (int nw = 1; nw <= 8; nw *= 2) {
      k<<<1, 32 * nw, 116000>>>(out, cyc, mode); cudaDeviceSynchronize();
      printf("mode %d warps %d: %6.1f cycles per row\n", mode, nw, (double)cyc[0] / REP / 16);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
