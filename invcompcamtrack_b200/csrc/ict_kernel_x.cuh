// ict_kernel_x.cuh — pieces shared by the reference-order (Eigen packet sum) kernels K2x (ict_kernel_x.cu, 32x32
// patches) and K2x8 (ict_kernel_x8.cu, 8x8 patches): producer/chain-warp roles, the steepest-descent values and Hessian
// products with the reference's roundings, Eigen's redux tail, the ring hand-off wait.
#pragma once
#include "ict_device.cuh"

namespace ict {

#define KX_PROD 7                       /* producer warps; warp KX_PROD is the chain warp */

// Producer-side wait: poll with a back-off.  A producer is a round ahead of the chain warp most of the time; a tight
// try_wait loop (SYNCS + YIELD + BRA) was 29 % of all issued instructions (ncu).  Backing off costs nothing (the
// chain warp is the critical resource) and leaves the issue slots to it.
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* b, unsigned parity) {
  unsigned done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(ict_saddr(b)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
  }
}

struct __align__(16) KxShared {
  unsigned long long full[2];             // ring half written: all lanes of all producer warps arrive
  unsigned long long empty[2];            // ring half read: the 32 chain lanes arrive
  float G[12];
  float p[8];
  float sum[8];
  float dp[8];
  float Hsum[24];
  Lu6 f;
  int cont, it, nv;
};

// pair (a <= b) of the q-th Hessian entry in the order of ComputeHessian (odometer.cpp:430-455)
__host__ __device__ constexpr int kx_pair_a(int q) { return q < 6 ? 0 : q < 11 ? 1 : q < 15 ? 2 : q < 18 ? 3 : q < 20 ? 4 : 5; }
__host__ __device__ constexpr int kx_pair_b(int q) {
  return q < 6 ? q : q < 11 ? q - 5 : q < 15 ? q - 9 : q < 18 ? q - 12 : q < 20 ? q - 14 : 5;
}

// the six steepest-descent values of one pixel with the reference's roundings (odometer.cpp:317-326);
// ab = {A_0..A_5, B_0..B_5} of the pixel's point (A_1 = B_0 = 0: sd1 and sd2 have one term)
__device__ __forceinline__ void kx_sd(float gx, float gy, const float* ab, float* sd) {
  sd[0] = gx * ab[0];
  sd[1] = gy * ab[7];
  sd[2] = gx * ab[2] + gy * ab[8];
  sd[3] = gx * ab[3] + gy * ab[9];
  sd[4] = gx * ab[4] + gy * ab[10];
  sd[5] = gx * ab[5] + gy * ab[11];
}

template <int PASS>
__device__ __forceinline__ void kx_hess_products(const float* sd, float* out) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    constexpr int dummy = 0;
    (void)dummy;
    const int q = 6 * PASS + k;
    out[k] = q < 21 ? sd[kx_pair_a(q < 21 ? q : 0)] * sd[kx_pair_b(q < 21 ? q : 0)] : 0.0f;
  }
}

// Eigen's redux tail for chains held one per lane in groups of eight: (c0+c4 + c2+c6) + (c1+c5 + c3+c7).
// Returns the sum in the lanes with (lane & 7) == 0.
__device__ __forceinline__ float kx_finish(float s) {
  const unsigned FULL = 0xffffffffu;
  const float p0 = s + __shfl_down_sync(FULL, s, 4);        // lanes c < 4: ch[c] + ch[c+4]
  const float t = p0 + __shfl_down_sync(FULL, p0, 2);       // lane c = 0: p0[0] + p0[2]; c = 1: p0[1] + p0[3]
  return t + __shfl_down_sync(FULL, t, 1);
}

}  // namespace ict
