// ict_kernel_v2.cuh — device helpers shared by the float4-row-quad kernels for 32x32 patches:
// K2v2 (ict_kernel_v2.cu, production) and K2x (ict_kernel_x.cu, reference-order sums).
#pragma once
#include "ict_device.cuh"

namespace ict {

__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }

// project_pt (pose.cpp:307-397) of one point at level intrinsics (fx, fy, cx, cy) + util_getPatch placement
// (utilities.cpp:65-94): the reference's operation order, no fused operations.  Writes {base, vis, -, -}, {w0..w3}.
__device__ __forceinline__ int place_point(const float* G, float X, float Y, float Z, float fx, float fy, float cx,
                                           float cy, float swo, float sho, int width, float4* dst) {
  const float tx = G[0] * X + G[1] * Y + G[2] * Z + G[3];
  const float ty = G[4] * X + G[5] * Y + G[6] * Z + G[7];
  const float tz = G[8] * X + G[9] * Y + G[10] * Z + G[11];
  const float mx = (tx / tz) * fx + cx, my = (ty / tz) * fy + cy;
  const int vis = (mx >= 0) & (my >= 0) & (mx <= swo) & (my <= sho);   // odometer.cpp:369-371 (NaN -> outside)
  PatchPlace pl = {0, 0.f, 0.f, 0.f, 0.f};
  if (vis) pl = patch_place(mx, my, 16, width);
  dst[0] = make_float4(__int_as_float(pl.base), __int_as_float(vis), 0.0f, 0.0f);
  dst[1] = make_float4(pl.w0, pl.w1, pl.w2, pl.w3);
  return vis;
}

// util_getPatch_grad (utilities.cpp:160-185) for KT consecutive rows of one patch column: p points at the row ABOVE
// the first one (the bilinear sample of row r reads rows r and r-1, columns c and c-1); unfused, in the reference's
// order ((w0*a + w1*b) + w2*c) + w3*d.  Writes KT/4 float4 row-quads at dst, dst + 32, ...
template <int KT>
__device__ __forceinline__ void gather_plane(const float* __restrict__ p, int width, const float4 w, float4* dst) {
  float a[KT + 1], b[KT + 1];
#pragma unroll
  for (int j = 0; j <= KT; ++j) {
    a[j] = __ldg(p + j * width);
    b[j] = __ldg(p + j * width - 1);
  }
#pragma unroll
  for (int jq = 0; jq < KT / 4; ++jq) {
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = 4 * jq + j + 1;
      r[j] = ((w.x * a[row] + w.y * b[row]) + w.z * a[row - 1]) + w.w * b[row - 1];
    }
    dst[jq * 32] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

}  // namespace ict
