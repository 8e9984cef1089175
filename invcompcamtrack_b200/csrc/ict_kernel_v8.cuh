// ict_kernel_v8.cuh — the lane layout of an 8x8 patch shared by K2v8 (ict_kernel_v8.cu) and K2x8 (ict_kernel_x8.cu):
// lane l takes the pixels (2j, c) and (2j+1, c), j = l/8, c = l%8.
#pragma once
#include "ict_device.cuh"

namespace ict {

// the three rows a lane needs of one plane: rows 2j-1, 2j, 2j+1 at columns c and c-1 (o = offset of row 2j-1, col c)
struct V8Rows { float a0, b0, a1, b1, a2, b2; };
__device__ __forceinline__ V8Rows v8_load(const float* __restrict__ pl, int o, int width) {
  V8Rows r;
  r.a0 = __ldg(pl + o);             r.b0 = __ldg(pl + o - 1);
  r.a1 = __ldg(pl + o + width);     r.b1 = __ldg(pl + o + width - 1);
  r.a2 = __ldg(pl + o + 2 * width); r.b2 = __ldg(pl + o + 2 * width - 1);
  return r;
}
// util_getPatch_grad (utilities.cpp:160-185), unfused, reference order: the lane's two pixels
__device__ __forceinline__ float2 v8_bilin_exact(const V8Rows& r, const float4 w) {
  float2 v;
  v.x = ((w.x * r.a1 + w.y * r.b1) + w.z * r.a0) + w.w * r.b0;
  v.y = ((w.x * r.a2 + w.y * r.b2) + w.z * r.a1) + w.w * r.b1;
  return v;
}
// util_getPatch (utilities.cpp:107) with fused multiply-adds in the same association order
__device__ __forceinline__ float2 v8_bilin_fma(const V8Rows& r, const float4 w) {
  float2 v;
  v.x = fmaf(w.w, r.b0, fmaf(w.z, r.a0, fmaf(w.y, r.b1, w.x * r.a1)));
  v.y = fmaf(w.w, r.b1, fmaf(w.z, r.a1, fmaf(w.y, r.b2, w.x * r.a2)));
  return v;
}
__device__ __forceinline__ float v8_warp_total(float v) {   // butterfly: every lane gets the total, fixed order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ict
