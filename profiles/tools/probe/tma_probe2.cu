// tma_probe2.cu — the CUDA programming guide's libcu++ TMA example, box BW x BH floats, as a known-good reference.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#ifndef BW
#define BW 36
#endif
#ifndef BH
#define BH 17
#endif
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float* out) {
  __shared__ alignas(128) float smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}

int main(int argc, char** argv) {
  const int W = 304, H = 199, x = 37, y = 51;
  std::vector<float> h((size_t)W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, sizeof(float) * W * H);
  cudaMalloc(&out, sizeof(float) * BW * BH);
  cudaMemcpy(d, h.data(), sizeof(float) * W * H, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 2; }
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (box %d x %d)\n", (int)r, BW, BH);
  if (r) return 3;
  kernel<<<1, 64>>>(tm, x, y, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("libcu++ example: %s\n", cudaGetErrorString(e));
  if (e) return 1;
  std::vector<float> o(BW * BH);
  cudaMemcpy(o.data(), out, sizeof(float) * BW * BH, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; ++r2)
    for (int c = 0; c < BW; ++c) bad += o[r2 * BW + c] != (float)((y + r2) * W + x + c);
  printf("%d mismatches\n", bad);
  return bad != 0;
}
