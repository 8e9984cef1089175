// tma_probe.cu — which way of handing a 2-D tensor map to cp.async.bulk.tensor works on this B200 / driver:
//   mode 0: __grid_constant__ kernel parameter   mode 1: global memory, no fence   mode 2: global memory + acquire fence
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe <mode>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ void body(const void* tm, int x, int y, float* out, int fence) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ unsigned long long bar;
  float* win = reinterpret_cast<float*>(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(tm) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar)), "r"(36 * 17 * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            saddr(win)),
        "l"(tm), "r"(x), "r"(y), "r"(saddr(&bar))
        : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(done)
        : "r"(saddr(&bar)), "r"(0)
        : "memory");
  }
  for (int i = threadIdx.x; i < 36 * 17; i += blockDim.x) out[i] = win[i];
}

__global__ void k_param(const __grid_constant__ CUtensorMap tm, int x, int y, float* out) { body(&tm, x, y, out, 0); }
__global__ void k_global(const CUtensorMap* tm, int x, int y, float* out, int fence) { body(tm, x, y, out, fence); }

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 304, H = 199, x = 37, y = 51;
  std::vector<float> h((size_t)W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, sizeof(float) * W * H);
  cudaMalloc(&out, sizeof(float) * 36 * 17);
  cudaMemcpy(d, h.data(), sizeof(float) * W * H, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 2; }
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {36, 17}, es[2] = {1, 1};
  CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  if (r) return 3;
  if (mode == 0) {
    k_param<<<1, 64, 4096>>>(tm, x, y, out);
  } else {
    CUtensorMap* dtm;
    cudaMalloc(&dtm, 4 * sizeof(CUtensorMap));
    cudaMemcpy(dtm + 1, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    k_global<<<1, 64, 4096>>>(dtm + 1, x, y, out, mode == 2);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d: %s\n", mode, cudaGetErrorString(e));
  if (e) return 1;
  std::vector<float> o(36 * 17);
  cudaMemcpy(o.data(), out, sizeof(float) * 36 * 17, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < 17; ++r2)
    for (int c = 0; c < 36; ++c) bad += o[r2 * 36 + c] != (float)((y + r2) * W + x + c);
  printf("mode %d: %d mismatches\n", mode, bad);
  return bad != 0;
}
