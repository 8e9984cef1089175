// tma_probe3.cu — raw-PTX 2-D TMA load with run-time box / swizzle / L2 promotion; prints the descriptor words.
//   ./tma_probe3 bw bh swizzle(0..3) l2promo(0..3) [direct: 1 = call cuTensorMapEncodeTiled through libcuda linked directly]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_param(const __grid_constant__ CUtensorMap tm, int x, int y, float* out, int nfl) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ unsigned long long bar;
  float* win = reinterpret_cast<float*>(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar)), "r"(nfl * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            saddr(win)),
        "l"(&tm), "r"(x), "r"(y), "r"(saddr(&bar))
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
  for (int i = threadIdx.x; i < nfl; i += blockDim.x) out[i] = win[i];
}

int main(int argc, char** argv) {
  const int bw = argc > 1 ? atoi(argv[1]) : 36, bh = argc > 2 ? atoi(argv[2]) : 17;
  const int swz = argc > 3 ? atoi(argv[3]) : 0, l2p = argc > 4 ? atoi(argv[4]) : 0, direct = argc > 5 ? atoi(argv[5]) : 0;
  const int W = 304, H = 199, x = argc > 6 ? atoi(argv[6]) : 36, y = argc > 7 ? atoi(argv[7]) : 51;
  std::vector<float> h((size_t)W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, sizeof(float) * W * H);
  cudaMalloc(&out, sizeof(float) * bw * bh);
  cudaMemcpy(d, h.data(), sizeof(float) * W * H, cudaMemcpyHostToDevice);
  int drv = 0, rt = 0;
  cudaDriverGetVersion(&drv);
  cudaRuntimeGetVersion(&rt);
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("driver %d runtime %d device %s cc %d.%d\n", drv, rt, pr.name, pr.major, pr.minor);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r;
  if (direct) {
    r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               (CUtensorMapSwizzle)swz, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 2; }
    r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (CUtensorMapSwizzle)swz, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("encode: %d (box %d x %d swizzle %d l2promo %d direct %d) base %p\n", (int)r, bw, bh, swz, l2p, direct, (void*)d);
  const unsigned long long* wd = (const unsigned long long*)&tm;
  for (int i = 0; i < 16; ++i) printf("%016llx%s", wd[i], (i & 3) == 3 ? "\n" : " ");
  if (r) return 3;
  cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k_param<<<1, 64, 32768>>>(tm, x, y, out, bw * bh);
  cudaError_t e = cudaDeviceSynchronize();
  printf("result: %s\n", cudaGetErrorString(e));
  if (e) return 1;
  std::vector<float> o(bw * bh);
  cudaMemcpy(o.data(), out, sizeof(float) * bw * bh, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2)
    for (int c = 0; c < bw; ++c) bad += o[r2 * bw + c] != (float)((y + r2) * W + x + c);
  printf("%d mismatches (swizzled layouts are expected to mismatch)\n", bad);
  return 0;
}
