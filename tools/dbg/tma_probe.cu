// Standalone probe: which TMA box/type combinations work?  usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../modelcompression_b200/csrc/ptx_sm100.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe4d(const __grid_constant__ CUtensorMap tm, float* out, int nfloat, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, nfloat * 4);
    ptx::tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
  }
  unsigned int spins = 0;
  while (!ptx::mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { if (threadIdx.x == 0) out[0] = -12345.f; return; } }
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
__global__ void probe3d(const __grid_constant__ CUtensorMap tm, float* out, int nfloat, int c0, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, nfloat * 4);
    ptx::tma_load_3d(smem, &tm, &bar, c0, c1, c2);
  }
  unsigned int spins = 0;
  while (!ptx::mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { if (threadIdx.x == 0) out[0] = -12345.f; return; } }
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int B = 2, C = 3, H = 416, W = 416;
  std::vector<float> h((size_t)B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 1 << 20);
  CUtensorMap tm;
  CUresult r;
  int nfloat = 0;
  if (variant <= 3) {
    // 4-D fp32 [B][C][H][W]
    cuuint32_t bw = variant == 0 ? 36 : variant == 1 ? 32 : variant == 2 ? 64 : 36;
    cuuint32_t bc = variant == 3 ? 1 : 3;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {bw, 18, bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode rc=%d box=%u,18,%u,1\n", variant, (int)r, bw, bc);
    nfloat = bw * 18 * bc;
    probe4d<<<1, 128, nfloat * 4 + 1024>>>(tm, out, nfloat, -1, -1, 0, 1);
  } else {
    // 3-D fp32 view [B*C*H][W] ... as dims (W, H, B*C)
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
    cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {36, 18, 3};
    cuuint32_t es[3] = {1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d (3-D) encode rc=%d\n", variant, (int)r);
    nfloat = 36 * 18 * 3;
    probe3d<<<1, 128, nfloat * 4 + 1024>>>(tm, out, nfloat, -1, -1, 3);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(nfloat);
    cudaMemcpy(o.data(), out, nfloat * 4, cudaMemcpyDeviceToHost);
    printf("out[0..5]= %g %g %g %g %g %g ; out[37]=%g (expect row0 zeros, then %g)\n", o[0], o[1], o[2], o[3], o[4], o[5],
           o[37], h[(size_t)1 * C * H * W + 0]);
  }
  return 0;
}
