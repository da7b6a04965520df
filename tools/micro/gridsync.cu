// microbenchmark: cost of cooperative grid.sync() on B200 in the shapes weight_prune_kernel uses
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256) k_sync(int reps, unsigned long long* out) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = 0, t1 = 0;
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int i = 0; i < reps; ++i) grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    out[0] = t1 - t0;
  }
}

// each block streams `words` floats (read a, write b) then syncs; reports loop end and sync end per block 0
__global__ void __launch_bounds__(256) k_stream_sync(const float4* a, float4* b, long long n4, unsigned long long* out, int do_write) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = 0, t1 = 0, t2 = 0;
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (long long i = (long long)blockIdx.x * 1024 + threadIdx.x; i < n4; i += (long long)gridDim.x * 1024) {
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = (i + u * 256 < n4) ? __ldcs(a + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 m;
      m.x = fabsf(q[u].x) > 0.5f ? 1.f : 0.f; m.y = fabsf(q[u].y) > 0.5f ? 1.f : 0.f;
      m.z = fabsf(q[u].z) > 0.5f ? 1.f : 0.f; m.w = fabsf(q[u].w) > 0.5f ? 1.f : 0.f;
      if (do_write && i + u * 256 < n4) __stcs(b + i + u * 256, m);
      else if (m.x == 2.f) b[0] = m;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t2));
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
}

// same-line global atomics: every block adds to `nbins` consecutive counters
__global__ void __launch_bounds__(256) k_atomics(unsigned int* hist, int nbins, int stride, unsigned long long* out) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = 0, t1 = 0;
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int i = threadIdx.x; i < nbins; i += 256) atomicAdd(&hist[i * stride], 1u);
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    out[0] = t1 - t0;
  }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 64);
  unsigned long long h[2];
  const long long n = 50634592;
  float4 *a, *b;
  cudaMalloc(&a, n * 4);
  cudaMalloc(&b, n * 4);
  cudaMemset(a, 0, n * 4);
  unsigned int* hist;
  cudaMalloc(&hist, 1024 * 32 * 4);
  cudaMemset(hist, 0, 1024 * 32 * 4);
  for (int bps = 1; bps <= 5; ++bps) {
    int grid = 148 * bps;
    int reps = 100;
    void* args[] = {&reps, &d_out};
    for (int it = 0; it < 2; ++it) cudaLaunchCooperativeKernel((void*)k_sync, dim3(grid), dim3(256), args, 0, 0);
    cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("grid %d: grid.sync = %.2f us each (%s)\n", grid, h[0] / 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
    long long n4 = n / 4;
    for (int w = 0; w < 2; ++w) {
      void* a2[] = {&a, &b, &n4, &d_out, &w};
      for (int it = 0; it < 2; ++it) cudaLaunchCooperativeKernel((void*)k_stream_sync, dim3(grid), dim3(256), a2, 0, 0);
      cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
      printf("grid %d: stream %s: block0 loop end %.1f us, after grid.sync %.1f us (%s)\n", grid, w ? "read+write" : "read only",
             h[0] / 1e3, h[1] / 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    for (int stride = 1; stride <= 32; stride *= 32) {
      int nbins = 1024;
      void* a3[] = {&hist, &nbins, &stride, &d_out};
      for (int it = 0; it < 2; ++it) cudaLaunchCooperativeKernel((void*)k_atomics, dim3(grid), dim3(256), a3, 0, 0);
      cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
      printf("grid %d: %d x 1024 atomics, bin stride %d words: %.1f us incl. one sync\n", grid, grid, stride, h[0] / 1e3);
    }
  }
  return 0;
}
