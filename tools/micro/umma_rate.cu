// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, cta_group::1, both operands in shared
// memory, K-major SWIZZLE_128B) as a function of N and of how often tcgen05.commit is interleaved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vid2vid_b200/csrc -o gpurun_out/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace v2v;

// layout: 2 = SWIZZLE_128B (4 MMAs per 128-byte row), 4 = SWIZZLE_64B (2 per 64-byte row), 6 = SWIZZLE_32B (1 per row)
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int commit_every, int kstep_bytes, int layout, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 1) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a = smem_u32(sm), b = smem_u32(sm + 32768);
    t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        const int rowb = layout == 2 ? 128 : (layout == 4 ? 64 : 32), per_row = rowb / 32;
        const uint64_t ad = make_kmajor_desc(a + (it & 1) * 16384, 8 * rowb, layout), bd = make_kmajor_desc(b + (it & 1) * 32768 / 2, 8 * rowb, layout);
        // 4 MMAs per iteration: per_row K sub-steps inside a row, then the next 128-row operand block
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
          const uint64_t off = (uint64_t)(((kq % per_row) * kstep_bytes + (kq / per_row) * 128 * rowb) >> 4);
          umma_bf16(tm, ad + off, bd + off, idesc, 1u);
        }
        if (commit_every && (it % commit_every) == commit_every - 1) umma_commit(&bar[1]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], ph);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int layout : {2, 4, 6}) for (int ce : {0, 1}) for (int ks : {32, 0}) for (int N : {16, 32, 64, 128, 256}) {
    if (layout != 2 && (ce == 1 || ks == 0)) continue;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 128, 100 * 1024>>>(N, 10, ce, ks, layout, d);     // warm
    cudaEventRecord(e0);
    k<<<148, 128, 100 * 1024>>>(N, iters, ce, ks, layout, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = 4.0 * iters;
    printf("layout=%d N=%3d commit_every=%d kstep=%2dB: %7.1f cycles/MMA  (%6.1f ns/MMA by events)  -> %.0f TFLOP/s chip  %s\n", layout, N, ce, ks, cyc / mmas,
           ms * 1e6 / mmas, 148 * mmas * 2.0 * 128 * N * 16 / (ms * 1e-3) / 1e12, cudaGetErrorString(e));
  }
  return 0;
}
