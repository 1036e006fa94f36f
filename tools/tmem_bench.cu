// Microbenchmark: TMEM read-out (tcgen05.ld 32x32b.x32) throughput of one SM, alone and while the tensor
// pipe is accumulating into the other half of TMEM (tcgen05.mma M=128 N=256 K=16, bf16, operands in smem).
// Answers: is the epilogue of a K = 256 GEMM tile (128 KB of fp32 accumulators) limited by the TMEM read
// port, and does it contend with the MMAs of the next tile?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I3d_vit_ensemble_b200/csrc -o tmem_bench tools/tmem_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace vit3d::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// warps: 0 idle, 1 = MMA issuer, 2..17 = readers (reader r uses lane quarter (warp & 3), column slice (r >> 2))
__global__ void __launch_bounds__(576, 1) tmem_kernel(int readers, int do_mma, int reps, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar, bar_done, bar_slot[4];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&bar_done, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_slot[i], 1);
    fence_barrier_init(); stop = 0;
    mbar_arrive(&bar_done);          // phase 0 of bar_done is complete from now on
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  long long t0 = 0, t1 = 0;
  uint32_t acc = 0;
  if (warp == 1) {
    if (do_mma && lane == 0) {
      const uint32_t idesc = make_idesc(UMMA_FMT_BF16, 128, 256, 0, 0);
      const uint32_t sa = smem_u32(smem), sb = sa + 16384;
      long long n = 0;
      t0 = clock64();
      while (!stop) {
        // one K = 256 tile = 16 MMAs into columns 256..511, then commit + wait (like the GEMM main loop)
        for (int k = 0; k < 64; ++k) {      // four K = 256 tiles per commit (keeps the commit round trip small)
          if (do_mma >= 2 && (k & 3) == 0) {          // per-k-block bookkeeping of a real main loop
            mbar_wait(&bar_done, 0);                  // an already-completed barrier (operand landed)
            tc_fence_after();
          }
          const uint64_t ad = make_smem_desc(sa + (k & 3) * 32, 16, 1024, UMMA_LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(sb + (k & 3) * 32, 16, 1024, UMMA_LAYOUT_SW128);
          umma<false>(tmem_base + 256, ad, bd, idesc, (k & 15) > 0 ? 1u : 0u);
          if (do_mma >= 3 && (k & 3) == 3) umma_commit(&bar_slot[(k >> 2) & 3]);   // release the ring slot
        }
        umma_commit(&bar);
        mbar_wait(&bar, (uint32_t)(n & 1));
        n += 1;
      }
      t1 = clock64();
      out[2] = t1 - t0;
      out[3] = n;
    }
  } else if (warp >= 2 && warp - 2 < readers) {
    const int r = warp - 2, q = warp & 3;
    const int slices = readers / 4 > 0 ? readers / 4 : 1;       // column slices per quarter
    const int cw = 256 / slices;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (r >> 2) * cw;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      for (int c = 0; c < cw; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
    }
    t1 = clock64();
    if (r == 0 && lane == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  }
  // readers done -> stop the MMA loop
  if (warp == 0 || (warp >= 2 && warp - 2 < readers))
    asm volatile("bar.sync 1, %0;" ::"r"(32 * (readers + 1)) : "memory");   // readers + warp 0
  if (threadIdx.x == 0) stop = 1;
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int main() {
  long long* out; uint32_t* sink;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 4096));
  CK(cudaFuncSetAttribute(tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int reps = 2000;
  for (int do_mma = 0; do_mma <= 3; ++do_mma)
    for (int readers : {16}) {
      CK(cudaMemset(out, 0, 64));
      tmem_kernel<<<1, 576, 65536>>>(readers, do_mma, reps, out, sink);
      CK(cudaDeviceSynchronize());
      long long h[4];
      CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
      const double clk_per_tile = (double)h[0] / reps;          // one pass over 128 x 256 fp32 by all readers
      printf("mma=%d readers=%2d: %.0f clk per 128x256 fp32 read-out (%.1f B/clk/SM)", do_mma, readers, clk_per_tile, 131072.0 / clk_per_tile);
      if (do_mma) printf("   concurrent MMA (%s): %.0f clk per K=256 tile (ideal 2048)", do_mma == 1 ? "bare" : (do_mma == 2 ? "+wait/k-block" : "+wait+commit/k-block"), (double)h[2] / (4.0 * (double)h[3]));
      printf("\n");
    }
  return 0;
}
