// Microbenchmark: issue rate of tcgen05.mma (M=128 N=256 K=16, bf16, SS mode, SW128) for the four operand-major
// combinations.  One thread issues back-to-back MMAs on resident shared-memory tiles; ideal = 128 clk per MMA.
// The weight-gradient GEMMs (dW = dY^T X, reduction over token rows) read BOTH operands MN-major.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I3d_vit_ensemble_b200/csrc -Iinclude -o mma_major_bench tools/mma_major_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace vit3d::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int SMEM = 4 * 49152;

__global__ void __launch_bounds__(64, 1) major_kernel(int a_mn, int b_mn, int n_cols, int nmma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SMEM / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(UMMA_FMT_BF16, 128, n_cols, a_mn, b_mn);
      const uint64_t a0 = make_smem_desc(smem_u32(smem), a_mn ? 8192u : 16u, 1024, UMMA_LAYOUT_SW128);
      const uint64_t b0 = make_smem_desc(smem_u32(smem) + 16384, b_mn ? 8192u : 16u, 1024, UMMA_LAYOUT_SW128);
      const uint64_t ka = a_mn ? (2048u >> 4) : (32u >> 4), kb = b_mn ? (2048u >> 4) : (32u >> 4);
      const long long t0 = clock64();
      for (int i = 0; i < nmma; i += 4) {
        const uint64_t st = (uint64_t)(((i >> 2) & 3) * 49152) >> 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<false>(tmem_base, a0 + st + k * ka, b0 + st + k * kb, idesc, (i | k) ? 1u : 0u);
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  long long* out;
  CK(cudaMalloc(&out, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(major_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  long long* h = (long long*)malloc(sizeof(long long) * sms);
  const int nmma = 16384;
  for (int n_cols : {256, 128})
    for (int a_mn : {0, 1})
      for (int b_mn : {0, 1}) {
        for (int rep = 0; rep < 2; ++rep) {
          major_kernel<<<sms, 64, SMEM>>>(a_mn, b_mn, n_cols, nmma, out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double c = 0;
        for (int b = 0; b < sms; ++b) c += h[b];
        printf("M=128 N=%d K=16 bf16  A %s-major  B %s-major : %6.1f clk per MMA (ideal %d)\n", n_cols, a_mn ? "MN" : "K ",
               b_mn ? "MN" : "K ", c / sms / nmma, n_cols / 2);
      }
  return 0;
}
