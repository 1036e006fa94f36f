// Microbenchmark: HBM -> shared-memory throughput of the operand pattern of the deep-K GEMMs.
// One CTA per SM streams [128 rows x 128 B] boxes (TMA, SWIZZLE_128B) of its own row tiles through a ring of S stages;
// the consumer only hands the slots back.  Layouts of the same [M, 3072] bf16 matrix:
//   strided : row-major, a box = 128 pieces of 128 B, 6144 B apart (what a K-major operand tile is today)
//   tiled   : every [128 x 64] block contiguous (16 KB), a box = one contiguous chunk
//   pair    : strided, two adjacent k-blocks (256 B per row) issued back to back per ring slot pair
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I3d_vit_ensemble_b200/csrc -Iinclude -o tma_tile_read_bench tools/tma_tile_read_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace vit3d::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int BOX_BYTES = 128 * 128;

__global__ void __launch_bounds__(64, 1) read_kernel(const __grid_constant__ CUtensorMap tm, int tiled, int tiles, int nkb, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[12], empty[12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x)
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], BOX_BYTES);
        if (tiled) tma_load_2d(smem + s * BOX_BYTES, &tm, &full[s], 0, (t * nkb + kb) * 128);
        else tma_load_2d(smem + s * BOX_BYTES, &tm, &full[s], kb * 64, t * 128);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x)
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CK(cudaFuncSetAttribute(read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * BOX_BYTES));
  for (int K : {3072, 2048}) {
    const int nkb = K / 64;
    const int tiles = 4 * sms;                       // 4 row tiles per CTA
    const long long M = (long long)tiles * 128;
    const size_t bytes = (size_t)M * K * 2;
    uint8_t* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    uint8_t* flush;
    CK(cudaMalloc(&flush, 256 << 20));
    for (int tiled : {0, 1}) {
      CUtensorMap tm;
      cuuint64_t gdim[2], gstr[1];
      cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
      if (tiled) { gdim[0] = 64; gdim[1] = (cuuint64_t)M * nkb; gstr[0] = 128; }
      else { gdim[0] = K; gdim[1] = M; gstr[0] = (cuuint64_t)K * 2; }
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      for (int stages : {2, 3, 4, 6, 8, 12}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaMemset(flush, rep, 256 << 20));     // evict the matrix from L2
          cudaEvent_t e0, e1;
          CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
          CK(cudaEventRecord(e0));
          read_kernel<<<sms, 64, 12 * BOX_BYTES>>>(tm, tiled, tiles, nkb, stages);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        printf("K=%d %-8s stages=%2d (%3d KB in flight / SM): %7.1f us  %6.0f GB/s\n", K, tiled ? "tiled" : "strided", stages,
               stages * 16, best * 1e3, bytes / best / 1e6);
      }
    }
    CK(cudaFree(buf));
    CK(cudaFree(flush));
  }
  return 0;
}
