// Microbenchmark: do SS-mode tcgen05.mma operand reads and asynchronous shared-memory fills (bulk copies, the
// path TMA tensor loads take) share the SM's shared-memory bandwidth?
//   warp 1: back-to-back tcgen05.mma M=128 N=256 K=16 (bf16, A and B in 128B-swizzled shared memory: 4 KB + 8 KB
//           read per instruction, 128 cycles each at full rate = 96 B/clk)
//   warp 0: a ring of 32 KB fills global -> shared memory, paced by `gap` cycles between issues (gap < 0: none).
//           mode 0: 1-D bulk copies from an L2-resident buffer
//           mode 1: 2-D tensor boxes {64 bf16, 256 rows} of a [2048 x 256] bf16 matrix (the fc1 weight stream of
//                   the fused MLP: 256 rows of 128 B, 512 B apart)
//           mode 2: the same boxes of a [256 x 2048] matrix (the fc2 weight stream: rows 4096 B apart)
//           mode 3: the patch-embedding gather: 5-D boxes {32 floats, 8, 1, 8, 4 volumes} (256 rows of 128 B out of
//                   320-byte patch rows) walking over 2368 volumes (776 MB: from HBM), with 1..4 boxes in flight
// Printed: cycles per MMA and fill bytes per clock, on one SM and on all SMs at once.  If the two streams share
// 128 B/clk, the MMA rate must drop as soon as the fills exceed ~32 B/clk - the model DESIGN.md section 3.1 uses
// for the fused MLP (80 B/clk of fills and GELU-tile writes) and the patch embedding (96 B/clk).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I3d_vit_ensemble_b200/csrc -o smem_bw_bench tools/smem_bw_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace vit3d::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int OPER_BYTES = 49152;          // A [128 x 64 bf16] 16 KB + B [256 x 64 bf16] 32 KB
constexpr int SLOT_BYTES = 32768;
constexpr int SLOTS = 4;
constexpr int SMEM_BYTES = OPER_BYTES + SLOTS * SLOT_BYTES;

__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// out[4 * block + 0] = MMA cycles, +1 = MMAs issued, +2 = fill cycles, +3 = fill bytes
__global__ void __launch_bounds__(96, 1) smem_bw_kernel(const uint8_t* src, const __grid_constant__ CUtensorMap tm, int mode,
                                                       int n_mma_groups, int gap, int do_mma, int depth, int ntiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar_mma, bar_slot[SLOTS];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < OPER_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < SLOTS; ++i) mbar_init(&bar_slot[i], 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc(UMMA_FMT_BF16, 128, 256, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const long long t0 = clock64();
    long long n = 0;
    if (do_mma) {
      for (int g = 0; g < n_mma_groups; ++g) {
        for (int k = 0; k < 64; ++k) {      // 64 MMAs (four K = 256 tiles) per commit round trip
          const uint64_t ad = make_smem_desc(sa + (k & 3) * 32, 16, 1024, UMMA_LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(sb + (k & 3) * 32, 16, 1024, UMMA_LAYOUT_SW128);
          umma<false>(tmem_base + 256, ad, bd, idesc, (k & 15) > 0 ? 1u : 0u);
        }
        umma_commit(&bar_mma);
        mbar_wait(&bar_mma, (uint32_t)(g & 1));
        n += 64;
      }
    } else {
      while (clock64() - t0 < 128ll * 64 * n_mma_groups) {}
    }
    const long long t1 = clock64();
    out[4 * blockIdx.x + 0] = t1 - t0;
    out[4 * blockIdx.x + 1] = n;
    stop = 1;
  } else if (warp == 0 && lane == 0 && gap >= 0) {
    const uint8_t* my = src + (size_t)(blockIdx.x % 16) * (SLOTS * SLOT_BYTES);     // 2 MB of source in all: L2 resident
    const long long t0 = clock64();
    long long bytes = 0, next = t0;
    int it = 0;
    while (!stop) {
      const int s = it % depth;
      if (it >= depth) mbar_wait(&bar_slot[s], (uint32_t)((it / depth - 1) & 1));     // previous copy into this slot landed
      while (clock64() < next) {}
      next = clock64() + gap;
      mbar_arrive_expect_tx(&bar_slot[s], SLOT_BYTES);
      uint8_t* dst = smem + OPER_BYTES + s * SLOT_BYTES;
      if (mode == 0) bulk_g2s(dst, my + s * SLOT_BYTES, SLOT_BYTES, &bar_slot[s]);
      else if (mode == 1) tma_load_2d(dst, &tm, &bar_slot[s], (it & 3) * 64, ((it >> 2) & 7) * 256);
      else if (mode == 2) tma_load_2d(dst, &tm, &bar_slot[s], ((it >> 2) & 7) * 256 + (it & 3) * 64, 0);
      else {
        // k-block it of this CTA's stream: 48 k-blocks (16 patch rows x 3 column blocks) per tile of 4 volumes
        const int tile = (int)blockIdx.x + (it / 48) * (int)gridDim.x, kb = it % 48;
        tma_load_5d(dst, &tm, &bar_slot[s], (kb % 3) * 32, 0, kb / 3, 0, (tile % ntiles) * 4);
      }
      bytes += SLOT_BYTES;
      ++it;
    }
    // drain
    for (int j = (it > depth ? it - depth : 0); j < it; ++j) mbar_wait(&bar_slot[j % depth], (uint32_t)((j / depth) & 1));
    out[4 * blockIdx.x + 2] = clock64() - t0;
    out[4 * blockIdx.x + 3] = bytes;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int dev = 0;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  const int sms = p.multiProcessorCount;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  uint8_t *src, *wbuf, *vols;
  long long* out;
  const int NVOL = 2368;
  CK(cudaMalloc(&src, 16 * SLOTS * SLOT_BYTES));
  CK(cudaMemset(src, 0, 16 * SLOTS * SLOT_BYTES));
  CK(cudaMalloc(&wbuf, 2048 * 256 * 2));
  CK(cudaMemset(wbuf, 0, 2048 * 256 * 2));
  CK(cudaMalloc(&vols, (size_t)NVOL * 327680));
  CK(cudaMemset(vols, 0, (size_t)NVOL * 327680));
  CK(cudaMalloc(&out, sizeof(long long) * 4 * sms));
  CUtensorMap tms[4];
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    cuuint64_t d1[2] = {256, 2048}, s1[1] = {512};
    cuuint32_t b1[2] = {64, 256};
    if (enc(&tms[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wbuf, d1, s1, b1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 2;
    cuuint64_t d2[2] = {2048, 256}, s2[1] = {4096};
    if (enc(&tms[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wbuf, d2, s2, b1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 2;
    // volumes (B,128,128,5) fp32 viewed as [80 floats][8 patches][16 rows][8 patches][B]
    cuuint64_t d3[5] = {80, 8, 16, 8, (cuuint64_t)NVOL};
    cuuint64_t s3[4] = {320, 2560, 16 * 2560, 327680};
    cuuint32_t b3[5] = {32, 8, 1, 8, 4};
    if (enc(&tms[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, vols, d3, s3, b3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 2;
    tms[0] = tms[1];
  }
  CK(cudaFuncSetAttribute(smem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  long long* h = (long long*)malloc(sizeof(long long) * 4 * sms);
  const int groups = 400;     // 25,600 MMAs = 3.3 M cycles at full rate
  const char* names[4] = {"1-D bulk (L2)", "fc1 weight boxes (L2)", "fc2 weight boxes (L2)", "patch gather 5-D (HBM)"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int grid : {1, sms}) {
      for (int do_mma : {1, 0}) {
        for (int gap : {-1, 1024, 512, 384, 256, 0}) {
          for (int depth : {4, 3, 2, 1}) {
            if (gap < 0 && (!do_mma || mode > 0)) continue;
            if (mode > 0 && gap != 0) continue;
            if (depth != 4 && (mode != 3 || grid == 1)) continue;
            if (gap < 0 && depth != 4) continue;
            CK(cudaMemset(out, 0, sizeof(long long) * 4 * sms));
            smem_bw_kernel<<<grid, 96, SMEM_BYTES>>>(src, tms[mode], mode, groups, gap, do_mma, depth, NVOL / 4, out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
            double mc = 0, mm = 0, fc = 0, fb = 0;
            for (int b = 0; b < grid; ++b) { mc += h[4 * b]; mm += h[4 * b + 1]; fc += h[4 * b + 2]; fb += h[4 * b + 3]; }
            const double cyc_per_mma = mm > 0 ? mc / mm : 0.0;
            const double fill = fc > 0 ? fb / fc : 0.0;
            printf("%-24s grid=%3d mma=%d gap=%5d boxes in flight=%d: %6.1f clk per MMA   fills %6.1f B/clk/SM (%5.2f clk per 128-B row, %6.0f B/clk chip)\n",
                   names[mode], grid, do_mma, gap, depth, cyc_per_mma, fill, fill > 0 ? 128.0 / fill : 0.0, fill * grid);
          }
        }
      }
    }
  }
  return 0;
}
