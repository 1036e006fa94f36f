// Microbenchmark: how fast can the epilogue warps of a persistent GEMM CTA push a [M, N] bf16 output to HBM?
// No MMA, no math: 16 warps per CTA (1 CTA / SM) replay the store pattern of the fc1 GEMM
// (tile 128 x 256, CTA = (n-tile, m-tile group), warp = 32 rows x 64 columns) with different store paths:
//   0  bulk tensor store, box 32 x 32 (64-byte rows, SWIZZLE_64B), one staging buffer per warp
//   1  same, two staging buffers per warp (one store may still be reading while the next is staged)
//   2  bulk tensor store, box 32 x 64 (128-byte rows, SWIZZLE_128B), one buffer per warp
//   3  same, two buffers
//   4  coalesced st.global.v4 from the staged panel: a warp instruction covers 4 rows x 128 bytes
//   5  st.global.v4 row-owner: every lane writes its own row (32 lines per instruction)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_store_bench tools/tma_store_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- load-latency probe: a 17th warp fetches 16 KB A-tile blocks (128 rows x 64 bf16) while the 16 store
// warps run, either with a bulk tensor load (loadpath 1) or with 16-byte cp.async (loadpath 2), and
// accumulates the issue -> data-landed latency.
__device__ __forceinline__ void mbar_init_(uint64_t* b) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(b))); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ void load_probe(const CUtensorMap* tmA, const __nv_bfloat16* A, uint8_t* dst, uint64_t* bar, int loadpath, volatile int* stop, long long* out) {
  const int lane = threadIdx.x & 31;
  long long total = 0, n = 0;
  uint32_t ph = 0;
  int row0 = (blockIdx.x * 977) % 60000;
  while (!*stop) {
    row0 = (row0 + 128 * 148) % 66000;
    const long long t0 = clock64();
    if (loadpath == 1) {
      if (lane == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)tmA), "r"(smem_u32(bar)), "r"(0), "r"(row0) : "memory");
      }
      while (!mbar_try(bar, ph)) {}
      ph ^= 1;
    } else {
      // 128 rows x 128 bytes = 1024 16-byte chunks, 32 per lane
      for (int i = 0; i < 32; ++i) {
        const int c = i * 32 + lane, r = c >> 3, ch = c & 7;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + r * 128 + ((ch ^ (r & 7)) << 4))), "l"(A + (size_t)(row0 + r) * 256 + ch * 8) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
    }
    total += clock64() - t0;
    ++n;
  }
  if (lane == 0) { atomicAdd((unsigned long long*)&out[0], (unsigned long long)total); atomicAdd((unsigned long long*)&out[1], (unsigned long long)n); }
}

template <int MODE>
__global__ void __launch_bounds__(544, 1) store_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* out, int M, int N, int tiles_m, int tiles_n, int delay,
                                                       const __grid_constant__ CUtensorMap tmA, const __nv_bfloat16* A, int loadpath, long long* lat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t lbar;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { mbar_init_(&lbar); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x >= 512) {
    if (loadpath > 0) load_probe(&tmA, A, smem + 16 * 4096 * 2, &lbar, loadpath, &stop, lat);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, part = warp >> 2;
  constexpr bool WIDE = MODE == 2 || MODE == 3 || MODE == 4 || MODE == 5;
  constexpr int NBUF = (MODE == 1 || MODE == 3) ? 2 : 1;
  constexpr int PANEL = WIDE ? 4096 : 2048;
  uint8_t* base = smem + warp * PANEL * NBUF;
  // fill staging with something
  for (int i = lane; i < PANEL * NBUF / 4; i += 32) reinterpret_cast<uint32_t*>(base)[i] = warp * 1000 + i;
  __syncwarp();
  const int tn = blockIdx.x % tiles_n, g = blockIdx.x / tiles_n, cpn = gridDim.x / tiles_n;
  int cnt = 0;
  for (int t = g; t < tiles_m; t += cpn) {
    const int m_base = t * 128 + q * 32, n_base = tn * 256 + part * 64;
    if (MODE <= 3) {
      constexpr int ROUNDS = WIDE ? 1 : 2;
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r, ++cnt) {
        uint8_t* buf = base + (cnt % NBUF) * PANEL;
        // simulated epilogue math between stores
        if (delay > 0) { long long t0 = clock64(); while (clock64() - t0 < delay / ROUNDS) {} }
        if (lane == 0) { if (NBUF == 1) wait_read<0>(); else wait_read<1>(); }
        __syncwarp();
        // touch the buffer like the epilogue does (4 or 8 x st.shared.v4 per lane)
        for (int j = 0; j < (WIDE ? 8 : 4); ++j)
          *reinterpret_cast<uint4*>(buf + lane * (WIDE ? 128 : 64) + ((j ^ (WIDE ? (lane & 7) : ((lane >> 1) & 3))) << 4)) = make_uint4(cnt, lane, j, t);
        fence_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&tm, smem_u32(buf), n_base + r * 32, m_base); commit(); }
      }
    } else if (MODE == 4) {
      if (delay > 0) { long long t0 = clock64(); while (clock64() - t0 < delay) {} }
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(base + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(cnt, lane, j, t);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), ch = lane & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(base + row * 128 + ((ch ^ (row & 7)) << 4));
        if (m_base + row < M) *reinterpret_cast<uint4*>(out + (size_t)(m_base + row) * N + n_base + ch * 8) = v;
      }
      __syncwarp();
    } else {
      if (delay > 0) { long long t0 = clock64(); while (clock64() - t0 < delay) {} }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (m_base + lane < M) *reinterpret_cast<uint4*>(out + (size_t)(m_base + lane) * N + n_base + j * 8) = make_uint4(cnt, lane, j, t);
    }
  }
  if (MODE <= 3 && lane == 0) wait_all();
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (threadIdx.x == 0) stop = 1;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap g_tmA;
static __nv_bfloat16* g_A = nullptr;
static long long* g_lat = nullptr;
static int g_loadpath = 0;
static double g_last_lat = 0;

template <int MODE>
float run(const CUtensorMap& tm, __nv_bfloat16* out, int M, int N, int delay) {
  const int tiles_m = (M + 127) / 128, tiles_n = N / 256;
  int cpn = 148 / tiles_n; if (cpn > tiles_m) cpn = tiles_m;
  const int grid = cpn * tiles_n;
  const int smem = 16 * 4096 * 2 + 16384;
  CK(cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) store_kernel<MODE><<<grid, 544, smem>>>(tm, out, M, N, tiles_m, tiles_n, delay, g_tmA, g_A, g_loadpath, g_lat);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(g_lat, 0, 16));
  CK(cudaEventRecord(e0));
  const int iters = 10;
  for (int i = 0; i < iters; ++i) store_kernel<MODE><<<grid, 544, smem>>>(tm, out, M, N, tiles_m, tiles_n, delay, g_tmA, g_A, g_loadpath, g_lat);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[2]; CK(cudaMemcpy(h, g_lat, 16, cudaMemcpyDeviceToHost));
  g_last_lat = h[1] ? (double)h[0] / (double)h[1] : 0.0;
  return ms / iters * 1e3f;
}

int main(int argc, char** argv) {
  const int M = 66560;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fn;
  CK(cudaMalloc(&g_A, (size_t)66560 * 256 * 2)); CK(cudaMemset(g_A, 0, (size_t)66560 * 256 * 2));
  CK(cudaMalloc(&g_lat, 16));
  {
    cuuint64_t gd[2] = {256, 66560}; cuuint64_t gs[1] = {512}; cuuint32_t es[2] = {1, 1}; cuuint32_t bx[2] = {64, 128};
    if (enc(&g_tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_A, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
  }
  for (int N : {2048}) {
    __nv_bfloat16* out; CK(cudaMalloc(&out, (size_t)M * N * 2));
    CUtensorMap t32, t64;
    cuuint64_t gd[2] = {(cuuint64_t)N, (cuuint64_t)M}; cuuint64_t gs[1] = {(cuuint64_t)N * 2}; cuuint32_t es[2] = {1, 1};
    cuuint32_t b32[2] = {32, 32}, b64[2] = {64, 32};
    if (enc(&t32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gd, gs, b32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
    if (enc(&t64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gd, gs, b64, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
    const double mb = (double)M * N * 2 / 1e6;
    for (int loadpath : {0, 1, 2})
    for (int delay : {0, 2500}) {
      g_loadpath = loadpath;
      float t[6]; double l[6];
      t[0] = run<0>(t32, out, M, N, delay); l[0] = g_last_lat; t[1] = run<1>(t32, out, M, N, delay); l[1] = g_last_lat;
      t[2] = run<2>(t64, out, M, N, delay); l[2] = g_last_lat; t[3] = run<3>(t64, out, M, N, delay); l[3] = g_last_lat;
      t[4] = run<4>(t64, out, M, N, delay); l[4] = g_last_lat; t[5] = run<5>(t64, out, M, N, delay); l[5] = g_last_lat;
      printf("N=%d (%.0f MB) math %d clk/tile, concurrent 16 KB loads: %s\n", N, mb, delay, loadpath == 0 ? "none" : (loadpath == 1 ? "bulk tensor (TMA)" : "cp.async 16 B"));
      const char* names[6] = {"tma32x32", "tma32x32x2buf", "tma32x64", "tma32x64x2buf", "stg_coalesced", "stg_rowowner"};
      for (int i = 0; i < 6; ++i) printf("    stores %-14s %.1f us (%.2f TB/s)   load latency %.0f clk\n", names[i], t[i], mb / t[i], l[i]);
    }
    CK(cudaFree(out));
  }
  return 0;
}
