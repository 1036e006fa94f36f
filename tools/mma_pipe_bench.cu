// Microbenchmark: what does the producer <-> MMA-issuer mbarrier hand-shake of a TMA/tcgen05 main loop cost?
//   warp 0 (one lane): producer - waits for empty[s], (optionally issues a real 2-D bulk load of the stage,
//                      else just) arrives on full[s]
//   warp 1: MMA issuer - waits for full[s], issues `mpk` tcgen05.mma (M=128 N=256 K=16, bf16), commits to empty[s];
//           all 32 lanes walk the loop, one elected lane issues (the structure of the library's kernels), or a
//           single lane does everything (--style 1)
// Printed: cycles per MMA for ring depths 2..8 and 4 / 8 MMAs per stage (ideal 128).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I3d_vit_ensemble_b200/csrc -Iinclude -o mma_pipe_bench tools/mma_pipe_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace vit3d::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int STAGE_BYTES = 49152;     // A 16 KB + B 32 KB
constexpr int MAXST = 4;

__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// var: 0 = full hand-shake; 1 = per-stage commit only (no waits); 2 = per-stage wait only, on a barrier that is
// already complete (no commits); 3 = neither (bare issue); 4 = hand-shake with the wait for stage s+1 hoisted above
// the MMAs of stage s; 5 = bare without tcgen05.fence::after_thread_sync; 6 = hand-shake without that fence
template <int MPK>
__global__ void __launch_bounds__(96, 1) pipe_kernel(const uint8_t* src, int stages, int mpk, int nkb, int style, int load,
                                                    int var, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t full[8], empty[8], done, ready, dummy;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < MAXST * STAGE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    mbar_init(&ready, 1);
    mbar_init(&dummy, 1);
    fence_barrier_init();
    mbar_arrive(&ready);            // phase 0 of `ready` is complete from now on
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const int sring = stages < MAXST ? stages : MAXST;      // smem slots (more barriers than slots just alias the data)
  if (warp == 0) {
    if (lane == 0 && (var == 0 || var == 4 || var == 6 || var == 7 || var == 8)) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (load) {
          mbar_arrive_expect_tx(&full[s], 32768);
          bulk_g2s(smem + (s % sring) * STAGE_BYTES + 16384, src + (size_t)((kb + blockIdx.x) % 32) * 32768, 32768, &full[s]);
        } else {
          mbar_arrive(&full[s]);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(UMMA_FMT_BF16, 128, 256, 0, 0);
    const uint64_t ring = make_smem_desc(smem_u32(smem), 16, 1024, UMMA_LAYOUT_SW128);
    long long t0 = clock64();
    if (style == 0) {
      int s = 0; uint32_t ph = 0;
      if (var == 4) mbar_wait(&full[0], 0);
      for (int kb = 0; kb < nkb; ++kb) {
        if (var == 0) mbar_wait(&full[s], ph);
        if (var == 2) mbar_wait(&ready, 0);
        if (var == 6) mbar_wait(&full[s], ph);
        if (var < 5) tc_fence_after();
        const uint64_t ad = ring + (uint64_t)(((s % sring) * STAGE_BYTES) >> 4);
        const uint64_t bd = ad + (uint64_t)(16384 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < MPK; ++k) umma<false>(tmem_base, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, (kb | k) ? 1u : 0u);
          if (var == 0 || var == 4 || var == 6) umma_commit(&empty[s]);
          if (var == 1) umma_commit(&dummy);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
        if (var == 4 && kb + 1 < nkb) mbar_wait(&full[s], ph);     // next stage's operands, while this stage's MMAs run
      }
      if (elect_one()) umma_commit(&done);
      __syncwarp();
      mbar_wait(&done, 0);
    } else if (style == 2) {
      // one election for the whole loop: a single thread waits, issues and commits (var 7), optionally with the
      // try_wait for the NEXT stage issued between the 2nd and 3rd MMA of the current one (var 8), i.e. while the
      // thread would be blocked on the full MMA queue anyway
      if (elect_one()) {
        int s = 0; uint32_t ph = 0;
        bool ready = false;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!ready) mbar_wait(&full[s], ph);
          const uint64_t ad = ring + (uint64_t)(((s % sring) * STAGE_BYTES) >> 4);
          const uint64_t bd = ad + (uint64_t)(16384 >> 4);
          int s2 = s + 1; uint32_t ph2 = ph;
          if (s2 == stages) { s2 = 0; ph2 ^= 1; }
          ready = false;
#pragma unroll
          for (int k = 0; k < MPK; ++k) {
            umma<false>(tmem_base, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, (kb | k) ? 1u : 0u);
            if (var == 8 && k == 1 && kb + 1 < nkb) ready = mbar_try_wait(&full[s2], ph2);
          }
          umma_commit(&empty[s]);
          s = s2; ph = ph2;
        }
        umma_commit(&done);
      }
      __syncwarp();
      mbar_wait(&done, 0);
    } else if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t ad = ring + (uint64_t)(((s % sring) * STAGE_BYTES) >> 4);
        const uint64_t bd = ad + (uint64_t)(16384 >> 4);
        for (int k = 0; k < mpk; ++k) umma<false>(tmem_base, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, (kb | k) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int main() {
  int dev = 0;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  const int sms = p.multiProcessorCount;
  uint8_t* src;
  long long* out;
  CK(cudaMalloc(&src, 32 * 32768));
  CK(cudaMemset(src, 0, 32 * 32768));
  CK(cudaMalloc(&out, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(pipe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXST * STAGE_BYTES));
  CK(cudaFuncSetAttribute(pipe_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXST * STAGE_BYTES));
  CK(cudaFuncSetAttribute(pipe_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXST * STAGE_BYTES));
  long long* h = (long long*)malloc(sizeof(long long) * sms);
  const int total_mma = 16384;
  const char* vn[9] = {"hand-shake", "commit only", "wait only (ready barrier)", "bare", "hand-shake, wait hoisted", "bare, no tcgen05.fence", "hand-shake, no tcgen05.fence", "one thread, hand-shake", "one thread, early try_wait"};
  for (int var : {0, 3, 5, 6, 7, 8})
    for (int mpk : {4, 8, 16})
      for (int stages : {4}) {
        const int nkb = total_mma / mpk;
        if (mpk == 4) pipe_kernel<4><<<sms, 96, MAXST * STAGE_BYTES>>>(src, stages, mpk, nkb, var >= 7 ? 2 : 0, 0, var, out);
        else if (mpk == 8) pipe_kernel<8><<<sms, 96, MAXST * STAGE_BYTES>>>(src, stages, mpk, nkb, var >= 7 ? 2 : 0, 0, var, out);
        else pipe_kernel<16><<<sms, 96, MAXST * STAGE_BYTES>>>(src, stages, mpk, nkb, var >= 7 ? 2 : 0, 0, var, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double c = 0;
        for (int b = 0; b < sms; ++b) c += h[b];
        printf("%-28s MMAs/stage=%2d stages=%d (unrolled): %6.1f clk per MMA (ideal 128)\n", vn[var], mpk, stages, c / sms / total_mma);
      }
  return 0;
}
