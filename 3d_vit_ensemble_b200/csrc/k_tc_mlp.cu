// Fused MLP forward (a3, modeling.py:118-124, inference):  out = x + fc2(GELU(fc1(xn)))  in ONE kernel.
// The [M, d] intermediate (d = 2048 / 3072) never leaves the SM: per 128-row tile, fc1 is computed 64
// output columns at a time into a double-buffered TMEM accumulator, 16 epilogue warps add the bias, apply
// the GELU and write the bf16 result straight into shared memory in the K-major 128B-swizzled layout that
// tcgen05 reads, and the same CTA immediately feeds it to fc2, whose [128 x 256] accumulator stays in TMEM
// for the whole tile.  HBM traffic per row drops from 2*(d*2) + ... bytes to the 512 B in / 1 KB out of
// the tile itself; the weights stream from L2 through a 3-stage 32 KB TMA ring.
//
//   warp 0: TMA producer     warp 1: tcgen05.mma issuer (+TMEM alloc)     warps 2..17: epilogue
//   TMEM (512 cols): [0,128) fc1 accumulators (2 x 64), [128,384) fc2 accumulator.
//   smem: xn tile 64 KB | weight ring 3 x 32 KB | GELU(A) tiles 2 x 16 KB | barriers.
#include <stdlib.h>

#include "ptx.cuh"
#include "tc.cuh"
#include "tc_epilogue.cuh"

namespace vit3d {

using namespace ptx;

constexpr int ML_H = 256;             // hidden size (K of fc1, N of fc2)
constexpr int ML_NC = 64;             // fc1 columns per chunk (= K of one fc2 step)
constexpr int ML_XN_BYTES = 128 * ML_H * 2;       // 65536
constexpr int ML_WST = 3;
constexpr int ML_W_BYTES = 32768;                 // [64 x 256] of W1 or [256 x 64] of W2, bf16
constexpr int ML_A_BYTES = 128 * ML_NC * 2;       // 16384
constexpr int ML_THREADS = 64 + 512;
constexpr int ML_MAX_D = 4096;                    // fc1 bias is staged in shared memory (16 KB)
constexpr int ML_SMEM = ML_XN_BYTES + ML_WST * ML_W_BYTES + 2 * ML_A_BYTES + ML_MAX_D * 4 + 1024 + 256;

__global__ void __launch_bounds__(ML_THREADS, 1)
tc_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ b1, TcEpilogue ep, int M, int d,
              long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_xn = smem;
  uint8_t* s_w = smem + ML_XN_BYTES;
  uint8_t* s_a = s_w + ML_WST * ML_W_BYTES;
#define DBG(slot) do { if (dbg && blockIdx.x == 0) dbg[(slot)] = clock64(); } while (0)
  float* s_b1 = reinterpret_cast<float*>(s_a + 2 * ML_A_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + 2 * ML_A_BYTES + ML_MAX_D * 4);
  uint64_t* xn_full = bars;            // 1
  uint64_t* xn_free = bars + 1;        // 512: the final epilogue (which reuses the xn region) is done
  uint64_t* w_full = bars + 2;         // [3]
  uint64_t* w_empty = bars + 5;        // [3]
  uint64_t* acc1_full = bars + 8;      // [2]
  uint64_t* acc1_empty = bars + 10;    // [2] 256 each
  uint64_t* a_full = bars + 12;        // [2] 256 each
  uint64_t* a_empty = bars + 14;       // [2]
  uint64_t* acc2_full = bars + 16;
  uint64_t* acc2_empty = bars + 17;    // 512
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (M + 127) / 128;
  const int nch = d / ML_NC;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    mbar_init(xn_full, 1);
    mbar_init(xn_free, 16);
    for (int s = 0; s < ML_WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 8);      // one elected arrival per epilogue warp of the parity group
      mbar_init(&a_full[b], 8);
      mbar_init(&a_empty[b], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  // the fc1 bias sits on the GELU critical path once per chunk: keep it in shared memory (29-cycle LDS
  // instead of a global load that misses the ~30 KB of L1 left beside 200 KB of shared memory)
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_b1[i] = b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Weight-ring schedule per tile (same in producer and MMA warp):
  //   op 0: W1[0]; then for c = 1..nch-1: W1[c], W2[c-1]; finally W2[nch-1]      (2*nch ops)
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        mbar_wait(xn_free, (it & 1) ^ 1);
        mbar_arrive_expect_tx(xn_full, ML_XN_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(s_xn + kb * 16384, &tmX, xn_full, kb * 64, tile * 128);
        for (int op = 0; op < 2 * nch; ++op) {
          // op -> (which matrix, chunk)
          int c;
          bool is_w1;
          if (op == 0) { is_w1 = true; c = 0; }
          else if (op == 2 * nch - 1) { is_w1 = false; c = nch - 1; }
          else { is_w1 = (op & 1) != 0; c = is_w1 ? (op + 1) / 2 : op / 2 - 1; }
          mbar_wait(&w_empty[stage], wphase ^ 1);
          if (it == 0 && op < 64) DBG(op);
          uint8_t* dst = s_w + stage * ML_W_BYTES;
          mbar_arrive_expect_tx(&w_full[stage], ML_W_BYTES);
          if (is_w1) {
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + kb * 8192, &tmW1, &w_full[stage], kb * 64, c * ML_NC);
          } else {
            tma_load_2d(dst, &tmW2, &w_full[stage], c * ML_NC, 0);
          }
          if (++stage == ML_WST) { stage = 0; wphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 128, ML_NC, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_ab(UMMA_FMT_F16, UMMA_FMT_F16, 128, ML_H);   // fp16 GELU tile x fp16 W2
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      // running counts of uses of the double-buffered resources (phase = (use index >> 1) & 1)
      int g1_count = 0, g2_count = 0;
      const uint32_t xn_addr = smem_u32(s_xn);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        mbar_wait(xn_full, it & 1);
        for (int op = 0; op < 2 * nch; ++op) {
          int c;
          bool is_w1;
          if (op == 0) { is_w1 = true; c = 0; }
          else if (op == 2 * nch - 1) { is_w1 = false; c = nch - 1; }
          else { is_w1 = (op & 1) != 0; c = is_w1 ? (op + 1) / 2 : op / 2 - 1; }
          const uint32_t wst = smem_u32(s_w + stage * ML_W_BYTES);
          if (is_w1) {
            // fc1 chunk c -> acc1[c & 1]
            const int buf = g1_count & 1;
            if (it == 0 && op < 64) DBG(64 + op * 4);
            mbar_wait(&acc1_empty[buf], ((g1_count >> 1) & 1) ^ 1);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 1);
            mbar_wait(&w_full[stage], wphase);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 2);
            tc_fence_after();
            const uint32_t dcol = tmem_base + buf * ML_NC;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_smem_desc(xn_addr + kb * 16384 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                const uint64_t bd = make_smem_desc(wst + kb * 8192 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                umma<false>(dcol, ad, bd, idesc1, (kb | k) ? 1u : 0u);
              }
            umma_commit(&w_empty[stage]);
            umma_commit(&acc1_full[buf]);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 3);
            ++g1_count;
          } else {
            // fc2 step c: acc2 += A[c & 1] (128 x 64) * W2[:, c*64 .. +64]^T
            const int buf = g2_count & 1;
            if (it == 0 && op < 64) DBG(64 + op * 4);
            if (c == 0) mbar_wait(acc2_empty, (it & 1) ^ 1);
            mbar_wait(&a_full[buf], (g2_count >> 1) & 1);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 1);
            mbar_wait(&w_full[stage], wphase);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 2);
            tc_fence_after();
            const uint32_t aaddr = smem_u32(s_a + buf * ML_A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc(aaddr + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
              const uint64_t bd = make_smem_desc(wst + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
              umma<false>(tmem_base + 128, ad, bd, idesc2, (c | k) ? 1u : 0u);
            }
            umma_commit(&w_empty[stage]);
            umma_commit(&a_empty[buf]);
            if (c == nch - 1) umma_commit(acc2_full);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 3);
            ++g2_count;
          }
          if (++stage == ML_WST) { stage = 0; wphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                  // TMEM lane quarter
    const int grp = (warp - 2) >> 2;         // 0..3
    const int parity = grp >> 1;             // which fc1 accumulator / A buffer this warp serves
    const int half = grp & 1;                // which 32 of the chunk's 64 columns
    const int row = q * 32 + lane;           // row of the tile this thread owns in the GELU phase
    const uint32_t a_row = smem_u32(s_a + parity * ML_A_BYTES) + row * 128;
    const uint32_t stage = smem_u32(s_xn + (warp - 2) * 2048);   // final-epilogue transpose tile (xn region)
    int it = 0;
    int use = 0;                             // how many chunks this warp has processed (its buffer's use index)
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      for (int c = parity; c < nch; c += 2, ++use) {
        const bool stamp = it == 0 && c < 32 && lane == 0 && (warp == 2 || warp == 10);
        if (stamp) DBG(320 + c * 6);
        mbar_wait(&acc1_full[parity], use & 1);
        if (stamp) DBG(320 + c * 6 + 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + parity * ML_NC + half * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[parity]);
        if (stamp) DBG(320 + c * 6 + 2);
        float v[32];
        const float* bias = s_b1 + c * ML_NC + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bias + j);
          v[j] = __uint_as_float(r[j]) + b.x;
          v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
          v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = gelu_pair_f16(v[2 * j], v[2 * j + 1]);
        if (stamp) DBG(320 + c * 6 + 3);
        mbar_wait(&a_empty[parity], (use & 1) ^ 1);       // the fc2 step that read this A buffer has retired
        if (stamp) DBG(320 + c * 6 + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = half * 4 + j;                  // 16-byte chunk of the 128-byte row
          st_shared_v4(a_row + ((chunk ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[parity]);
        if (stamp) DBG(320 + c * 6 + 5);
      }
      // ---- tile epilogue: out = acc2 + b2 + residual (fp32), 64 columns per warp
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 128 + grp * 64;
      epilogue_rows<64, true>(ep, nullptr, nullptr, taddr, stage, lane, tile * 128 + q * 32, grp * 64, M, ML_H);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc2_empty);
        mbar_arrive(xn_free);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

extern long long* g_mlp_dbg;

// ============================================================================ CTA-pair version
// Same algorithm with two SMs cooperating on a 256-row tile (tcgen05 cta_group::2) and 128 fc1 columns
// per chunk.  Each CTA owns 128 rows (its xn tile, its GELU tile, its half of both accumulators in its own
// TMEM) and loads only HALF of every weight chunk (64 of the 128 fc1 rows, 128 of the 256 fc2 rows); the
// leader CTA's single issuing lane drives both tensor cores.
//   Why 128-column chunks: a clock64 timeline of the single-CTA kernel showed the tensor pipe needs ~78
//   cycles per tcgen05.mma however small N is, so N = 64 fc1 instructions run at 40 % of the rate of
//   N >= 128 ones; and each mbarrier wait of the issuing lane costs ~125 cycles even when already complete,
//   so fewer, larger steps win.  Why a pair: per-CTA weight traffic halves (64 KB per 128-column chunk),
//   which fits the 96 KB ring beside the 64 KB xn tile and two 32 KB GELU tiles.
//   Barriers that gate the MMA issuer live in the LEADER (peer TMA bytes and peer epilogue arrivals are
//   signalled remotely); barriers that gate producers / epilogues are per CTA, released by multicast commit.
constexpr int MP_NC = 128;                      // fc1 columns per chunk
constexpr int MP_WST = 3;
constexpr int MP_W_BYTES = 32768;               // per CTA: [64 x 256] of W1 or [128 x 128] of W2
constexpr int MP_A_BYTES = 128 * MP_NC * 2;     // 32768: GELU tile, two K-major k-blocks of [128 x 64]
constexpr int MP_SMEM = ML_XN_BYTES + MP_WST * MP_W_BYTES + 2 * MP_A_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ML_THREADS, 1)
tc_mlp_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ b1, TcEpilogue ep, int M, int d,
                   long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_xn = smem;
  uint8_t* s_w = smem + ML_XN_BYTES;
  uint8_t* s_a = s_w + MP_WST * MP_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + 2 * MP_A_BYTES);
  uint64_t* xn_full = bars;            // leader: 1 arrive + bytes of both CTAs
  uint64_t* xn_free = bars + 1;        // per CTA: 512 (own final epilogue done)
  uint64_t* w_full = bars + 2;         // [3] leader
  uint64_t* w_empty = bars + 5;        // [3] per CTA (multicast commit)
  uint64_t* acc1_full = bars + 8;      // [2] per CTA (multicast commit)
  uint64_t* acc1_empty = bars + 10;    // [2] leader: 1024 = 512 threads of each CTA
  uint64_t* a_full = bars + 12;        // [2] leader: 1024
  uint64_t* a_empty = bars + 14;       // [2] per CTA (multicast commit)
  uint64_t* acc2_full = bars + 16;     // per CTA (multicast commit)
  uint64_t* acc2_empty = bars + 17;    // leader: 1024
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ptiles = (M + 255) / 256;
  const int nclusters = gridDim.x / 2;
  const int cid = blockIdx.x / 2;
  const int nch = d / MP_NC;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    mbar_init(xn_full, 1);
    mbar_init(xn_free, 16);
    for (int s = 0; s < MP_WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 16);     // one elected arrival per warp of the group that serves buffer b, both CTAs
      mbar_init(&a_full[b], 16);
      mbar_init(&a_empty[b], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 32);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // op schedule per tile: W1[0]; then W1[c], W2[c-1] for c = 1..nch-1; then W2[nch-1]
  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs, own halves)
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      for (int pt = cid; pt < ptiles; pt += nclusters, ++it) {
        mbar_wait(xn_free, (it & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(xn_full, 2 * ML_XN_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(s_xn + kb * 16384, &tmX, xn_full, kb * 64, pt * 256 + (int)rank * 128);
        for (int op = 0; op < 2 * nch; ++op) {
          int c;
          bool is_w1;
          if (op == 0) { is_w1 = true; c = 0; }
          else if (op == 2 * nch - 1) { is_w1 = false; c = nch - 1; }
          else { is_w1 = (op & 1) != 0; c = is_w1 ? (op + 1) / 2 : op / 2 - 1; }
          mbar_wait(&w_empty[stage], wphase ^ 1);
          if (it == 0 && op < 64 && leader) DBG(op);
          uint8_t* dst = s_w + stage * MP_W_BYTES;
          if (leader) mbar_arrive_expect_tx(&w_full[stage], 2 * MP_W_BYTES);
          if (is_w1) {      // this CTA's 64 of the chunk's 128 fc1 rows: 4 k-blocks of [64 rows x 64 K]
            for (int kb = 0; kb < 4; ++kb)
              tma_load_2d_pair(dst + kb * 8192, &tmW1, &w_full[stage], kb * 64, c * MP_NC + (int)rank * 64);
          } else {          // this CTA's 128 of the 256 fc2 rows: 2 k-blocks of [128 rows x 64 K]
            for (int kb = 0; kb < 2; ++kb)
              tma_load_2d_pair(dst + kb * 16384, &tmW2, &w_full[stage], c * MP_NC + kb * 64, (int)rank * 128);
          }
          if (++stage == MP_WST) { stage = 0; wphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 256, MP_NC, 0, 0);
      // fc2 runs in fp16 x fp16: the GELU tile is written as fp16 and W2 arrives as an fp16 shadow
      constexpr uint32_t idesc2 = make_idesc_ab(UMMA_FMT_F16, UMMA_FMT_F16, 256, ML_H);
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      int g1_count = 0, g2_count = 0;
      const uint32_t xn_addr = smem_u32(s_xn);
      for (int pt = cid; pt < ptiles; pt += nclusters, ++it) {
        mbar_wait(xn_full, it & 1);
        for (int op = 0; op < 2 * nch; ++op) {
          int c;
          bool is_w1;
          if (op == 0) { is_w1 = true; c = 0; }
          else if (op == 2 * nch - 1) { is_w1 = false; c = nch - 1; }
          else { is_w1 = (op & 1) != 0; c = is_w1 ? (op + 1) / 2 : op / 2 - 1; }
          const uint32_t wst = smem_u32(s_w + stage * MP_W_BYTES);
          if (is_w1) {
            const int buf = g1_count & 1;
            if (it == 0 && op < 64) DBG(64 + op * 4);
            mbar_wait(&acc1_empty[buf], ((g1_count >> 1) & 1) ^ 1);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 1);
            mbar_wait(&w_full[stage], wphase);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 2);
            tc_fence_after();
            const uint32_t dcol = tmem_base + buf * MP_NC;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_smem_desc(xn_addr + kb * 16384 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                const uint64_t bd = make_smem_desc(wst + kb * 8192 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                umma_pair_bf16(dcol, ad, bd, idesc1, (kb | k) ? 1u : 0u);
              }
            umma_commit_pair(&w_empty[stage], 3);
            umma_commit_pair(&acc1_full[buf], 3);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 3);
            ++g1_count;
          } else {
            const int buf = g2_count & 1;
            if (it == 0 && op < 64) DBG(64 + op * 4);
            if (c == 0) mbar_wait(acc2_empty, (it & 1) ^ 1);
            mbar_wait(&a_full[buf], (g2_count >> 1) & 1);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 1);
            mbar_wait(&w_full[stage], wphase);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 2);
            tc_fence_after();
            const uint32_t aaddr = smem_u32(s_a + buf * MP_A_BYTES);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_smem_desc(aaddr + kb * 16384 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                const uint64_t bd = make_smem_desc(wst + kb * 16384 + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
                umma_pair_bf16(tmem_base + 256, ad, bd, idesc2, (c | kb | k) ? 1u : 0u);
              }
            umma_commit_pair(&w_empty[stage], 3);
            umma_commit_pair(&a_empty[buf], 3);
            if (c == nch - 1) umma_commit_pair(acc2_full, 3);
            if (it == 0 && op < 64) DBG(64 + op * 4 + 3);
            ++g2_count;
          }
          if (++stage == MP_WST) { stage = 0; wphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps (both CTAs, own 128 rows)
    // Two groups of 8 warps alternate chunks (group p serves accumulator / GELU buffer p), so two chunk
    // epilogues are in flight at once: one group's TMEM read-out (64 KB at ~64 B/clk = ~1k cycles, a
    // hardware floor) overlaps the other's GELU arithmetic and shared-memory stores.
    const int q = warp & 3;                  // TMEM lane quarter
    const int grp = (warp - 2) >> 2;         // 0..3
    const int parity = grp >> 1;             // which fc1 accumulator / GELU buffer this warp serves
    const int half = grp & 1;                // which 64 of the chunk's 128 columns
    const int row = q * 32 + lane;
    const uint32_t stage = smem_u32(s_xn + (warp - 2) * 2048);
    // columns half*64 .. +64 of the 128-wide GELU tile are exactly k-block `half` (128 bytes per row)
    const uint32_t a_row = smem_u32(s_a + parity * MP_A_BYTES) + half * 16384 + row * 128;
    int it = 0;
    int use = 0;                             // chunks this warp has processed
    for (int pt = cid; pt < ptiles; pt += nclusters, ++it) {
      for (int c = parity; c < nch; c += 2, ++use) {
        const bool stamp = it == 0 && c < 32 && lane == 0 && (warp == 2 || warp == 10) && leader;
        if (stamp) DBG(320 + c * 6);
        mbar_wait(&acc1_full[parity], use & 1);
        if (stamp) DBG(320 + c * 6 + 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + parity * MP_NC + half * 64;
        tmem_ld_32x32b_x32(tcol, r0);
        tmem_ld_32x32b_x32(tcol + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&acc1_empty[parity], 0);
        if (stamp) DBG(320 + c * 6 + 2);
        const float* bias = b1 + c * MP_NC + half * 64;
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + j));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + 32 + j));
          pk[j / 2] = gelu_pair_f16(__uint_as_float(r0[j]) + ba.x, __uint_as_float(r0[j + 1]) + ba.y);
          pk[j / 2 + 1] = gelu_pair_f16(__uint_as_float(r0[j + 2]) + ba.z, __uint_as_float(r0[j + 3]) + ba.w);
          pk[16 + j / 2] = gelu_pair_f16(__uint_as_float(r1[j]) + bb.x, __uint_as_float(r1[j + 1]) + bb.y);
          pk[16 + j / 2 + 1] = gelu_pair_f16(__uint_as_float(r1[j + 2]) + bb.z, __uint_as_float(r1[j + 3]) + bb.w);
        }
        if (stamp) DBG(320 + c * 6 + 3);
        mbar_wait(&a_empty[parity], (use & 1) ^ 1);
        if (stamp) DBG(320 + c * 6 + 4);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(a_row + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&a_full[parity], 0);
        if (stamp) DBG(320 + c * 6 + 5);
      }
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + grp * 64;
      epilogue_rows<64, true>(ep, nullptr, nullptr, taddr, stage, lane, pt * 256 + (int)rank * 128 + q * 32, grp * 64, M, ML_H);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_remote(acc2_empty, 0);
        mbar_arrive(xn_free);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody leaves while the partner may still signal or read this CTA
  if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

static int launch_mlp_pair(const void* xn, const void* w1, const float* b1, const void* w2, const TcEpilogue& ep, int M,
                           int H, int d, cudaStream_t st) {
  CUtensorMap tx, tw1, tw2;
  int rc = make_tmap_2d(&tx, xn, 2, M, H, H, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw1, w1, 2, d, H, H, 64, 64, 128);               // W1 [d, 256]: box [64 rows x 64 cols]
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw2, w2, 2, H, d, d, 128, 64, 128);              // W2 [256, d]: box [128 rows x 64 cols]
  if (rc != VIT3D_OK) return rc;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(tc_mlp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MP_SMEM));
    configured_dev = dev;
  }
  const int ptiles = ceil_div(M, 256);
  int clusters = sm_count() / 2;
  if (clusters > ptiles) clusters = ptiles;
  tc_mlp_pair_kernel<<<2 * clusters, ML_THREADS, MP_SMEM, st>>>(tx, tw1, tw2, b1, ep, M, d, g_mlp_dbg);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

long long* g_mlp_dbg = nullptr;   // debug timeline buffer (tools/probe_mlp.py --timeline)
extern "C" __attribute__((visibility("default"))) void vit3d_debug_set_mlp_timeline(long long* p) { g_mlp_dbg = p; }

bool tc_mlp_supported(int M, int H, int d) {
  return M > 0 && H == ML_H && d % (2 * ML_NC) == 0 && d >= 2 * ML_NC && d <= ML_MAX_D;
}

extern long long* g_mlp_dbg;
int tc_mlp_fwd(const void* xn, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual,
               float* out, int M, int H, int d, cudaStream_t st) {
  if (!tc_mlp_supported(M, H, d)) V3_UNSUPPORTED("fused MLP: unsupported shape M=%d H=%d d=%d", M, H, d);
  CUtensorMap tx, tw1, tw2;
  int rc = make_tmap_2d(&tx, xn, 2, M, H, H, 128, 64, 128);          // box [128 rows x 64 cols]
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw1, w1, 2, d, H, H, ML_NC, 64, 128);            // W1 [d, 256]: box [64 rows x 64 cols]
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw2, w2, 2, H, d, d, ML_H, 64, 128);             // W2 [256, d]: box [256 rows x 64 cols]
  if (rc != VIT3D_OK) return rc;
  TcEpilogue ep;
  ep.bias = b2; ep.residual = residual; ep.out = out; ep.out_f32 = 1;
  static const bool no_pair = getenv("VIT3D_MLP_NO_PAIR") != nullptr;
  // the pair kernel alternates its two epilogue groups over 128-column chunks: needs an even chunk count
  if (M > 128 && d % (2 * MP_NC) == 0 && !no_pair) return launch_mlp_pair(xn, w1, b1, w2, ep, M, H, d, st);
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(tc_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM));
    configured_dev = dev;
  }
  const int tiles = ceil_div(M, 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  tc_mlp_kernel<<<grid, ML_THREADS, ML_SMEM, st>>>(tx, tw1, tw2, b1, ep, M, d, g_mlp_dbg);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

}  // namespace vit3d
