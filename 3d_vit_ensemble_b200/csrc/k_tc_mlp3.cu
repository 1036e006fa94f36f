// Fused MLP block, forward, in the 128-column-chunk layout of the backward kernel (k_tc_mlp_bwd.cu):
//
//     y = x + fc2(GELU(fc1(xn)))    (+ the next LayerNorm, LN = 1 bf16 / LN = 2 fp32-only, as k_tc_mlp2.cu)
//
// k_tc_mlp2.cu runs 256-column chunks with ONE fc1 accumulator and ONE GELU tile: its 16 epilogue warps are busy ~79 %
// of the time and wait for an accumulator the other 21 % (profiles/r02_fused_mlp_fwd_stall_sites.txt) - the chunk's
// fc1 cannot start before the previous chunk has been read out of TMEM, its fc2 not before the whole GELU tile is
// written.  Here a chunk is 128 fc1 columns, the fc1 accumulator (2 x 128 TMEM columns) and the GELU tile
// (2 x 32 KB, fp16, K-major 128B-swizzled) are double buffered: fc1(c+1) and fc2(c-1) run while the warps work on
// chunk c, and the warps walk from chunk to chunk without waiting as long as the tensor pipe keeps up.
//
//   MMA1(c): acc1[b] [128 x 128] = xn[128 x 256] * W1[c*128 .., :]^T      16 tcgen05.mma (N = 128)
//   EPI(c) : the eight warps of group b = c & 1, one row x 64 columns per thread: + b1, GELU (packed half, tanh form;
//            2 gelu - the 1/2 is applied to the fc2 accumulator), fp16 tile -> buffer b.  The two groups work one chunk
//            period out of phase.
//   MMA2(c): acc2 [128 x 256] += tile[b] [128 x 128] * W2[:, c*128 ..]^T   8 tcgen05.mma (N = 256), TMEM cols 256..511
// Issue order MMA1(c+1), MMA2(c), ...  Weights stream from L2 through a ring of three 32 KB slots (a slot = two
// [128 x 64] k-blocks of W1 or one [256 x 64] k-block of W2).  The final epilogue is k_tc_mlp2.cu's (TMA panels,
// residual prefetched, LayerNorm through a TMEM stash); a warp's panel buffer is a 4 KB slice of the (then idle) GELU
// buffers, shared with its lane-quarter siblings' GELU rows - the quarter's four warps meet at a named barrier
// before any of them writes the next tile's GELU values.
//
//   warp 0: TMA producer     warp 1: TMEM allocator + MMA issuer     warps 2..17: GELU / final epilogue
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "tc.cuh"
#include "tc_epilogue.cuh"

namespace vit3d {

using namespace ptx;

constexpr int M3_H = 256;
constexpr int M3_NC = 128;                      // fc1 columns per chunk
constexpr int M3_X_BYTES = 128 * M3_H * 2;      // 65536: xn tile, 4 k-blocks of [128 x 64]
constexpr int M3_A_BYTES = 128 * M3_NC * 2;     // 32768: GELU tile, 2 k-blocks of [128 x 64]
constexpr int M3_SLOT = 32 * 1024;
constexpr int M3_NST = 3;
constexpr int M3_THREADS = 64 + 512;
constexpr int M3_SMEM = M3_X_BYTES + 2 * M3_A_BYTES + M3_NST * M3_SLOT + 16 * 128 /*b1 slices*/ + 512 /*barriers*/;
static_assert(M3_SMEM <= 232448, "over the 227 KB shared-memory limit");

struct Mlp3Args {
  const float* b1 = nullptr;      // [d]
  const float* b2 = nullptr;      // [H]
  const float* gamma = nullptr;   // [H] (LN)
  const float* beta = nullptr;    // [H] (LN)
  float eps = 1e-6f;
  int M = 0, d = 0;
};

// 2 gelu(x) of a packed-half pair (tanh form, see k_tc_mlp2.cu gelu_h2)
__device__ __forceinline__ uint32_t gelu2_h2(__half2 x) {
  const __half2 x2 = __hmul2(x, x);
  const __half2 p = __hfma2(x2, __float2half2_rn(0.0356774081f), __float2half2_rn(0.797884561f));
  const __half2 u = __hmul2(x, p);
  uint32_t ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 th = *reinterpret_cast<const __half2*>(&ti);
  const __half2 y = __hfma2(x, th, x);
  return *reinterpret_cast<const uint32_t*>(&y);
}

template <int LN>
__global__ void __launch_bounds__(M3_THREADS, 1)
tc_mlp3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmLn, Mlp3Args args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_a = smem + M3_X_BYTES;                              // two GELU buffers, 32 KB each
  uint8_t* s_w = s_a + 2 * M3_A_BYTES;
  uint8_t* s_b1 = s_w + M3_NST * M3_SLOT;                        // 16 x 128 B: packed-half b1 slice per warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1 + 16 * 128);
  uint64_t* x_full = bars;             // xn tile landed
  uint64_t* x_empty = bars + 1;        // (commit) the last fc1 of the tile has read xn
  uint64_t* w_full = bars + 2;         // [3]
  uint64_t* w_empty = bars + 5;        // [3] (commit)
  uint64_t* acc1_full = bars + 8;      // [2] (commit)
  uint64_t* acc1_empty = bars + 10;    // [2] 8 arrivals (warp group b)
  uint64_t* a2_full = bars + 12;       // [2] 8 arrivals
  uint64_t* a2_empty = bars + 14;      // [2] (commit)
  uint64_t* acc2_full = bars + 16;     // (commit)
  uint64_t* acc2_empty = bars + 17;    // 16 arrivals
  uint64_t* xpanel_free = bars + 18;   // 16: the final epilogue no longer uses the xn region
  uint64_t* res_bar = bars + 19;       // [16] per warp: residual panel 0 landed
  uint64_t* res_bar1 = bars + 35;      // [16] per warp: residual panel 1 landed (buffer = a 4 KB slice of the xn tile)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 51);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int M = args.M, d = args.d;
  const int nch = d / M3_NC;
  const int tiles = (M + 127) / 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < M3_NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 8);        // the eight warps of group b
      mbar_init(&a2_full[b], 8);
      mbar_init(&a2_empty[b], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    mbar_init(xpanel_free, 16);
    for (int w = 0; w < 16; ++w) { mbar_init(&res_bar[w], 1); mbar_init(&res_bar1[w], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      auto slot = [&]() -> uint8_t* {
        mbar_wait(&w_empty[stage], wphase ^ 1);
        mbar_arrive_expect_tx(&w_full[stage], M3_SLOT);
        return s_w + stage * M3_SLOT;
      };
      auto advance = [&]() { if (++stage == M3_NST) { stage = 0; wphase ^= 1; } };
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        mbar_wait(x_empty, (it & 1) ^ 1);           // the last fc1 of the previous tile has read xn ...
        mbar_wait(xpanel_free, (it & 1) ^ 1);       // ... and its final epilogue is done with the region
        mbar_arrive_expect_tx(x_full, M3_X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(s_x + kb * 16384, &tmX, x_full, kb * 64, t * 128);
        for (int s = 0; s <= nch; ++s) {
          if (s < nch) {
            for (int j = 0; j < 2; ++j) {                    // W1 rows s*128.., k-blocks 2j, 2j+1
              uint8_t* dst = slot();
              tma_load_2d(dst, &tmW1, &w_full[stage], (2 * j) * 64, s * M3_NC);
              tma_load_2d(dst + 16384, &tmW1, &w_full[stage], (2 * j + 1) * 64, s * M3_NC);
              advance();
            }
          }
          if (s >= 1) {
            for (int j = 0; j < 2; ++j) {                    // W2 [256 rows, cols (s-1)*128 + j*64 ..]
              uint8_t* dst = slot();
              tma_load_2d(dst, &tmW2, &w_full[stage], (s - 1) * M3_NC + j * 64, 0);
              advance();
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected thread walks the schedule)
    constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 128, M3_NC, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_ab(UMMA_FMT_F16, UMMA_FMT_F16, 128, M3_H);   // fp16 GELU tile x fp16 W2
    const uint64_t xdesc0 = make_smem_desc(smem_u32(s_x), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t adesc0 = make_smem_desc(smem_u32(s_a), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t wdesc0 = make_smem_desc(smem_u32(s_w), 16, 1024, UMMA_LAYOUT_SW128);
    if (elect_one()) {
      int stage = 0;
      uint32_t wphase = 0;
      uint32_t g1 = 0, g2 = 0;
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        mbar_wait(x_full, it & 1);
        for (int s = 0; s <= nch; ++s) {
          if (s < nch) {
            // ---- fc1(s): acc1[b] = xn * W1[s]^T
            const uint32_t b = g1 & 1;
            mbar_wait(&acc1_empty[b], ((g1 >> 1) & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (M3_SLOT >> 4));
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t ad = xdesc0 + (uint64_t)((2 * j + kk) * 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma<false>(tmem_base + b * M3_NC, ad + 2 * k, bd + kk * 1024 + 2 * k, idesc1, (j | kk | k) ? 1u : 0u);
              }
              umma_commit(&w_empty[stage]);
              if (++stage == M3_NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(&acc1_full[b]);
            if (s == nch - 1) umma_commit(x_empty);
            ++g1;
          }
          if (s >= 1) {
            // ---- fc2(s-1): acc2 += GELU tile[b] * W2[:, s-1]^T
            const int c = s - 1;
            const uint32_t b = g2 & 1;
            if (c == 0) mbar_wait(acc2_empty, (it & 1) ^ 1);
            mbar_wait(&a2_full[b], (g2 >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t ad = adesc0 + (uint64_t)(b * (M3_A_BYTES >> 4) + j * 1024);
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (M3_SLOT >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma<false>(tmem_base + 256, ad + 2 * k, bd + 2 * k, idesc2, (c | j | k) ? 1u : 0u);
              umma_commit(&w_empty[stage]);
              if (++stage == M3_NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(&a2_empty[b]);
            if (c == nch - 1) umma_commit(acc2_full);
            ++g2;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================== GELU / final epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                     // TMEM lane quarter
    const int part = ew >> 2;                   // 32-column slice of the 128-column chunk
    const int row = q * 32 + lane;
    const int sw7 = row & 7;
    uint8_t* b1_slot = s_b1 + ew * 128;
    const uint32_t b1_s = smem_u32(b1_slot);
    // final-epilogue panel buffer: rows q*32.. of k-block (part >> 1) in GELU buffer (part & 1): 4 KB, private to this
    // warp during the final epilogue (its GELU rows belong to this warp and its quarter sibling part ^ 1)
    auto panel = [&](int p_) -> uint8_t* { return s_a + (p_ & 1) * M3_A_BYTES + (p_ >> 1) * 16384 + q * 4096; };
    uint8_t* buf_ptr = panel(part);
    const uint32_t buf_s = smem_u32(buf_ptr);
    const uint32_t my_row = buf_s + lane * 128;
    uint64_t* rbar = &res_bar[ew];
    // second panel buffer: a 4 KB slice of the xn tile, idle from the last fc1 of a tile until the next tile's xn lands
    uint8_t* bufx_ptr = s_x + (part * 4 + q) * 4096;
    const uint32_t bufx_s = smem_u32(bufx_ptr);
    const uint32_t my_rowx = bufx_s + lane * 128;
    uint64_t* rbar1 = &res_bar1[ew];
    uint32_t rphase = 0;
    uint32_t g = 0;
    int it = 0;
    // GELU work split: the chunks of accumulator / GELU buffer b belong to warp group b (ew >> 3): eight warps, two per
    // lane quarter, 64 columns per thread.  The groups run one chunk period out of phase, so one group's MUFU burst
    // overlaps the other's TMEM read-out / shared-memory stores (the special-function pipe and the half2 pipe are
    // both ~2 k cycles per 256 columns and SM sub-partition: tools/mufu_bench.cu)
    const int grp = ew >> 3;
    const int hlf = (ew >> 2) & 1;              // 64-column half of the 128-column chunk = k-block `hlf` of the GELU buffer
    const uint32_t g_row = smem_u32(s_a) + grp * M3_A_BYTES + hlf * 16384 + row * 128;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m_base = t * 128 + q * 32;
      for (int c = grp; c < nch; c += 2, ++g) {
        const uint32_t ph = g & 1;              // this group's buffer completes one phase per own chunk
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 16) bv = __ldg(reinterpret_cast<const float4*>(args.b1 + c * M3_NC + hlf * 64) + lane);
        mbar_wait(&acc1_full[grp], ph);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + grp * M3_NC + hlf * 64;
        tmem_ld_32x32b_x32(tcol, r0);
        tmem_ld_32x32b_x32(tcol + 32, r1);
        if (lane < 16) {
          const __half2 h0 = __floats2half2_rn(bv.x, bv.y), h1 = __floats2half2_rn(bv.z, bv.w);
          *reinterpret_cast<uint2*>(b1_slot + lane * 8) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[grp]);  // fc1 of chunk c+2 may overwrite this accumulator
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const uint4 ba = ld_shared_v4(b1_s + j * 4);            // halves 2j .. 2j+7
          const uint4 bb = ld_shared_v4(b1_s + 64 + j * 4);
          const uint32_t a4[4] = {ba.x, ba.y, ba.z, ba.w}, b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 x = __floats2half2_rn(__uint_as_float(r0[2 * (j + i)]), __uint_as_float(r0[2 * (j + i) + 1]));
            pk[j + i] = gelu2_h2(__hadd2(x, *reinterpret_cast<const __half2*>(&a4[i])));
            const __half2 z = __floats2half2_rn(__uint_as_float(r1[2 * (j + i)]), __uint_as_float(r1[2 * (j + i) + 1]));
            pk[16 + j + i] = gelu2_h2(__hadd2(z, *reinterpret_cast<const __half2*>(&b4[i])));
          }
        }
        mbar_wait(&a2_empty[grp], ph ^ 1);  // fc2 of chunk c-2 has read this buffer
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(g_row + ((j ^ sw7) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[grp]);
      }
      // residual panel 1 goes into the xn region once every fc1 of the tile has completed (both accumulators' last
      // chunks: acc1_full[1] of chunk nch-1 was waited for by group 1, group 0 waits for it here)
      if (grp == 0) mbar_wait(&acc1_full[1], (g + 1) & 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(rbar1, 4096);
        tma_load_2d(bufx_ptr, &tmRes, rbar1, part * 64 + 32, m_base);
      }
      // ---------------- final epilogue of the tile: y = acc2 / 2 + b2 + residual (+ LayerNorm)
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      if (lane == 0) {
        mbar_arrive_expect_tx(rbar, 4096);
        tma_load_2d(buf_ptr, &tmRes, rbar, part * 64, m_base);
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + part * 64;
      const int sl7 = lane & 7;
      float s1 = 0.f, s2 = 0.f;
      // panel 1 first (its residual has been in the xn-region buffer for a while), then panel 0 (fetched above)
#pragma unroll
      for (int pi = 0; pi < 2; ++pi) {
        const int p = 1 - pi;
        const uint32_t prow = p ? my_rowx : my_row;
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + p * 32, r);
        mbar_wait(p ? rbar1 : rbar, rphase);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = prow + ((j ^ sl7) << 4);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(args.b2 + part * 64 + p * 32) + j);
          const uint4 x = ld_shared_v4(a);
          const float v0 = fmaf(__uint_as_float(r[4 * j]), 0.5f, bb.x) + __uint_as_float(x.x);
          const float v1 = fmaf(__uint_as_float(r[4 * j + 1]), 0.5f, bb.y) + __uint_as_float(x.y);
          const float v2 = fmaf(__uint_as_float(r[4 * j + 2]), 0.5f, bb.z) + __uint_as_float(x.z);
          const float v3 = fmaf(__uint_as_float(r[4 * j + 3]), 0.5f, bb.w) + __uint_as_float(x.w);
          if constexpr (LN != 2)
            st_shared_v4(a, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          if constexpr (LN != 0) {
            s1 += (v0 + v1) + (v2 + v3);
            s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
            r[4 * j] = __float_as_uint(v0); r[4 * j + 1] = __float_as_uint(v1);
            r[4 * j + 2] = __float_as_uint(v2); r[4 * j + 3] = __float_as_uint(v3);
          }
        }
        if constexpr (LN != 0) tmem_st_32x32b_x32(taddr + p * 32, r);
        if constexpr (LN != 2) fence_proxy_async_smem();
        __syncwarp();
        if constexpr (LN != 2) {
          if (lane == 0) {
            tma_store_2d(&tmY, p ? bufx_s : buf_s, part * 64 + p * 32, m_base);
            bulk_store_commit();
          }
        }
      }
      rphase ^= 1;
      if (lane == 0) {
        if constexpr (LN != 2) bulk_store_wait_read();   // both result panels have been read out of shared memory
        mbar_arrive(xpanel_free);                     // the producer may load the next tile's xn
      }
      __syncwarp();
      if constexpr (LN != 0) {
        tmem_st_wait();
        *reinterpret_cast<float2*>(buf_ptr + lane * 8) = make_float2(s1, s2);
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(128) : "memory");
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const float2 s = *reinterpret_cast<const float2*>(panel(pp) + lane * 8);
          t1 += s.x; t2 += s.y;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(128) : "memory");
        const float mean = t1 * (1.0f / M3_H);
        const float rstd = rsqrtf(fmaxf(t2 * (1.0f / M3_H) - mean * mean, 0.f) + args.eps);
        const float shift = -mean * rstd;
        if constexpr (LN == 2) {
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + p * 32, r);
            tmem_ld_wait();
            if (p == 1) {
              if (lane == 0) bulk_store_wait_read();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 ga = __ldg(reinterpret_cast<const float4*>(args.gamma + part * 64 + p * 32) + j);
              const float4 be = __ldg(reinterpret_cast<const float4*>(args.beta + part * 64 + p * 32) + j);
              const float y0 = fmaf(fmaf(__uint_as_float(r[4 * j]), rstd, shift), ga.x, be.x);
              const float y1 = fmaf(fmaf(__uint_as_float(r[4 * j + 1]), rstd, shift), ga.y, be.y);
              const float y2 = fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rstd, shift), ga.z, be.z);
              const float y3 = fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rstd, shift), ga.w, be.w);
              st_shared_v4(my_row + ((j ^ sl7) << 4), __float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2),
                           __float_as_uint(y3));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmY, buf_s, part * 64 + p * 32, m_base);
              bulk_store_commit();
            }
          }
          if (lane == 0) bulk_store_wait_read();
          __syncwarp();
        } else {
          const int sw3 = (lane >> 1) & 3;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + p * 32, r);
            tmem_ld_wait();
            const uint32_t prow = buf_s + p * 2048 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t wv[4];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 ga = __ldg(reinterpret_cast<const float4*>(args.gamma + part * 64 + p * 32) + 2 * j + h);
                const float4 be = __ldg(reinterpret_cast<const float4*>(args.beta + part * 64 + p * 32) + 2 * j + h);
                const float y0 = fmaf(fmaf(__uint_as_float(r[8 * j + 4 * h]), rstd, shift), ga.x, be.x);
                const float y1 = fmaf(fmaf(__uint_as_float(r[8 * j + 4 * h + 1]), rstd, shift), ga.y, be.y);
                const float y2 = fmaf(fmaf(__uint_as_float(r[8 * j + 4 * h + 2]), rstd, shift), ga.z, be.z);
                const float y3 = fmaf(fmaf(__uint_as_float(r[8 * j + 4 * h + 3]), rstd, shift), ga.w, be.w);
                wv[2 * h] = pack2_bf16(y0, y1);
                wv[2 * h + 1] = pack2_bf16(y2, y3);
              }
              st_shared_v4(prow + ((j ^ sw3) << 4), wv[0], wv[1], wv[2], wv[3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmLn, buf_s, part * 64, m_base);
            tma_store_2d(&tmLn, buf_s + 2048, part * 64 + 32, m_base);
            bulk_store_commit();
            bulk_store_wait_read();
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);         // fc2 of the next tile may overwrite acc2
      // the quarter's four warps share GELU rows with each other's panel buffers: nobody writes the next tile's GELU
      // values before all four have left their panels
      asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(128) : "memory");
    }
    if (lane == 0) bulk_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool tc_mlp3_supported(int M, int H, int d) { return M > 0 && H == M3_H && d % (2 * M3_NC) == 0 && d >= 2 * M3_NC; }

template <int LN>
static int launch_mlp3(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& ty,
                       const CUtensorMap& tr, const CUtensorMap& tl, const Mlp3Args& a, cudaStream_t st) {
  auto kern = tc_mlp3_kernel<LN>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, M3_SMEM));
    configured_dev = dev;
  }
  const int tiles = ceil_div(a.M, 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  V3_CUDA(launch_pdl(kern, dim3(grid), dim3(M3_THREADS), (size_t)M3_SMEM, st, tx, tw1, tw2, ty, tr, tl, a));
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// same contract as tc_mlp2 (k_tc_mlp2.cu): y / ln_out semantics per ln_f32
int tc_mlp3(const void* xn, const void* w1, const float* b1, const void* w2_h, const float* b2, const float* residual,
            float* y, const float* gamma, const float* beta, float eps, void* ln_out, int ln_f32, int M, int H, int d,
            cudaStream_t st) {
  if (!tc_mlp3_supported(M, H, d)) V3_UNSUPPORTED("fused MLP (128-column chunks): unsupported shape M=%d H=%d d=%d", M, H, d);
  CUtensorMap tx, tw1, tw2, ty, tr, tl;
  int rc = make_tmap_2d(&tx, xn, 2, M, H, H, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw1, w1, 2, d, H, H, 128, 64, 128);          // W1 [d, 256]: 128 chunk rows x 64 K per box
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw2, w2_h, 2, H, d, d, 256, 64, 128);        // W2 [256, d] (fp16 bits): 256 rows x 64 K per box
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&ty, ln_f32 ? ln_out : (void*)y, 4, M, H, H, 32, 32, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tr, residual, 4, M, H, H, 32, 32, 128);
  if (rc != VIT3D_OK) return rc;
  tl = ty;
  if (ln_out && !ln_f32) {
    rc = make_tmap_2d(&tl, ln_out, 2, M, H, H, 32, 32, 64);
    if (rc != VIT3D_OK) return rc;
  }
  Mlp3Args a;
  a.b1 = b1; a.b2 = b2; a.gamma = gamma; a.beta = beta; a.eps = eps; a.M = M; a.d = d;
  if (!ln_out) return launch_mlp3<0>(tx, tw1, tw2, ty, tr, tl, a, st);
  if (ln_f32) return launch_mlp3<2>(tx, tw1, tw2, ty, tr, tl, a, st);
  return launch_mlp3<1>(tx, tw1, tw2, ty, tr, tl, a, st);
}

}  // namespace vit3d
