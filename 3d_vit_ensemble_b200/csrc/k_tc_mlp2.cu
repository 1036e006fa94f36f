// Fused MLP block (a3 + a4, modeling.py:118-124, :194-196; inference):
//
//     y  = x + fc2(GELU(fc1(xn) + b1)) + b2        (fp32 residual stream, y may alias x)
//     yn = LayerNorm(y) * gamma + beta              (bf16, optional: the next Block's attention_norm)
//
// in ONE kernel.  The [M, d] GELU intermediate (d = 2048 / 3072: 2 x 272 MB of HBM traffic per layer at
// batch 1024, more than everything else in the layer together) never leaves the SM.
//
// Per 128-row tile (CTA) the d fc1 columns are processed in chunks of 256:
//   fc1(c): acc1[128 x 256] = xn[128 x 256] * W1[c*256 .., :]^T     16 tcgen05.mma (N = 256, K = 16), TMEM cols 0..255
//   GELU  : 16 epilogue warps read acc1 (tcgen05.ld, ~400 cycles for the whole tile), add b1 in packed half,
//           apply the fitted tanh-form GELU in packed half and write the fp16 tile A2[128 x 256] into shared
//           memory in the K-major 128B-swizzled layout tcgen05 reads (64 KB, single buffer)
//   fc2(c): acc2[128 x 256] += A2 * W2[:, c*256 ..]^T               16 tcgen05.mma, TMEM cols 256..511, fp16 x fp16
// The issue order fc1(c), fc2(c-1), fc1(c+1), fc2(c), ... gives the GELU of a chunk the duration of two MMA
// groups (4096 cycles at full rate; it needs ~2200 issue slots per scheduler), so the tensor pipe never waits
// for it in steady state.  Weights stream from L2 through a TMA ring (96 KB).
//
// MC = true: the kernel runs on clusters of two CTAs that stay independent (own tile, own tensor core, own
// barriers) except for the weight ring: every k-block of W1 / W2 is fetched ONCE per cluster - each CTA loads
// half of its 256 rows and multicasts it into both CTAs' rings - and a ring slot is refilled when BOTH CTAs'
// MMAs have released it (multicast tcgen05.commit, barrier count 2).  A single CTA needs 2 MB of weights per
// 128-row tile = 64 B/clk at full MMA rate, more than the ~42 B/clk/SM the L2 delivers when every SM streams;
// sharing halves it.  (A cta_group::2 version was measured slower: its remote release-arrives from 32 epilogue
// warps per chunk cost more than the bandwidth they saved.)
//
// The final epilogue (once per tile) is the TMA-panel epilogue of k_tc_gemm_res.cu: residual panels fetched by
// bulk tensor loads, result panels stored by bulk tensor stores, LayerNorm through a TMEM stash; its panel
// buffers alias the warp's own 4 KB slice of the (then idle) GELU tile.
//
//   warp 0: TMA producer     warp 1: TMEM allocator + MMA issuer     warps 2..17: GELU / final epilogue
#include "ptx.cuh"
#include "tc.cuh"
#include "tc_epilogue.cuh"

namespace vit3d {

using namespace ptx;

constexpr int M2_H = 256;                       // hidden size: K of fc1, N of fc2
constexpr int M2_NC = 256;                      // fc1 columns per chunk
constexpr int M2_X_BYTES = 128 * M2_H * 2;      // 65536: xn tile, 4 k-blocks of [128 x 64]
constexpr int M2_A_BYTES = 128 * M2_NC * 2;     // 65536: GELU tile, 4 k-blocks of [128 x 64]
constexpr int M2_THREADS = 64 + 512;
constexpr int M2_RING_BYTES = 96 * 1024;
constexpr int M2_SMEM = M2_X_BYTES + M2_A_BYTES + M2_RING_BYTES + 16 * 128 /*b1 slices*/ + 512 /*barriers*/;
static_assert(M2_SMEM <= 232448, "over the 227 KB shared-memory limit");

struct Mlp2Args {
  const float* b1 = nullptr;      // [d]
  const float* b2 = nullptr;      // [H]
  const float* gamma = nullptr;   // [H] (LN)
  const float* beta = nullptr;    // [H] (LN)
  float eps = 1e-6f;
  int M = 0, d = 0;
};

// GELU of a packed-half pair, result packed fp16 (the A operand of fc2), tanh form
//     y = hx + hx * tanh(x * (c0 + c1 x^2)),  hx = x / 2,  c0 = sqrt(2/pi), c1 = 0.044715 c0
// = 7 HFMA2-pipe ops + 2 MUFU + 1 PRMT per pair.  The argument is monotone, so no clamp is needed (x^2
// overflowing to +inf in half precision gives tanh(+-inf) = +-1, the correct limit).  Against the exact erf
// GELU of the reference (modeling.py:52) the form itself is off by <= 4.7e-4 absolute; measured on the logits
// of conf 5 / conf 18 that is 2-4e-4, below the 7.6e-4 the fp16 rounding of the GELU tile contributes and far
// inside the 2e-2 bf16-mode tolerance (the training path keeps the 3-term fit with 3e-5).
__device__ __forceinline__ uint32_t gelu_h2(__half2 x) {
  const __half2 x2 = __hmul2(x, x);
  const __half2 p = __hfma2(x2, __float2half2_rn(0.0356774081f), __float2half2_rn(0.797884561f));
  const __half2 u = __hmul2(x, p);
  uint32_t ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 th = *reinterpret_cast<const __half2*>(&ti);
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 y = __hfma2(hx, th, hx);
  return *reinterpret_cast<const uint32_t*>(&y);
}

// LN: 0 = y only; 1 = y and bf16 LayerNorm(y) (the next Block's attention_norm); 2 = fp32 LayerNorm(y) ONLY
// (encoder_norm after the last Block, modeling.py:253: the block output itself is not needed at inference, so
// the kernel writes the normalised rows through tmY and the separate LayerNorm pass - 68 MB in, 68 MB out at
// batch 1024 - disappears)
template <bool MC, int LN>
__global__ void __cluster_dims__(MC ? 2 : 1, 1, 1) __launch_bounds__(M2_THREADS, 1)
tc_mlp2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmLn, Mlp2Args args) {
  constexpr int W_ROWS = MC ? 128 : 256;            // B-operand rows this CTA loads per k-block
  constexpr int W_BYTES = 256 * 128;                // 32 KB per ring slot (both halves)
  constexpr int NST = M2_RING_BYTES / W_BYTES;      // 3
  constexpr int NCTA = MC ? 2 : 1;
  constexpr uint32_t EPI_ARRIVALS = 16;             // one elected arrival per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_a = smem + M2_X_BYTES;
  uint8_t* s_w = s_a + M2_A_BYTES;
  uint8_t* s_b1 = s_w + M2_RING_BYTES;                           // 16 x 128 B: packed-half b1 slice per warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1 + 16 * 128);
  uint64_t* x_full = bars;             // xn tile landed
  uint64_t* x_empty = bars + 1;        // (commit) the last fc1 of the tile has read xn
  uint64_t* w_full = bars + 2;         // [NST] weight k-block landed (both halves)
  uint64_t* w_empty = bars + 8;        // [NST] (commit of every CTA of the cluster) slot may be refilled
  uint64_t* acc1_full = bars + 14;     // (commit)
  uint64_t* acc1_empty = bars + 15;    // EPI_ARRIVALS
  uint64_t* a2_full = bars + 16;       // EPI_ARRIVALS
  uint64_t* a2_empty = bars + 17;      // (commit)
  uint64_t* acc2_full = bars + 18;     // (commit)
  uint64_t* acc2_empty = bars + 19;    // EPI_ARRIVALS
  uint64_t* res_bar = bars + 20;       // [16] per warp: residual panel 0 landed (buffer = the warp's GELU-tile slice)
  uint64_t* res_bar1 = bars + 36;      // [16] per warp: residual panel 1 landed (buffer = a 4 KB slice of the xn tile)
  uint64_t* xpanel_free = bars + 52;   // 16: the final epilogue no longer uses the xn region
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 53);

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the role branches below are uniform
  // control flow and the MMA / TMA warps keep their operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = MC ? cluster_ctarank() : 0u;
  const int M = args.M, d = args.d;
  const int nch = d / M2_NC;
  const int tiles = (M + 128 * NCTA - 1) / (128 * NCTA);         // tiles of 128 (256 for a pair) rows
  const int ngroups = gridDim.x / NCTA;
  const int gid = blockIdx.x / NCTA;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], NCTA); }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, EPI_ARRIVALS);
    mbar_init(a2_full, EPI_ARRIVALS);
    mbar_init(a2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EPI_ARRIVALS);
    for (int w = 0; w < 16; ++w) { mbar_init(&res_bar[w], 1); mbar_init(&res_bar1[w], 1); }
    mbar_init(xpanel_free, 16);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();       // both CTAs' barriers exist before any multicast signal
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
      for (int t = gid; t < tiles; t += ngroups, ++it) {
        mbar_wait(x_empty, (it & 1) ^ 1);           // the last fc1 of the previous tile has read xn ...
        mbar_wait(xpanel_free, (it & 1) ^ 1);       // ... and its final epilogue is done with the region
        mbar_arrive_expect_tx(x_full, M2_X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(s_x + kb * 16384, &tmX, x_full, kb * 64, (t * NCTA + (int)rank) * 128);
        for (int s = 0; s <= nch; ++s) {
          for (int which = 0; which < 2; ++which) {
            const bool is_w1 = which == 0;
            const int c = is_w1 ? s : s - 1;
            if (is_w1 ? (s >= nch) : (s < 1)) continue;
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(&w_empty[stage], wphase ^ 1);          // every CTA of the cluster has released the slot
              mbar_arrive_expect_tx(&w_full[stage], W_BYTES);
              uint8_t* dst = s_w + stage * W_BYTES + (int)rank * (W_ROWS * 128);   // this CTA's half of the rows
              const int c0 = is_w1 ? kb * 64 : c * M2_NC + kb * 64;
              const int c1 = is_w1 ? c * M2_NC + (int)rank * W_ROWS : (int)rank * W_ROWS;
              const CUtensorMap* tw = is_w1 ? &tmW1 : &tmW2;
              if constexpr (MC) tma_load_2d_mcast(dst, tw, &w_full[stage], c0, c1, (uint16_t)3);
              else tma_load_2d(dst, tw, &w_full[stage], c0, c1);
              if (++stage == NST) { stage = 0; wphase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer: ONE thread, elected once, walks the whole
    // schedule (barrier waits, tcgen05.mma, commits).  tools/mma_pipe_bench.cu: with an election and a __syncwarp
    // per k-block (all lanes walking the loop) a 4-MMA stage costs 186-193 cycles per MMA - the warp-level
    // election / reconvergence sits between the last MMA of one stage and the first of the next and the issue
    // queue is only ~2 MMAs deep - while a single elected thread running the same hand-shake reaches 128.1.
    constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 128, M2_NC, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_ab(UMMA_FMT_F16, UMMA_FMT_F16, 128, M2_H);   // fp16 GELU tile x fp16 W2
    const uint64_t xdesc0 = make_smem_desc(smem_u32(s_x), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t adesc0 = make_smem_desc(smem_u32(s_a), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t wdesc0 = make_smem_desc(smem_u32(s_w), 16, 1024, UMMA_LAYOUT_SW128);
    auto release_slot = [&](uint64_t* bar) {        // ring slot consumed: tell every producer of the cluster
      if constexpr (MC) umma_commit_mcast(bar, (uint16_t)3); else umma_commit(bar);
    };
    if (elect_one()) {
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      uint32_t g1 = 0, g2 = 0;              // chunk counters of fc1 / fc2 (phases of the per-chunk barriers)
      for (int t = gid; t < tiles; t += ngroups, ++it) {
        mbar_wait(x_full, it & 1);
        for (int s = 0; s <= nch; ++s) {
          if (s < nch) {
            // ---- fc1(s): acc1 = xn * W1[s]^T
            mbar_wait(acc1_empty, (g1 & 1) ^ 1);          // the GELU warps have read the previous chunk out
            tc_fence_after();
            // no tcgen05 fence after a weight k-block has landed: the mbarrier's complete_tx already orders the TMA
            // writes before the MMAs' reads (the fence above pairs with the epilogue warps' tcgen05.ld)
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t ad = xdesc0 + (uint64_t)(kb * 1024);                     // k-block kb: + 16 KB (>> 4)
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (W_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma<false>(tmem_base, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) ? 1u : 0u);
              release_slot(&w_empty[stage]);
              if (++stage == NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(acc1_full);
            if (s == nch - 1) umma_commit(x_empty);      // xn may be replaced by the next tile's
            ++g1;
          }
          if (s >= 1) {
            // ---- fc2(s-1): acc2 += GELU tile * W2[:, s-1]^T
            const int c = s - 1;
            if (c == 0) mbar_wait(acc2_empty, (it & 1) ^ 1);   // previous tile's final epilogue has drained acc2
            mbar_wait(a2_full, g2 & 1);
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t ad = adesc0 + (uint64_t)(kb * 1024);
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (W_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma<false>(tmem_base + 256, ad + 2 * k, bd + 2 * k, idesc2, (c | kb | k) ? 1u : 0u);
              release_slot(&w_empty[stage]);
              if (++stage == NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(a2_empty);
            if (c == nch - 1) umma_commit(acc2_full);
            ++g2;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================== GELU / final epilogue warps (own 128 rows)
    const int q = warp & 3;                     // TMEM lane quarter
    const int part = (warp - 2) >> 2;           // 64-column slice of a chunk = k-block `part` of the GELU tile
    const int row = q * 32 + lane;
    const uint32_t a_row = smem_u32(s_a) + part * 16384 + row * 128;
    const int sw7 = row & 7;
    uint8_t* b1_slot = s_b1 + (warp - 2) * 128;
    const uint32_t b1_s = smem_u32(b1_slot);
    // final-epilogue panel buffer = this warp's own 4 KB slice of the GELU tile (rows q*32.., k-block part)
    uint8_t* buf_ptr = s_a + part * 16384 + q * 4096;
    const uint32_t buf_s = smem_u32(buf_ptr);
    const uint32_t my_row = buf_s + lane * 128;
    uint64_t* rbar = &res_bar[warp - 2];
    // second panel buffer: a 4 KB slice of the xn tile, idle from the last fc1 of a tile until the next tile's
    // xn is loaded - its residual panel is fetched two MMA groups before the final epilogue needs it
    uint8_t* bufx_ptr = s_x + (part * 4 + q) * 4096;
    const uint32_t bufx_s = smem_u32(bufx_ptr);
    const uint32_t my_rowx = bufx_s + lane * 128;
    uint64_t* rbar1 = &res_bar1[warp - 2];
    uint32_t rphase = 0;
    uint32_t g = 0;                             // chunk counter
    int it = 0;
    for (int t = gid; t < tiles; t += ngroups, ++it) {
      const int m_base = (t * NCTA + (int)rank) * 128 + q * 32;
      for (int c = 0; c < nch; ++c, ++g) {
        // this warp's 64 fc1 biases as packed halves (the global load overlaps the wait for the accumulator)
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 16) bv = __ldg(reinterpret_cast<const float4*>(args.b1 + c * M2_NC + part * 64) + lane);
        mbar_wait(acc1_full, g & 1);
        tc_fence_after();
        if (c == nch - 1 && lane == 0) {
          // every fc1 of this tile has completed: the xn region is free -> prefetch residual panel 1 into it
          mbar_arrive_expect_tx(rbar1, 4096);
          tma_load_2d(bufx_ptr, &tmRes, rbar1, part * 64 + 32, m_base);
        }
        uint32_t r0[32], r1[32];
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + part * 64;
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
        if (lane == 0) mbar_arrive(acc1_empty);       // fc1 of the next chunk may overwrite acc1
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const uint4 ba = ld_shared_v4(b1_s + j * 4);            // halves 2j .. 2j+7
          const uint4 bb = ld_shared_v4(b1_s + 64 + j * 4);
          const uint32_t a4[4] = {ba.x, ba.y, ba.z, ba.w}, b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __half2 x = __floats2half2_rn(__uint_as_float(r0[2 * (j + i)]), __uint_as_float(r0[2 * (j + i) + 1]));
            pk[j + i] = gelu_h2(__hadd2(x, *reinterpret_cast<const __half2*>(&a4[i])));
            __half2 z = __floats2half2_rn(__uint_as_float(r1[2 * (j + i)]), __uint_as_float(r1[2 * (j + i) + 1]));
            pk[16 + j + i] = gelu_h2(__hadd2(z, *reinterpret_cast<const __half2*>(&b4[i])));
          }
        }
        mbar_wait(a2_empty, (g & 1) ^ 1);   // fc2 of the previous chunk has read the GELU tile
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(a_row + ((j ^ sw7) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full);
      }
      // ---------------- final epilogue of the tile: y = acc2 + b2 + residual (+ LayerNorm)
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
          const float4 b = __ldg(reinterpret_cast<const float4*>(args.b2 + part * 64 + p * 32) + j);
          const uint4 x = ld_shared_v4(a);
          const float v0 = __uint_as_float(r[4 * j]) + b.x + __uint_as_float(x.x);
          const float v1 = __uint_as_float(r[4 * j + 1]) + b.y + __uint_as_float(x.y);
          const float v2 = __uint_as_float(r[4 * j + 2]) + b.z + __uint_as_float(x.z);
          const float v3 = __uint_as_float(r[4 * j + 3]) + b.w + __uint_as_float(x.w);
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
          const float2 s = *reinterpret_cast<const float2*>(s_a + pp * 16384 + q * 4096 + lane * 8);
          t1 += s.x; t2 += s.y;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(128) : "memory");
        const float mean = t1 * (1.0f / M2_H);
        const float rstd = rsqrtf(fmaxf(t2 * (1.0f / M2_H) - mean * mean, 0.f) + args.eps);
        const float shift = -mean * rstd;
        if constexpr (LN == 2) {
          // fp32 rows, one [32 x 32] panel at a time through the warp's 4 KB buffer (SWIZZLE_128B, as the y panels)
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
    }
    if (lane == 0) bulk_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();         // nobody leaves while the partner may still multicast into this CTA
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool tc_mlp2_supported(int M, int H, int d) { return M > 0 && H == M2_H && d % M2_NC == 0 && d >= M2_NC; }

template <bool MC, int LN>
static int launch_mlp2(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& ty,
                       const CUtensorMap& tr, const CUtensorMap& tl, const Mlp2Args& a, cudaStream_t st) {
  auto kern = tc_mlp2_kernel<MC, LN>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM));
    configured_dev = dev;
  }
  const int ncta = MC ? 2 : 1;
  const int tiles = ceil_div(a.M, 128 * ncta);
  int groups = sm_count() / ncta;
  if (groups > tiles) groups = tiles;
  V3_CUDA(launch_pdl(kern, dim3(groups * ncta), dim3(M2_THREADS), (size_t)M2_SMEM, st, tx, tw1, tw2, ty, tr, tl, a));
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// xn [M,256] bf16, w1 [d,256] bf16, w2 [256,d] FP16, residual / y [M,256] fp32 (may alias), ln_out [M,256] bf16 or null.
// ln_f32 != 0: ln_out is FP32 and is the only output (y is not written and may be null).
int tc_mlp2_fwd(const void* xn, const void* w1, const float* b1, const void* w2_h, const float* b2, const float* residual,
                float* y, const float* gamma, const float* beta, float eps, void* ln_out, int ln_f32, int M, int H, int d,
                cudaStream_t st) {
  if (!tc_mlp2_supported(M, H, d)) V3_UNSUPPORTED("fused MLP: unsupported shape M=%d H=%d d=%d", M, H, d);
  if (ln_f32 && !ln_out) { set_error("fused MLP: fp32 LayerNorm output requested without a buffer"); return VIT3D_ERR_INVALID; }
  const bool pair = M > 128 && tuning(VIT3D_TUNE_MLP_PAIR) != 0;
  CUtensorMap tx, tw1, tw2, ty, tr, tl;
  int rc = make_tmap_2d(&tx, xn, 2, M, H, H, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw1, w1, 2, d, H, H, pair ? 128 : 256, 64, 128);      // W1 [d, 256]: a CTA of a pair loads half the rows
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw2, w2_h, 2, H, d, d, pair ? 128 : 256, 64, 128);    // W2 [256, d] (fp16 bits)
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
  Mlp2Args a;
  a.b1 = b1; a.b2 = b2; a.gamma = gamma; a.beta = beta; a.eps = eps; a.M = M; a.d = d;
  const int mode = ln_f32 ? 2 : (ln_out ? 1 : 0);
  if (pair) {
    if (mode == 2) return launch_mlp2<true, 2>(tx, tw1, tw2, ty, tr, tl, a, st);
    return mode ? launch_mlp2<true, 1>(tx, tw1, tw2, ty, tr, tl, a, st) : launch_mlp2<true, 0>(tx, tw1, tw2, ty, tr, tl, a, st);
  }
  if (mode == 2) return launch_mlp2<false, 2>(tx, tw1, tw2, ty, tr, tl, a, st);
  return mode ? launch_mlp2<false, 1>(tx, tw1, tw2, ty, tr, tl, a, st) : launch_mlp2<false, 0>(tx, tw1, tw2, ty, tr, tl, a, st);
}

}  // namespace vit3d
