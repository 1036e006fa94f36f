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
  const __half2 y = __hfma2(x, th, x);        // 2 gelu(x): the factor 1/2 is applied to the fc2 accumulator (exact)
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
#include "tc_mlp2_epilogue.inc"
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();         // nobody leaves while the partner may still multicast into this CTA
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ============================================================================ CTA-pair version (cta_group::2)
// The same block on pairs of CTAs (one TPC): ONE tcgen05.mma of M = 256 covers the 128 rows of each CTA; every weight
// k-block is split between the two CTAs (each loads and holds 128 of its 256 rows, 16 KB instead of 32 KB), so a CTA's
// 96 KB ring holds SIX k-blocks - 3072 cycles of MMA work in flight instead of 1536.  That is what the one-CTA kernel
// lacks: its MMA thread spends most of its waits on w_full (ncu source view, profiles/r02_*), a ring slot comes back
// ~2500 cycles after its last MMA was issued, and halving the bytes per slot alone (timing experiment) changed nothing.
// Only the leader CTA (rank 0) issues MMAs and commits (multicast to both CTAs' barriers); both CTAs run their own TMA
// producer - whose loads signal the LEADER's full barriers - and their own 16 epilogue warps on their own rows.  The
// three signals that flow from the epilogue warps to the MMA thread (acc1 drained, GELU tile written, acc2 drained)
// stay CTA-local 16-arrival barriers; in the peer CTA the otherwise idle MMA-warp thread forwards each completion with
// ONE remote arrive (an earlier cta_group::2 attempt let all 32 epilogue warps arrive remotely - one cluster-scope
// release per warp and chunk - and lost more than it gained).
template <int LN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(M2_THREADS, 1)
tc_mlp2x_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmLn, Mlp2Args args) {
  constexpr int NCTA = 2;
  constexpr int W_SLOT = 128 * 128;                 // 16 KB: this CTA's 128 rows of one weight k-block
  constexpr int NST = M2_RING_BYTES / W_SLOT;       // 6
  constexpr uint32_t EPI_ARRIVALS = 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_a = smem + M2_X_BYTES;
  uint8_t* s_w = s_a + M2_A_BYTES;
  uint8_t* s_b1 = s_w + M2_RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1 + 16 * 128);
  uint64_t* x_full = bars;             // LEADER: both CTAs' xn tiles landed
  uint64_t* x_empty = bars + 1;        // (multicast commit) the last fc1 of the tile has read xn
  uint64_t* w_full = bars + 2;         // [NST] LEADER: both halves of the weight k-block landed
  uint64_t* w_empty = bars + 8;        // [NST] (multicast commit) slot may be refilled
  uint64_t* acc1_full = bars + 14;     // (multicast commit)
  uint64_t* acc1_empty = bars + 15;    // EPI_ARRIVALS, local
  uint64_t* a2_full = bars + 16;       // EPI_ARRIVALS, local
  uint64_t* a2_empty = bars + 17;      // (multicast commit)
  uint64_t* acc2_full = bars + 18;     // (multicast commit)
  uint64_t* acc2_empty = bars + 19;    // EPI_ARRIVALS, local
  uint64_t* res_bar = bars + 20;       // [16]
  uint64_t* res_bar1 = bars + 36;      // [16]
  uint64_t* xpanel_free = bars + 52;   // 16, local
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 53);
  uint64_t* p_acc1_empty = bars + 54;  // LEADER: 1 remote arrival per chunk (the peer's acc1_empty completed)
  uint64_t* p_a2_full = bars + 55;     // LEADER: 1 remote arrival per chunk
  uint64_t* p_acc2_empty = bars + 56;  // LEADER: 1 remote arrival per tile

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int M = args.M, d = args.d;
  const int nch = d / M2_NC;
  const int tiles = (M + 255) / 256;               // pairs of 128-row tiles
  const int ngroups = gridDim.x / NCTA;
  const int gid = blockIdx.x / NCTA;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, EPI_ARRIVALS);
    mbar_init(a2_full, EPI_ARRIVALS);
    mbar_init(a2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EPI_ARRIVALS);
    for (int w = 0; w < 16; ++w) { mbar_init(&res_bar[w], 1); mbar_init(&res_bar1[w], 1); }
    mbar_init(xpanel_free, 16);
    mbar_init(p_acc1_empty, 1);
    mbar_init(p_a2_full, 1);
    mbar_init(p_acc2_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // both CTAs' barriers and TMEM exist before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs; completion counted on the leader)
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0;
      int it = 0;
      const uint32_t lead_x_full = map_to_rank(x_full, 0);
      for (int t = gid; t < tiles; t += ngroups, ++it) {
        mbar_wait(x_empty, (it & 1) ^ 1);
        mbar_wait(xpanel_free, (it & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * M2_X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(s_x + kb * 16384, &tmX, lead_x_full, kb * 64, (t * NCTA + (int)rank) * 128);
        for (int s = 0; s <= nch; ++s) {
          for (int which = 0; which < 2; ++which) {
            const bool is_w1 = which == 0;
            const int c = is_w1 ? s : s - 1;
            if (is_w1 ? (s >= nch) : (s < 1)) continue;
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(&w_empty[stage], wphase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 2 * W_SLOT);
              const int c0 = is_w1 ? kb * 64 : c * M2_NC + kb * 64;
              const int c1 = is_w1 ? c * M2_NC + (int)rank * 128 : (int)rank * 128;       // this CTA's half of the rows
              tma_load_2d_pair(s_w + stage * W_SLOT, is_w1 ? &tmW1 : &tmW2, map_to_rank(&w_full[stage], 0), c0, c1);
              if (++stage == NST) { stage = 0; wphase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      if (rank == 0) {
        // ===================================================== MMA issuer (leader CTA): M = 256 over the pair
        constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 256, M2_NC, 0, 0);
        constexpr uint32_t idesc2 = make_idesc_ab(UMMA_FMT_F16, UMMA_FMT_F16, 256, M2_H);
        const uint64_t xdesc0 = make_smem_desc(smem_u32(s_x), 16, 1024, UMMA_LAYOUT_SW128);
        const uint64_t adesc0 = make_smem_desc(smem_u32(s_a), 16, 1024, UMMA_LAYOUT_SW128);
        const uint64_t wdesc0 = make_smem_desc(smem_u32(s_w), 16, 1024, UMMA_LAYOUT_SW128);
        int stage = 0;
        uint32_t wphase = 0;
        int it = 0;
        uint32_t g1 = 0, g2 = 0;
        for (int t = gid; t < tiles; t += ngroups, ++it) {
          mbar_wait(x_full, it & 1);
          for (int s = 0; s <= nch; ++s) {
            if (s < nch) {
              mbar_wait(acc1_empty, (g1 & 1) ^ 1);
              mbar_wait_cluster(p_acc1_empty, (g1 & 1) ^ 1);
              tc_fence_after();
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) {
                mbar_wait(&w_full[stage], wphase);
                const uint64_t ad = xdesc0 + (uint64_t)(kb * 1024);
                const uint64_t bd = wdesc0 + (uint64_t)(stage * (W_SLOT >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma2(tmem_base, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                umma2_commit(&w_empty[stage]);
                if (++stage == NST) { stage = 0; wphase ^= 1; }
              }
              umma2_commit(acc1_full);
              if (s == nch - 1) umma2_commit(x_empty);
              ++g1;
            }
            if (s >= 1) {
              const int c = s - 1;
              if (c == 0) {
                mbar_wait(acc2_empty, (it & 1) ^ 1);
                mbar_wait_cluster(p_acc2_empty, (it & 1) ^ 1);
              }
              mbar_wait(a2_full, g2 & 1);
              mbar_wait_cluster(p_a2_full, g2 & 1);
              tc_fence_after();
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) {
                mbar_wait(&w_full[stage], wphase);
                const uint64_t ad = adesc0 + (uint64_t)(kb * 1024);
                const uint64_t bd = wdesc0 + (uint64_t)(stage * (W_SLOT >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma2(tmem_base + 256, ad + 2 * k, bd + 2 * k, idesc2, (c | kb | k) ? 1u : 0u);
                umma2_commit(&w_empty[stage]);
                if (++stage == NST) { stage = 0; wphase ^= 1; }
              }
              umma2_commit(a2_empty);
              if (c == nch - 1) umma2_commit(acc2_full);
              ++g2;
            }
          }
        }
      } else {
        // ===================================================== relay (peer CTA): one remote arrive per local completion
        const uint32_t r_acc1 = map_to_rank(p_acc1_empty, 0), r_a2 = map_to_rank(p_a2_full, 0), r_acc2 = map_to_rank(p_acc2_empty, 0);
        uint32_t g = 0;
        int it = 0;
        for (int t = gid; t < tiles; t += ngroups, ++it) {
          for (int c = 0; c < nch; ++c, ++g) {
            mbar_wait(acc1_empty, g & 1);
            mbar_arrive_cluster(r_acc1);
            mbar_wait(a2_full, g & 1);
            mbar_arrive_cluster(r_a2);
          }
          mbar_wait(acc2_empty, it & 1);
          mbar_arrive_cluster(r_acc2);
        }
      }
    }
    __syncwarp();
  } else {
#include "tc_mlp2_epilogue.inc"
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // nobody leaves while the partner may still signal / read this CTA
  if (warp == 1) tmem_dealloc2<512>(tmem_base);
}

template <int LN>
static int launch_mlp2x(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& ty,
                        const CUtensorMap& tr, const CUtensorMap& tl, const Mlp2Args& a, cudaStream_t st) {
  auto kern = tc_mlp2x_kernel<LN>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM));
    configured_dev = dev;
  }
  const int tiles = ceil_div(a.M, 256);
  int groups = sm_count() / 2;
  if (groups > tiles) groups = tiles;
  V3_CUDA(launch_pdl(kern, dim3(groups * 2), dim3(M2_THREADS), (size_t)M2_SMEM, st, tx, tw1, tw2, ty, tr, tl, a));
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
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
  if (tuning(VIT3D_TUNE_MLP_PAIR) == 3 && tc_mlp3_supported(M, H, d))      // 128-column chunks, double-buffered (k_tc_mlp3.cu)
    return tc_mlp3(xn, w1, b1, w2_h, b2, residual, y, gamma, beta, eps, ln_out, ln_f32, M, H, d, st);
  const int pm = tuning(VIT3D_TUNE_MLP_PAIR);
  const int pair_mode = (M > 128 && pm != 3) ? pm : 0;                   // 1: multicast pairs (cta_group::1), 2: cta_group::2
  const bool pair = pair_mode != 0;
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
  if (pair_mode == 2) {
    if (mode == 2) return launch_mlp2x<2>(tx, tw1, tw2, ty, tr, tl, a, st);
    return mode ? launch_mlp2x<1>(tx, tw1, tw2, ty, tr, tl, a, st) : launch_mlp2x<0>(tx, tw1, tw2, ty, tr, tl, a, st);
  }
  if (pair) {
    if (mode == 2) return launch_mlp2<true, 2>(tx, tw1, tw2, ty, tr, tl, a, st);
    return mode ? launch_mlp2<true, 1>(tx, tw1, tw2, ty, tr, tl, a, st) : launch_mlp2<true, 0>(tx, tw1, tw2, ty, tr, tl, a, st);
  }
  if (mode == 2) return launch_mlp2<false, 2>(tx, tw1, tw2, ty, tr, tl, a, st);
  return mode ? launch_mlp2<false, 1>(tx, tw1, tw2, ty, tr, tl, a, st) : launch_mlp2<false, 0>(tx, tw1, tw2, ty, tr, tl, a, st);
}

}  // namespace vit3d
