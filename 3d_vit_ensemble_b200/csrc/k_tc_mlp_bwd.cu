// Fused MLP backward, data-gradient chain (a8: the autograd backward of Mlp.forward, modeling.py:118-124):
//
//     da  = gy W2                       gy [M,256] = d loss / d (fc2 output) (bf16, dropout of :123 already undone)
//     dh  = da * dact                   dact [M,d] = gelu'(pre) * keep / (1-p), written by the training fc1 epilogue
//                                       (vit3d_fc1_train_fwd): GELU derivative and the Dropout of :121 in one factor
//     dxn = dh W1                       [M,256] fp32: gradient w.r.t. the ffn_norm output
//     db1 += column sums of dh          fc1 bias gradient
//
// in ONE kernel: the [M,d] tensor `da` never exists, `dh` is written once (the fc1 / fc2 weight-gradient GEMMs
// read it) and never re-read by this chain.  Unfused, the chain was three passes - a GEMM writing da, an
// element-wise pass reading da + pre and writing dh, a GEMM reading dh - 5 x M x d x 2 bytes of HBM traffic per
// layer against 2 x here, and the element-wise pass alone took longer than both GEMMs together.
//
// Per 128-row tile (CTA) the d columns are processed in chunks of 128, the structure of k_tc_mlp2.cu:
//   MMA1(c): D1[128 x 128] = gy[128 x 256] * W2t[c*128 .., :]^T      16 tcgen05.mma (N = 128), TMEM cols b*128..
//   EPI(c) : 16 warps: D1 from TMEM (one row per thread, 32 columns) times `dact` from shared memory - the chunk of
//            dact was TMA-loaded into the SAME buffer the product goes to, in the K-major 128B-swizzled layout tcgen05
//            reads - dh written back IN PLACE as bf16; column sums reduced across the warp by a shuffle butterfly ->
//            32 atomics per warp (fc1 bias gradient)
//   MMA2(c): dxn[128 x 256] += dh[128 x 128] * W1t[:, c*128 ..]^T     8 tcgen05.mma (N = 256), TMEM cols 256..511
//   STORE(c): the same buffer leaves for HBM by two bulk tensor stores (dh for the weight gradients), issued by a
//            dedicated warp
// Issue order MMA1(c+1), MMA2(c), MMA1(c+2), ...; D1 and the dact/dh buffer are double buffered, so the epilogue of a
// chunk overlaps the MMAs of its neighbours.  Weights (W2t chunk 64 KB + W1t chunk 64 KB per chunk) stream from L2
// through a 96 KB TMA ring.  Measured history (conf 18, batch 256; the unfused chain takes 139 us):
//    85 us  this structure with `pre` in the buffer and GELU' x keep-mask evaluated here (~700 instructions per thread and
//           chunk: the epilogue, not the tensor pipe, was the bound - ncu: 31 % tensor-pipe activity, 39 % of the stall
//           samples on the keep-bit load)
//   101 us  `pre` fetched by the epilogue threads into registers one chunk ahead (row-owner 256-bit loads: every warp
//           instruction touches 32 rows 6 KB apart)
//    79 us  the factor gelu'(pre) x keep/(1-p) written by the forward (dact) and fetched to registers two chunks ahead
//   105 us  the same with 256-column chunks, single-buffered like the forward kernel
//    now    dact TMA-loaded into the dh buffer (this file): the loads are whole 128-byte rows, the epilogue is ~250
//           instructions per thread and chunk
//
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2..17: epilogue   warp 18: dh store
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "tc.cuh"
#include "tc_epilogue.cuh"

namespace vit3d {

using namespace ptx;

constexpr int MB_H = 256;                       // hidden size: K of MMA1, N of MMA2
constexpr int MB_NC = 128;                      // d columns per chunk
constexpr int MB_G_BYTES = 128 * MB_H * 2;      // 65536: gy tile, 4 k-blocks of [128 x 64]
constexpr int MB_D_BYTES = 128 * MB_NC * 2;     // 32768: pre / dh chunk, 2 k-blocks of [128 x 64]
constexpr int MB_NBUF = 2;
constexpr int MB_SLOT = 32 * 1024;
constexpr int MB_NST = 3;
constexpr int MB_THREADS = 64 + 512 + 32;
constexpr int MB_SMEM = MB_G_BYTES + MB_NBUF * MB_D_BYTES + MB_NST * MB_SLOT + 1024;
static_assert(MB_SMEM <= 232448, "over the 227 KB shared-memory limit");

struct MlpBwdArgs {
  float* db1 = nullptr;             // [d] (null: skip)
  int M = 0, d = 0;
};

__global__ void __launch_bounds__(MB_THREADS, 1)
tc_mlp_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmW2t,
                  const __grid_constant__ CUtensorMap tmW1t, const __grid_constant__ CUtensorMap tmDact,
                  const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDx, MlpBwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* s_g = smem;
  uint8_t* s_d = smem + MB_G_BYTES;
  uint8_t* s_w = s_d + MB_NBUF * MB_D_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + MB_NST * MB_SLOT);
  uint64_t* g_full = bars;                 // gy tile landed
  uint64_t* g_empty = bars + 1;            // (commit) the last MMA1 of the tile has read it
  uint64_t* w_full = bars + 2;             // [MB_NST]
  uint64_t* w_empty = bars + 6;            // [MB_NST] (commit)
  uint64_t* dact_full = bars + 10;          // [2] dact chunk landed in buffer b
  uint64_t* acc1_full = bars + 12;         // [2] (commit)
  uint64_t* acc1_empty = bars + 14;        // [2] 16 arrivals
  uint64_t* dh_full = bars + 16;           // [2] 16 arrivals: dh written
  uint64_t* dh_empty = bars + 18;          // [2] 2 arrivals: MMA2 commit + store warp
  uint64_t* acc2_full = bars + 20;         // (commit)
  uint64_t* acc2_empty = bars + 21;        // 16 arrivals: final epilogue done (accumulator drained, staging free)
  uint64_t* stores_done = bars + 22;       // 1 arrival per tile: every dh store of the tile has read its buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int M = args.M, d = args.d;
  const int nch = d / MB_NC;
  const int tiles = (M + 127) / 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmG);
    prefetch_tmap(&tmW2t);
    prefetch_tmap(&tmW1t);
    prefetch_tmap(&tmDact);
    mbar_init(g_full, 1);
    mbar_init(g_empty, 1);
    for (int s = 0; s < MB_NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&dact_full[b], 1);
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], 16);
      mbar_init(&dh_full[b], 16);
      mbar_init(&dh_empty[b], 2);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    mbar_init(stores_done, 1);
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
        mbar_arrive_expect_tx(&w_full[stage], MB_SLOT);
        return s_w + stage * MB_SLOT;
      };
      auto advance = [&]() { if (++stage == MB_NST) { stage = 0; wphase ^= 1; } };
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        mbar_wait(g_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(g_full, MB_G_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(s_g + kb * 16384, &tmG, g_full, kb * 64, t * 128);
        for (int s = 0; s <= nch; ++s) {
          if (s < nch) {
            for (int j = 0; j < 2; ++j) {                    // W2t rows s*128.., k-blocks 2j, 2j+1
              uint8_t* dst = slot();
              tma_load_2d(dst, &tmW2t, &w_full[stage], (2 * j) * 64, s * MB_NC);
              tma_load_2d(dst + 16384, &tmW2t, &w_full[stage], (2 * j + 1) * 64, s * MB_NC);
              advance();
            }
          }
          if (s >= 1) {
            for (int j = 0; j < 2; ++j) {                    // W1t [256 rows, cols (s-1)*128 + j*64 ..]
              uint8_t* dst = slot();
              tma_load_2d(dst, &tmW1t, &w_full[stage], (s - 1) * MB_NC + j * 64, 0);
              advance();
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected thread walks the schedule)
    constexpr uint32_t idesc1 = make_idesc(UMMA_FMT_BF16, 128, MB_NC, 0, 0);
    constexpr uint32_t idesc2 = make_idesc(UMMA_FMT_BF16, 128, MB_H, 0, 0);
    const uint64_t gdesc0 = make_smem_desc(smem_u32(s_g), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t ddesc0 = make_smem_desc(smem_u32(s_d), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t wdesc0 = make_smem_desc(smem_u32(s_w), 16, 1024, UMMA_LAYOUT_SW128);
    if (elect_one()) {
      int stage = 0;
      uint32_t wphase = 0;
      uint32_t g1 = 0, g2 = 0;
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        mbar_wait(g_full, it & 1);
        for (int s = 0; s <= nch; ++s) {
          if (s < nch) {
            // ---- MMA1(s): D1[b] = gy * W2t[s]^T
            const uint32_t b = g1 & 1;
            mbar_wait(&acc1_empty[b], ((g1 >> 1) & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (MB_SLOT >> 4));
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t ad = gdesc0 + (uint64_t)((2 * j + kk) * 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma<false>(tmem_base + b * MB_NC, ad + 2 * k, bd + kk * 1024 + 2 * k, idesc1, (j | kk | k) ? 1u : 0u);
              }
              umma_commit(&w_empty[stage]);
              if (++stage == MB_NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(&acc1_full[b]);
            if (s == nch - 1) umma_commit(g_empty);
            ++g1;
          }
          if (s >= 1) {
            // ---- MMA2(s-1): dxn += dh[b] * W1t[:, s-1]^T
            const int c = s - 1;
            const uint32_t b = g2 & 1;
            if (c == 0) mbar_wait(acc2_empty, (it & 1) ^ 1);
            mbar_wait(&dh_full[b], (g2 >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              mbar_wait(&w_full[stage], wphase);
              const uint64_t ad = ddesc0 + (uint64_t)(b * (MB_D_BYTES >> 4) + j * 1024);
              const uint64_t bd = wdesc0 + (uint64_t)(stage * (MB_SLOT >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma<false>(tmem_base + 256, ad + 2 * k, bd + 2 * k, idesc2, (c | j | k) ? 1u : 0u);
              umma_commit(&w_empty[stage]);
              if (++stage == MB_NST) { stage = 0; wphase ^= 1; }
            }
            umma_commit(&dh_empty[b]);
            if (c == nch - 1) umma_commit(acc2_full);
            ++g2;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 18) {
    // ===================================================== dh store / dact load warp: owns the life cycle of the two chunk
    // buffers (dact lands -> epilogue multiplies in place -> MMA2 + store read it -> next dact lands).  Kept out of the
    // weight producer on purpose: there a wait for a free buffer held up the weight stream behind it (measured 89 us).
    if (lane == 0) {
      const int total = ((tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nch;   // chunks of this CTA
      auto load_dact = [&](int gi) {            // chunk index gi of this CTA -> (tile, chunk)
        const int t = (int)blockIdx.x + (gi / nch) * (int)gridDim.x, c = gi % nch, b = gi & 1;
        mbar_arrive_expect_tx(&dact_full[b], MB_D_BYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(s_d + b * MB_D_BYTES + kb * 16384, &tmDact, &dact_full[b], c * MB_NC + kb * 64, t * 128);
      };
      if (total > 0) load_dact(0);
      if (total > 1) load_dact(1);
      int gi = 0;
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        for (int c = 0; c < nch; ++c, ++gi) {
          const int b = gi & 1;
          const uint32_t ph = (gi >> 1) & 1;
          mbar_wait(&dh_full[b], ph);
          for (int kb = 0; kb < 2; ++kb)
            tma_store_2d(&tmDh, smem_u32(s_d + b * MB_D_BYTES + kb * 16384), c * MB_NC + kb * 64, t * 128);
          bulk_store_commit();
          bulk_store_wait_read();
          mbar_arrive(&dh_empty[b]);
          if (gi + 2 < total && (gi + 2) / nch == gi / nch) {
            mbar_wait(&dh_empty[b], ph);                    // ... and MMA2 of this chunk has read the buffer too
            load_dact(gi + 2);
          }
        }
        mbar_arrive(stores_done);
        // the first two chunks of the NEXT tile land in buffers the final epilogue of this tile stages its panels in
        const int g0 = (it + 1) * nch;
        if (g0 < total) {
          mbar_wait(acc2_empty, it & 1);
          for (int k = 0; k < 2 && g0 + k < total; ++k) {
            const int gk = g0 + k;
            if (gk >= 2) mbar_wait(&dh_empty[gk & 1], ((gk >> 1) & 1) ^ 1);
            load_dact(gk);
          }
        }
      }
      bulk_store_wait_all();
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                     // TMEM lane quarter
    const int part = ew >> 2;                   // 32-column slice of the 128-column chunk
    const int row = q * 32 + lane;
    const int sw7 = row & 7;
    // this thread's 64 bytes of a chunk row: k-block part >> 1, 16-byte chunks (part & 1) * 4 .. + 3 (swizzled)
    const uint32_t d_row = smem_u32(s_d) + (part >> 1) * 16384 + row * 128;
    const int j0 = (part & 1) * 4;
    uint8_t* stage_ptr = s_d + ew * 4096;       // final-epilogue panel buffer (the pre / dh buffers are idle then)
    const uint32_t stage_s = smem_u32(stage_ptr);
    uint32_t g = 0;
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m = t * 128 + row;
      for (int c = 0; c < nch; ++c, ++g) {
        const int b = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        mbar_wait(&acc1_full[b], ph);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + b * MB_NC + part * 32, r);
        mbar_wait(&dact_full[b], ph);
        uint4 pv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pv[j] = ld_shared_v4(d_row + b * MB_D_BYTES + (((j0 + j) ^ sw7) << 4));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);
        float gv[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t pw[4] = {pv[j].x, pv[j].y, pv[j].z, pv[j].w};
          uint32_t ow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = 8 * j + 2 * i;
            const float v0 = __uint_as_float(r[e]) * __uint_as_float(pw[i] << 16);
            const float v1 = __uint_as_float(r[e + 1]) * __uint_as_float(pw[i] & 0xffff0000u);
            gv[e] = v0;
            gv[e + 1] = v1;
            ow[i] = pack2_bf16(v0, v1);
          }
          st_shared_v4(d_row + b * MB_D_BYTES + (((j0 + j) ^ sw7) << 4), ow[0], ow[1], ow[2], ow[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dh_full[b]);
        if (args.db1) {
          // column sums over the warp's 32 rows: butterfly reduce-scatter, lane j ends with column j
#pragma unroll
          for (int h = 16; h >= 1; h >>= 1) {
            const bool up = (lane & h) != 0;
#pragma unroll
            for (int i = 0; i < h; ++i) {
              const float send = up ? gv[i] : gv[i + h];
              const float recv = __shfl_xor_sync(0xffffffffu, send, h);
              gv[i] = (up ? gv[i + h] : gv[i]) + recv;
            }
          }
          // after the rounds lane l holds column bitrev-free index: round h keeps columns whose bit h equals the lane's
          atomicAdd(args.db1 + c * MB_NC + part * 32 + lane, gv[0]);
        }
      }
      // ---------------- final epilogue of the tile: dxn rows -> two [32 x 32] fp32 panels per warp
      mbar_wait(acc2_full, it & 1);
      mbar_wait(stores_done, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + part * 64;
      const uint32_t my_row = stage_s + lane * 128;
      const int sl7 = lane & 7;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + p * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(my_row + ((j ^ sl7) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmDx, stage_s, part * 64 + p * 32, t * 128 + q * 32);
          bulk_store_commit();
          bulk_store_wait_read();
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
    if (lane == 0) bulk_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool tc_mlp_bwd_supported(int M, int H, int d) { return M > 0 && H == MB_H && d % MB_NC == 0 && d >= MB_NC; }

int tc_mlp_bwd(const void* gy, const void* w2_t, const void* w1_t, const void* dact, void* dh, float* dxn, float* db1, int M,
               int H, int d, cudaStream_t st) {
  if (!tc_mlp_bwd_supported(M, H, d)) V3_UNSUPPORTED("fused MLP backward: unsupported shape M=%d H=%d d=%d", M, H, d);
  CUtensorMap tg, tw2, tw1, tp, td, tx;
  int rc = make_tmap_2d(&tg, gy, 2, M, H, H, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw2, w2_t, 2, d, H, H, 128, 64, 128);        // W2^T [d, 256]: rows = chunk columns, K = 256
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tw1, w1_t, 2, H, d, d, 256, 64, 128);        // W1^T [256, d]: N = 256 rows, K = chunk columns
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tp, dact, 2, M, d, d, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&td, dh, 2, M, d, d, 128, 64, 128);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tx, dxn, 4, M, H, H, 32, 32, 128);
  if (rc != VIT3D_OK) return rc;
  MlpBwdArgs a;
  a.db1 = db1; a.M = M; a.d = d;
  auto kern = tc_mlp_bwd_kernel;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    configured_dev = dev;
  }
  const int tiles = ceil_div(M, 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  V3_CUDA(launch_pdl(kern, dim3(grid), dim3(MB_THREADS), (size_t)MB_SMEM, st, tg, tw2, tw1, tp, td, tx, a));
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

}  // namespace vit3d
