// tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (both operands K-major), fp32 accumulation
// in TMEM, operands staged by TMA (128B swizzle) through a multi-stage mbarrier ring.
//
//   * persistent, warp-specialised CTA of 192 threads (1 CTA / SM):
//       warp 0   TMA producer (one lane)        warp 1   TMEM allocator + tcgen05.mma issuer (one lane)
//       warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / GELU / residual /
//                 position-embedding add -> global stores
//   * tile 128 x BLOCK_N, BLOCK_K = 128 bytes of K (64 bf16 / 32 tf32); the accumulator is double
//     buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * kind::f16 (bf16 operands) or kind::tf32 (fp32 operands read directly, no conversion pass).
//   * optional split-K (work item = tile x K-slice, fp32 atomics) for deep-K / small-output products.
#include <mutex>
#include <unordered_map>

#include "ptx.cuh"
#include "tc.cuh"

namespace vit3d {

using namespace ptx;

struct TcEpilogue {
  const float* bias = nullptr;      // [N]
  const float* residual = nullptr;  // [M,N] fp32
  const float* rowadd = nullptr;    // [(row_group+1), N] position table
  void* out = nullptr;              // [rows, N] fp32 or bf16
  void* pre = nullptr;              // pre-activation copy (type of out)
  int out_f32 = 1;
  int act = 0;
  int row_group = 0;                // >0: out row = m + m / row_group + 1
  int atomic = 0;                   // accumulate with fp32 atomics (split-K)
};

constexpr int TC_BLOCK_M = 128;
constexpr int TC_THREADS = 192;

template <int BN> struct TcCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = TC_BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epi_store_chunk(const TcEpilogue& ep, float (&v)[32], int m, int n, int N, bool full) {
  // v: 32 consecutive columns n..n+31 of row m (fp32 accumulators)
  const long long orow = ep.row_group > 0 ? (long long)m + m / ep.row_group + 1 : (long long)m;
  if (full) {
    if (ep.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (ep.rowadd) {
      const float* ra = ep.rowadd + (long long)((m % ep.row_group) + 1) * N + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ra + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (ep.atomic) {
      float* o = reinterpret_cast<float*>(ep.out) + orow * N + n;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(o + j, v[j]);
      return;
    }
    if (ep.pre) {
      if (ep.out_f32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.pre) + orow * N + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.pre) + orow * N + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
          u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
          o[j] = u;
        }
      }
    }
    if (ep.act == VIT3D_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
    }
    if (ep.residual) {
      const float4* r = reinterpret_cast<const float4*>(ep.residual + (long long)m * N + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(r + j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    if (ep.out_f32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow * N + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + orow * N + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
        u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
        o[j] = u;
      }
    }
    return;
  }
  // ragged right edge (N not a multiple of 32): element-wise
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const int nn = n + j;
    if (nn >= N) break;
    float x = v[j];
    if (ep.bias) x += ep.bias[nn];
    if (ep.rowadd) x += ep.rowadd[(long long)((m % ep.row_group) + 1) * N + nn];
    const long long oi = orow * N + nn;
    if (ep.atomic) { atomicAdd(reinterpret_cast<float*>(ep.out) + oi, x); continue; }
    if (ep.pre) {
      if (ep.out_f32) reinterpret_cast<float*>(ep.pre)[oi] = x;
      else reinterpret_cast<__nv_bfloat16*>(ep.pre)[oi] = __float2bfloat16(x);
    }
    if (ep.act == VIT3D_ACT_GELU) x = gelu_f(x);
    if (ep.residual) x += ep.residual[(long long)m * N + nn];
    if (ep.out_f32) reinterpret_cast<float*>(ep.out)[oi] = x;
    else reinterpret_cast<__nv_bfloat16*>(ep.out)[oi] = __float2bfloat16(x);
  }
}

template <bool TF32, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcEpilogue ep, int M,
               int N, int K, int tiles_m, int tiles_n, int splits) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BLOCK_K = TF32 ? 32 : 64;   // 128 bytes of K per stage row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = (nkb + splits - 1) / splits;
  const int total = tiles_m * tiles_n * splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int split = w % splits, tile = w / splits;
        const int tn = tile % tiles_n, tm = tile / tiles_n;
        const int kb0 = split * kb_per, kb1 = min(nkb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, tm * TC_BLOCK_M);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, tn * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32 ? UMMA_FMT_TF32 : UMMA_FMT_BF16, TC_BLOCK_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int split = w % splits;
        const int kb0 = split * kb_per, kb1 = min(nkb, kb0 + kb_per);
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x 32 bytes of K per stage
            const uint64_t ad = make_smem_desc(sa + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
            const uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
            umma<TF32>(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[buf]);       // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int tile = w / splits;
      const int tn = tile % tiles_n, tm = tile / tiles_n;
      const int buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const int m = tm * TC_BLOCK_M + q * 32 + lane;
      const int n0 = tn * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= N) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c, r);
        tmem_ld_wait();
        if (m < M) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epi_store_chunk(ep, v, m, n0 + c, N, n0 + c + 32 <= N);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D row-major [rows, cols] tensor, box = [box_rows, 128 bytes of columns], 128B swizzle, OOB -> 0
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, long long rows, long long cols, long long ld_elems,
                 int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return VIT3D_ERR_CUDA; }
  struct Key { const void* p; long long r, c, ld; int e, b; };
  struct Hash { size_t operator()(const Key& k) const {
    size_t h = (size_t)k.p; h = h * 1000003u ^ (size_t)k.r; h = h * 1000003u ^ (size_t)k.c;
    h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)(k.e * 1024 + k.b); return h; } };
  struct Eq { bool operator()(const Key& a, const Key& b) const {
    return a.p == b.p && a.r == b.r && a.c == b.c && a.ld == b.ld && a.e == b.e && a.b == b.b; } };
  static thread_local std::unordered_map<Key, CUtensorMap, Hash, Eq> cache;
  const Key key{ptr, rows, cols, ld_elems, elem_bytes, box_rows};
  auto itc = cache.find(key);
  if (itc != cache.end()) { *out = itc->second; return VIT3D_OK; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld_elems * elem_bytes) & 15)) {
    set_error("TMA operand must be 16-byte aligned (ptr %p, row stride %lld bytes)", ptr, ld_elems * elem_bytes);
    return VIT3D_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * elem_bytes)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VIT3D_ERR_CUDA; }
  if (cache.size() > 2048) cache.clear();
  cache.emplace(key, *out);
  return VIT3D_OK;
}

template <bool TF32, int BN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcEpilogue& ep, int M, int N, int K, int splits,
                     cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<TF32, BN>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  V3_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured_dev = dev;
  }
  const int tiles_m = ceil_div(M, TC_BLOCK_M), tiles_n = ceil_div(N, BN);
  const int total = tiles_m * tiles_n * splits;
  const int grid = total < sm_count() ? total : sm_count();
  kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, ep, M, N, K, tiles_m, tiles_n, splits);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// D[M,N] = A[M,K] B[N,K]^T with both operands dense row-major K-major; elem = 2 (bf16) or 4 (tf32)
int tc_gemm(bool tf32, const void* A, const void* B, int M, int N, int K, const TcEpilogue& ep, int splits,
            cudaStream_t st) {
  const int eb = tf32 ? 4 : 2;
  // pick the widest N tile that still gives every SM work
  const int sms = sm_count();
  const int tm = ceil_div(M, TC_BLOCK_M);
  int bn = 256;
  if (N <= 64 || (long long)tm * ceil_div(N, 256) * splits < sms) bn = 128;
  if (N <= 64 || (bn == 128 && (long long)tm * ceil_div(N, 128) * splits < sms)) bn = 64;
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, eb, M, K, K, TC_BLOCK_M);
  if (rc != VIT3D_OK) return rc;
  rc = make_tmap_2d(&tb, B, eb, N, K, K, bn);
  if (rc != VIT3D_OK) return rc;
  if (tf32) {
    if (bn == 256) return launch_tc<true, 256>(ta, tb, ep, M, N, K, splits, st);
    if (bn == 128) return launch_tc<true, 128>(ta, tb, ep, M, N, K, splits, st);
    return launch_tc<true, 64>(ta, tb, ep, M, N, K, splits, st);
  }
  if (bn == 256) return launch_tc<false, 256>(ta, tb, ep, M, N, K, splits, st);
  if (bn == 128) return launch_tc<false, 128>(ta, tb, ep, M, N, K, splits, st);
  return launch_tc<false, 64>(ta, tb, ep, M, N, K, splits, st);
}

bool tc_linear_supported(int prec, int M, int N, int K) {
  if (M <= 0) return false;
  if (prec == VIT3D_PREC_BF16) return K % 8 == 0 && N % 8 == 0 && K >= 64 && N >= 64;
  if (prec == VIT3D_PREC_TF32) return K % 4 == 0 && N % 4 == 0 && K >= 32 && N >= 64;
  return false;
}

int tc_linear_fwd(const TcLinear& t, cudaStream_t st) {
  TcEpilogue ep;
  ep.bias = t.bias; ep.residual = t.residual; ep.out = t.y; ep.pre = t.pre; ep.out_f32 = t.y_f32; ep.act = t.act;
  return tc_gemm(t.prec == VIT3D_PREC_TF32, t.x, t.w, t.M, t.N, t.K, ep, 1, st);
}

// placeholders until the dedicated kernels land
bool tc_patch_embed_supported(int, int, int, int, int, int, int, int) { return false; }
int tc_patch_embed_fwd(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int,
                       int, int, cudaStream_t) {
  return VIT3D_ERR_UNSUPPORTED;
}

}  // namespace vit3d
