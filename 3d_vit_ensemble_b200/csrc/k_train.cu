// Support kernels of the fused BF16 training step (hidden size 256; a8: the autograd backward of
// modeling.py:118-124, :187-197 as explicit kernels).  The GEMMs live in k_tc_gemm*.cu; this file holds the
// bandwidth-bound passes between them, each of which reads its operands ONCE and produces everything the
// next GEMMs need:
//
//   dropout_bits      all keep-masks of a training step (embedding, 2 per Block) as bit arrays, one launch;
//                     GEMM epilogues and the kernels below apply them (same Philox decision as vit3d_dropout)
//   ln256_fwd         LayerNorm 256 (+ dropout of its input): residual stream fp32, bf16 GEMM operand, statistics
//   ln256_bwd         LayerNorm backward + residual-gradient add + bf16 copy (optionally dropout-masked) of the
//                     result for the dgrad / wgrad GEMMs that follow + its column sums (the bias gradient of the
//                     Linear below) + gamma / beta gradients
//   mul_colsum_bwd    dh = da * dact (dact = gelu'(pre) * keep / (1 - p) from the training fc1 epilogue), with the fc1
//                     bias gradient (column sums) folded in - the unfused form of k_tc_mlp_bwd.cu's middle stage
//   head_bwd          classification-head backward: d(encoded) rows, head weight / bias gradients
//   refresh_shadows   every low-precision weight shadow of a model (bf16, transposed bf16, packed q|k|v, TF32
//                     patch filter) from the fp32 masters in one launch, driven by a device job table
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vit3d {

constexpr int T_H = 256;

// ============================================================================ dropout keep bits
struct DropSegs {
  int n;
  unsigned site[VIT3D_MAX_DROP_SEGS];
  long long word0[VIT3D_MAX_DROP_SEGS + 1];     // first 32-element word of each segment; [n] = total
};

__global__ void __launch_bounds__(256) dropout_bits_kernel(uint32_t* __restrict__ bits, DropSegs segs, uint32_t th,
                                                           unsigned long long seed, unsigned step,
                                                           const unsigned* __restrict__ step_dev) {
  if (step_dev) step += *step_dev;
  const long long total = segs.word0[segs.n];
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < segs.n && w >= segs.word0[s + 1]) ++s;
    const unsigned long long blk0 = (unsigned long long)(w - segs.word0[s]) * 4ull;   // Philox block = 8 elements
    const unsigned site = segs.site[s];
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long b = blk0 + j;
      const Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), site, step, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        word |= ((r.v[k] & 0xffffu) >= th ? 1u : 0u) << (8 * j + 2 * k);
        word |= ((r.v[k] >> 16) >= th ? 1u : 0u) << (8 * j + 2 * k + 1);
      }
    }
    bits[w] = word;
  }
}

int launch_dropout_bits(uint32_t* bits, int nseg, const unsigned* sites, const long long* nelems, float p,
                        unsigned long long seed, unsigned step, const unsigned* step_dev, cudaStream_t st) {
  DropSegs segs;
  segs.n = nseg;
  long long w = 0;
  for (int i = 0; i < nseg; ++i) {
    segs.site[i] = sites[i];
    segs.word0[i] = w;
    w += nelems[i] / 32;
  }
  segs.word0[nseg] = w;
  if (w <= 0) return VIT3D_OK;
  long long blocks = (w + 255) / 256;
  // two CTAs (512 threads) per SM: the masks are drawn on a side stream next to the forward's GEMMs, whose one CTA
  // per SM (576-608 threads, ~220 KB of shared memory) must still find its thread slots - eight CTAs per SM would
  // fill all 2048 and push the GEMMs behind this ALU-bound kernel
  // eight CTAs per SM.  (Measured: fewer CTAs so that the forward's GEMM CTAs co-reside, a matching shared-memory
  // carve-out, or drawing the next step's masks beside the backward's GEMMs all leave the step time unchanged or
  // worse - the Philox work, ~0.2 ms per conf-18 step, is not hidden by a side stream.)
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dropout_bits_kernel<<<(int)blocks, 256, 0, st>>>(bits, segs, dropout_thresh(p), seed, step, step_dev);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ LayerNorm 256, forward
// One warp per row, a lane owns 8 consecutive columns (two 16-byte loads, one 16-byte bf16 store).
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 o = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&o);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) ln256_fwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ bits,
                                                        float sc, float* __restrict__ xd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ y, float* __restrict__ yf,
                                                        float* __restrict__ mean, float* __restrict__ rstd, int M,
                                                        float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[8];
  load8(x + (size_t)row * T_H + lane * 8, v);
  if (bits) {
    const uint32_t m = bits[(size_t)row * (T_H / 8) + lane];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((m >> j) & 1u) ? v[j] * sc : 0.f;
    if (xd) store8(xd + (size_t)row * T_H + lane * 8, v);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mu = warp_sum(s) * (1.0f / T_H);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = v[j] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / T_H) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  float g[8], b[8], o[8];
  load8(gamma + lane * 8, g);
  load8(beta + lane * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf((v[j] - mu) * rs, g[j], b[j]);
  if (y) store8_bf16(y + (size_t)row * T_H + lane * 8, o);
  if (yf) store8(yf + (size_t)row * T_H + lane * 8, o);
}

int launch_ln256_fwd(const float* x, const uint8_t* bits, float sc, float* xd, const float* gamma, const float* beta,
                     void* y_bf16, float* y_f32, float* mean, float* rstd, int M, float eps, cudaStream_t st) {
  if (M <= 0) return VIT3D_OK;
  ln256_fwd_kernel<<<ceil_div(M, 8), 256, 0, st>>>(x, bits, sc, xd, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16),
                                                   y_f32, mean, rstd, M, eps);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ LayerNorm 256, backward (fused)
//   o   = dres + rstd * (g*dy - mean_h(g*dy) - xhat * mean_h(g*dy*xhat))          gradient wrt the LN input
//   dx  = o            (fp32; multiplied by the keep mask when mask_f32: the embedding dropout below layer 0)
//   dxb = bf16(o * keep * sc)   operand of the dgrad / wgrad GEMMs of the Linear whose output o is the gradient of
//   dbias  += column sums of (o * keep * sc)      (that Linear's bias gradient)
//   dgamma += sum dy * xhat;  dbeta += sum dy
// A warp walks over rows (stride = all warps of the grid) and keeps its column partials in registers; a block
// reduces them in shared memory and issues one atomic per column.
__global__ void __launch_bounds__(256) ln256_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ dres,
                                                        const uint8_t* __restrict__ bits, float sc, int mask_f32,
                                                        float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                        float* __restrict__ dbias, int M) {
  __shared__ float red[3][8][T_H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * 8;
  float g[8], dg[8], db[8], cs[8];
  load8(gamma + c0, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = db[j] = cs[j] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    float d[8], xv[8], o[8];
    load8(dy + (size_t)row * T_H + c0, d);
    load8(x + (size_t)row * T_H + c0, xv);
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xv[j] = (xv[j] - mu) * rs;                // xhat
      const float gd = g[j] * d[j];
      s1 += gd;
      s2 = fmaf(gd, xv[j], s2);
      dg[j] = fmaf(d[j], xv[j], dg[j]);
      db[j] += d[j];
    }
    s1 = warp_sum(s1) * (1.0f / T_H);
    s2 = warp_sum(s2) * (1.0f / T_H);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rs * (g[j] * d[j] - s1 - xv[j] * s2);
    if (dres) {
      float r[8];
      load8(dres + (size_t)row * T_H + c0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += r[j];
    }
    float ob[8];
    if (bits) {
      const uint32_t m = bits[(size_t)row * (T_H / 8) + lane];
#pragma unroll
      for (int j = 0; j < 8; ++j) ob[j] = ((m >> j) & 1u) ? o[j] * sc : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) ob[j] = o[j];
    }
    if (dx) {
      if (mask_f32) store8(dx + (size_t)row * T_H + c0, ob);
      else store8(dx + (size_t)row * T_H + c0, o);
    }
    if (dxb) store8_bf16(dxb + (size_t)row * T_H + c0, ob);
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] += ob[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][warp][c0 + j] = dg[j];
    red[1][warp][c0 + j] = db[j];
    red[2][warp][c0 + j] = cs[j];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { a0 += red[0][w][c]; a1 += red[1][w][c]; a2 += red[2][w][c]; }
  if (dgamma) atomicAdd(dgamma + c, a0);
  if (dbeta) atomicAdd(dbeta + c, a1);
  if (dbias) atomicAdd(dbias + c, a2);
}

int launch_ln256_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     const float* dres, const uint8_t* bits, float sc, int mask_f32, float* dx, void* dxb, float* dgamma,
                     float* dbeta, float* dbias, int M, cudaStream_t st) {
  if (M <= 0) return VIT3D_OK;
  int blocks = 2 * sm_count();
  if (blocks > ceil_div(M, 8)) blocks = ceil_div(M, 8);
  ln256_bwd_kernel<<<blocks, 256, 0, st>>>(dy, x, mean, rstd, gamma, dres, bits, sc, mask_f32, dx,
                                           reinterpret_cast<__nv_bfloat16*>(dxb), dgamma, dbeta, dbias, M);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ dh = da * dact (+ fc1 bias gradient)
// The unfused form of the MLP backward's element-wise stage (k_tc_mlp_bwd.cu does it inside the fused kernel; this
// serves mlp widths that kernel does not): dh[m, c] = da[m, c] * dact[m, c], db[c] += sum_m dh[m, c].  dact = gelu'(pre) *
// keep / (1 - p) as written by the training fc1 epilogue.  bf16 in / out, 8 per thread.
__global__ void __launch_bounds__(256) mul_colsum_bwd_kernel(const uint4* __restrict__ da, const uint4* __restrict__ dact,
                                                             uint4* __restrict__ dh, float* __restrict__ db, int M, int d,
                                                             int rows_per_block) {
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (col < d) {
    for (int m = r0 + warp; m < r1; m += 8) {
      const size_t idx = ((size_t)m * d + col) >> 3;
      const uint4 a = da[idx], h = dact[idx];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, hw[4] = {h.x, h.y, h.z, h.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = __uint_as_float(aw[j] << 16) * __uint_as_float(hw[j] << 16);
        const float g1 = __uint_as_float(aw[j] & 0xffff0000u) * __uint_as_float(hw[j] & 0xffff0000u);
        cs[2 * j] += g0;
        cs[2 * j + 1] += g1;
        const __nv_bfloat162 o = __floats2bfloat162_rn(g0, g1);
        ow[j] = *reinterpret_cast<const uint32_t*>(&o);
      }
      dh[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  if (!db) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane][j] = cs[j];
  __syncthreads();
  if (warp == 0 && col < d) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][lane][j];
      atomicAdd(db + col + j, s);
    }
  }
}

int launch_mul_colsum_bwd(const void* da, const void* dact, void* dh, float* db, int M, int d, cudaStream_t st) {
  if (M <= 0 || d <= 0) return VIT3D_OK;
  const int gx = ceil_div(d, 256);
  int gy = (8 * sm_count()) / gx;
  if (gy < 1) gy = 1;
  int rpb = ceil_div(M, gy);
  if (rpb < 32) rpb = 32;
  gy = ceil_div(M, rpb);
  mul_colsum_bwd_kernel<<<dim3(gx, gy), 256, 0, st>>>(reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(dact),
                                                      reinterpret_cast<uint4*>(dh), db, M, d, rpb);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ classification head, backward
// logits[b] = enc[b*S, :] . w + bias (modeling.py:281, num_classes = 1).  d(enc) is zero outside the cls rows.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dlogit, const float* __restrict__ enc,
                                                       const float* __restrict__ w, float* __restrict__ denc,
                                                       float* __restrict__ dw, float* __restrict__ db, int B, int S,
                                                       int write_blocks) {
  if ((int)blockIdx.x < write_blocks) {
    // d(enc): one float4 per thread
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // float4 index
    const long long total = (long long)B * S * (T_H / 4);
    if (i >= total) return;
    const int row = (int)(i / (T_H / 4)), c4 = (int)(i % (T_H / 4));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row % S == 0) {
      const float g = dlogit[row / S];
      const float4 wv = reinterpret_cast<const float4*>(w)[c4];
      o = make_float4(g * wv.x, g * wv.y, g * wv.z, g * wv.w);
    }
    reinterpret_cast<float4*>(denc)[i] = o;
    return;
  }
  // head weight / bias gradient: 64 volumes per block, thread = column
  const int b0 = ((int)blockIdx.x - write_blocks) * 64, b1 = min(B, b0 + 64);
  const int c = threadIdx.x;
  float acc = 0.f, sb = 0.f;
  for (int b = b0; b < b1; ++b) {
    const float g = dlogit[b];
    acc = fmaf(g, enc[(size_t)b * S * T_H + c], acc);
    sb += g;
  }
  atomicAdd(dw + c, acc);
  if (c == 0) atomicAdd(db, sb);
}

int launch_head_bwd(const float* dlogit, const float* enc, const float* w, float* denc, float* dw, float* db, int B, int S,
                    cudaStream_t st) {
  if (B <= 0) return VIT3D_OK;
  const int wb = ceil_div((long long)B * S * (T_H / 4), 256);
  head_bwd_kernel<<<wb + ceil_div(B, 64), 256, 0, st>>>(dlogit, enc, w, denc, dw, db, B, S, wb);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ weight-gradient partial tiles -> gradients
// One launch per training step: for every weight-gradient GEMM of the step, sum the split-K slices of each output tile
// (vit3d_wgrad_partial stored them densely, slice-minor) and ADD the sum to the parameter gradient(s).
struct WgradJob {
  const float* ws;       // [tiles][splits][128][bn]
  float* dst[3];         // rows [i*seg_rows, (i+1)*seg_rows) of the product go to dst[i] (seg_rows = 0: all to dst[0])
  int seg_rows, rows, cols, bn, splits, tiles_n;
  int block0;            // first block of this job in the launch
  int pad;
};
static_assert(sizeof(WgradJob) == VIT3D_WGRAD_JOB_BYTES, "WgradJob layout is part of the C ABI");
struct WgradJobs {
  int n;
  int pad;
  WgradJob j[VIT3D_MAX_WGRAD_JOBS];
};

constexpr int WR_F4_PER_BLOCK = 1024;     // float4 per block: 256 threads x 4

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradJobs jobs) {
  int k = 0;
  while (k + 1 < jobs.n && (int)blockIdx.x >= jobs.j[k + 1].block0) ++k;
  const WgradJob& jb = jobs.j[k];
  const int tile_f4 = 128 * jb.bn / 4;
  const long long total = (long long)(jb.rows / 128) * jb.tiles_n * tile_f4;       // float4 of the whole output
  const long long base = (long long)((int)blockIdx.x - jb.block0) * WR_F4_PER_BLOCK;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = base + u * 256 + threadIdx.x;
    if (e >= total) break;
    const int tile = (int)(e / tile_f4), r = (int)(e % tile_f4);
    const float4* src = reinterpret_cast<const float4*>(jb.ws) + ((long long)tile * jb.splits) * tile_f4 + r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < jb.splits; ++s) {
      const float4 v = __ldg(src + (long long)s * tile_f4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int tm = tile / jb.tiles_n, tn = tile % jb.tiles_n;
    const int row = tm * 128 + (r * 4) / jb.bn, col = tn * jb.bn + (r * 4) % jb.bn;
    float* d = jb.dst[0];
    int lrow = row;
    if (jb.seg_rows > 0) {
      const int seg = row / jb.seg_rows;
      d = jb.dst[seg];
      lrow = row - seg * jb.seg_rows;
    }
    float4* o = reinterpret_cast<float4*>(d + (long long)lrow * jb.cols + col);
    float4 cur = *o;
    cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
    *o = cur;
  }
}

int launch_wgrad_reduce(const void* host_jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return VIT3D_OK;
  WgradJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  jobs.n = njobs;
  memcpy(jobs.j, host_jobs, sizeof(WgradJob) * njobs);
  int blocks = 0;
  for (int i = 0; i < njobs; ++i) {
    WgradJob& j = jobs.j[i];
    j.block0 = blocks;
    const long long total = (long long)(j.rows / 128) * j.tiles_n * (128 * j.bn / 4);
    blocks += (int)((total + WR_F4_PER_BLOCK - 1) / WR_F4_PER_BLOCK);
  }
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(jobs);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

// ============================================================================ weight shadows
// kinds: 0 bf16 copy, 1 bf16 transposed copy, 2 fp16 copy, 3 fp32 copy rounded to TF32, 4 fp32 copy
struct ShadowJob {
  const float* src;     // [rows, cols] fp32, dense
  void* dst;            // element (r, c) at dst[r * ld + c] (kind 1: dst[c * ld + r])
  int rows, cols, ld, kind;
  int tile0;            // first 64 x 64 tile (VIT3D_SHADOW_TILE) of this job in the launch
  int tiles_c;          // tiles per row of tiles
};
static_assert(sizeof(ShadowJob) == VIT3D_SHADOW_JOB_BYTES, "ShadowJob layout is part of the C ABI");

// One CTA = one 64 x 64 tile (VIT3D_SHADOW_TILE): four 16-byte loads in flight per thread, 128-byte output rows
// (16 bytes per thread).  (32 x 32 tiles with scalar loads and 64-byte output rows ran at 1.3 TB/s: 83 us for the 118 MB
// of the conf-18 model; several of those tiles per CTA was slower still - the bound was bytes in flight, not CTA turnover.)
constexpr int SH_T = VIT3D_SHADOW_TILE;
__global__ void __launch_bounds__(256) refresh_shadows_kernel(const ShadowJob* __restrict__ jobs, int njobs,
                                                              unsigned* __restrict__ step_dev) {
  __shared__ float tile[SH_T][SH_T + 1];
  if (blockIdx.x == 0 && threadIdx.x == 0 && step_dev) *step_dev += 1u;
  const int t = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {                     // last job with tile0 <= t
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile0 <= t) lo = mid; else hi = mid - 1;
  }
  const ShadowJob jb = jobs[lo];
  const int local = t - jb.tile0;
  const int r0 = (local / jb.tiles_c) * SH_T, c0 = (local % jb.tiles_c) * SH_T;
  const bool vec_in = (jb.cols & 3) == 0 && (reinterpret_cast<uintptr_t>(jb.src) & 15) == 0;
  {
    const int c4 = (threadIdx.x & 15) * 4, rr = threadIdx.x >> 4;       // 16 threads per row, 16 rows per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + rr + 16 * i, c = c0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < jb.rows) {
        if (vec_in && c + 3 < jb.cols) {
          v = *reinterpret_cast<const float4*>(jb.src + (size_t)r * jb.cols + c);
        } else {
          const float* p = jb.src + (size_t)r * jb.cols;
          if (c < jb.cols) v.x = p[c];
          if (c + 1 < jb.cols) v.y = p[c + 1];
          if (c + 2 < jb.cols) v.z = p[c + 2];
          if (c + 3 < jb.cols) v.w = p[c + 3];
        }
      }
      float* d = &tile[rr + 16 * i][c4];
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  }
  __syncthreads();
  const bool half_out = jb.kind == 0 || jb.kind == 1 || jb.kind == 2;
  if (half_out) {
    // 8 consecutive output elements per thread: rows of the tile (kinds 0, 2) or its columns (kind 1: transposed)
    const bool tr = jb.kind == 1;
    const int orows = tr ? jb.cols : jb.rows, ocols = tr ? jb.rows : jb.cols;      // extent of the output matrix
    const int or0 = tr ? c0 : r0, oc0 = tr ? r0 : c0;
    const bool vec_out = (jb.ld & 7) == 0 && (reinterpret_cast<uintptr_t>(jb.dst) & 15) == 0;
    const int ch = (threadIdx.x & 7) * 8, rr = threadIdx.x >> 3;        // 8 threads per output row, 32 rows per pass
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int lr = rr + 32 * i;
      const int orow = or0 + lr, ocol = oc0 + ch;
      if (orow >= orows || ocol >= ocols) continue;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = tr ? tile[ch + k][lr] : tile[lr][ch + k];
      uint16_t* o = reinterpret_cast<uint16_t*>(jb.dst) + (size_t)orow * jb.ld + ocol;
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (jb.kind == 2) {
          const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
          w[k] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
          w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      if (vec_out && ocol + 7 < ocols) {
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (ocol + k < ocols) o[k] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
      }
    }
    return;
  }
  // fp32 outputs (kind 3: rounded to tf32, kind 4: copy): 4 consecutive elements per thread
  {
    const bool vec_out = (jb.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(jb.dst) & 15) == 0;
    const int c4 = (threadIdx.x & 15) * 4, rr = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lr = rr + 16 * i;
      const int r = r0 + lr, c = c0 + c4;
      if (r >= jb.rows || c >= jb.cols) continue;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = jb.kind == 3 ? round_tf32(tile[lr][c4 + k]) : tile[lr][c4 + k];
      float* o = reinterpret_cast<float*>(jb.dst) + (size_t)r * jb.ld + c;
      if (vec_out && c + 3 < jb.cols) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < jb.cols) o[k] = v[k];
      }
    }
  }
}

int launch_refresh_shadows(const void* jobs, int njobs, int total_tiles, unsigned* step_dev, cudaStream_t st) {
  if (njobs <= 0 || total_tiles <= 0) return VIT3D_OK;
  refresh_shadows_kernel<<<total_tiles, 256, 0, st>>>(reinterpret_cast<const ShadowJob*>(jobs), njobs, step_dev);
  V3_LAUNCH_CHECK();
  return VIT3D_OK;
}

}  // namespace vit3d
