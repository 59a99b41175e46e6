// fluxb200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T )      A, W bf16 (K contiguous), fp32 accumulation in TMEM.
//
// Replaces, on the hot path, every cuBLAS/cuBLASLt call the reference makes for a Linear
// (diffusion_rs_backend/src/unquantized/mod.rs:34-77, cublaslt/matmul.rs:502-586) and, in conv mode,
// the im2col + GEMM + NHWC->NCHW copy of diffusion_rs_common/src/core/cuda_backend/mod.rs:1545-1599.
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer   : 4-stage ring of {A 128x64, W 256x64} bf16 tiles, 128B swizzle
//   warp 1      MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, accumulators in TMEM
//   warps 2..9  epilogue       : tcgen05.ld -> fused bias / GELU / alpha / gate*x+residual in packed bf16x2 math
//                                (compile-time variants, no per-element branches) -> bf16 global stores
// TMEM holds two 128x256 fp32 accumulators so the epilogue of tile i overlaps the mainloop of tile i+1.
// A launch may carry up to 4 problems (grouped GEMM) so the 512-token text stream shares a wave with the
// 4096-token image stream instead of leaving 2/3 of the SMs idle.
#include <vector>

#include <stdlib.h>

#include "internal.h"
#include <cuda_fp16.h>

#include "ptx.cuh"

#ifndef FB_GEMM_TRACE
#define FB_GEMM_TRACE 0
#endif
#ifndef FB_FULL_WAIT_CLUSTER
#define FB_FULL_WAIT_CLUSTER 0  // 1: round-1 behaviour (acquire.cluster wait on the stage-full barrier), for A/B builds
#endif

namespace fb {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_N = 256;
static constexpr int BLOCK_K = 64;
static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
// CTA2 = a CTA pair (cta_group::2) computes one 256x256 tile: each CTA stages its own 128 rows of A and HALF of the
// 256 W rows, the MMA (M=256, issued by the leader) reads both halves -> 1.5x less L2->SM traffic per FLOP.
// QB = quantised B: every stage additionally holds the PACKED weights of the k-block (TMA-staged, <= 8 KB), which
// the dequant-producer warps expand into the stage's bf16 B tile.
static constexpr int PK_BYTES = 8192;
static constexpr int PK_AUX_OFF = 4096;  // second TMA box of a stage (absmax / Q4_K header) lands here
// CL4 = a cluster of two CTA pairs stacked along M computes a 512x256 tile (EXPERIMENT, off by default: the
// "gemm_cl4" flag).  Measured (scripts/gemm_trace.py): the MMA thread waits for TMA data 19-24 % of every tile.  In a
// CL4 cluster the two pairs need the same W tile: each CTA fetches only a quarter of it (64 rows) and TMA-multicasts
// it to its sibling in the other pair, which cuts the L2 reads by 25 %.  Result: bit-identical, but 8-27 % SLOWER - the
// limit is what each SM can ingest (its own 32 KB per k-block, however many CTAs requested it), not what L2 can
// serve, and only 33 clusters of 4 CTAs with 200 KB of shared memory are co-resident (132 of 148 SMs,
// scripts/ubench/cluster_probe.cu).  Data-ready signalling: every CTA arms its OWN full barrier with
// the 32 KB that land in its shared memory (own A, own B quarter, the sibling's B quarter); the non-leader CTA of a
// pair relays its barrier to the leader's `peer_full`; a stage is free again once BOTH pairs have consumed it
// (multicast commits from both leaders, empty barrier count 2).
template <bool CTA2, bool QB = false>
struct GemmCfg {
  static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + (QB ? PK_BYTES : 0);  // 32 (+8) KB (pair) / 48 KB (single)
  // 6 stages: a 7th (it fits) changes nothing (scripts/gemm_trace.py) - the feed is bound by each SM's ~51 B/clk of
  // achieved L2->SM ingest against the 64 B/clk a 128x256 tile needs at the MMA floor, not by bytes in flight.
  static constexpr int STAGES = QB ? 5 : (CTA2 ? 6 : 4);
  static constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + (QB ? 4096 : 0) /*code-pair LUTs*/;
};
static constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, then epilogue (+ dequant producer) warps
// dense B (TMA):      warps 2..9 = 8 epilogue warps
// quantised B (QB=1): warps 2..5 = 4 epilogue warps, warps 6..9 = 4 dequant-producer warps (one W row per thread)
template <bool QB> struct EpiCfg { static constexpr int WARPS = QB ? 4 : 8; };
static constexpr int MAX_PROBLEMS = 4;
static constexpr int GROUP_M = 8;
static constexpr int CONV_TH = 8, CONV_TW = 16;  // 8x16 output pixels = one 128-row M tile

struct alignas(64) GemmProblemDev {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  CUtensorMap tmap_b4;  // CL4: 64-row boxes of W
  int M, N, K, num_kb;
  int tiles_m, tiles_n, tile_begin, tile_end;
  int real_tiles_n;  // BIG kernel, N-direction: tiles_n counts 512-column positions, this is the number of 256-column tiles
  int sched_m;  // scheduling units along M: tiles_m (single CTA), ceil(tiles_m / 2) (CTA pair), ceil(tiles_m / 4) (CL4)
  int group_m;  // rasterisation: tiles are walked M-fastest inside groups of `group_m` M-units
  int conv, cH, cW, cC, c_chunks, ksize, tiles_h, tiles_w;
  bf16* out0;
  bf16* out1;
  long long ld0, ld1;
  int n_split, col_off1, ev0, ev1;  // EV_* epilogue variant of each output segment
  const bf16* bias;
  const bf16* gate;
  const bf16* res;
  long long gate_bstride;
  const int* step_ptr;  // denoising loop: device step counter; gate += *step_ptr * gate_step_stride
  long long gate_step_stride;
  int rows_per_batch, bias_mode, has_alpha;
  float alpha;
  // quantised B operand (fused dequant producer): per member a map over the packed bytes and one over the
  // auxiliary data (bnb: f32 absmax; Q4_K: the 16-byte block headers)
  CUtensorMap tmap_pk[4];
  CUtensorMap tmap_aux[4];
  QuantMember qm[4];
  int qcount;
  // fused QK-norm + RoPE epilogue (EV_QKROPE on segment 0)
  const bf16 *qk_wq, *qk_wk;
  const uint2* qk_pe2;
  long long qk_pe_bstride;
  bf16 *qk_Q, *qk_K, *qk_V;
  int qk_H, qk_L, qk_loff;
  float qk_eps;
};

struct GemmParams {
  GemmProblemDev p[MAX_PROBLEMS];
  int count;
  int total_tiles;
  // BIG kernel: tile_begin / sched_m of the problems count 512-row "positions"; the first n_big positions are computed
  // as 512x256 tiles, the remaining ones as their two 256x256 halves; total_items = n_big + 2 * (positions - n_big)
  int n_big, total_items;
  int big_ndir;  // BIG kernel: 0 = the two sub-tiles of an item are stacked along M (share W); 1 = side by side along N (share A)
  long long* trace;  // debug (fluxb200_debug_gemm_trace): per tile of scheduling unit 0, clock64 waits of the MMA thread
};

// ------------------------------------------------------------------------------------------------
// Epilogue math.  The reference rounds to bf16 after every tensor op; a bf16 HW multiply / add (HMUL2.BF16 /
// HADD2.BF16) is exactly "f32 op, then round-to-nearest-even to bf16" for bf16 inputs (products of two 8-bit
// mantissas are exact in f32; sums are exact in f32 whenever the smaller addend can matter), so the chain runs on
// packed bf16x2 registers with no conversions.
// ------------------------------------------------------------------------------------------------
// The _rn forms matter: nvcc contracts __hadd2(__hmul2(a, b), c) into one HFMA2.BF16 (a single rounding), which is not
// what the reference computes (gate * x and + residual are two tensor ops, two roundings).
typedef __nv_bfloat162 bf162;
__device__ __forceinline__ bf162 as_bf162(uint32_t u) { return *reinterpret_cast<bf162*>(&u); }
__device__ __forceinline__ uint32_t as_u32(bf162 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ bf162 bf162_const(float x) { return __float2bfloat162_rn(x); }

// tanh accurate to a few f32 ulps (far below bf16 resolution): 1 - 2/(exp(2|x|)+1), cubic near 0
__device__ __forceinline__ float tanh_f32(float x) {
  const float ax = fabsf(x);
  const float e = exp2f(ax * 2.8853900817779268f);
  float r = 1.0f - __fdividef(2.0f, e + 1.0f);
  const float x2 = ax * ax;
  const float small = ax * (1.0f - x2 * (0.3333333433f - 0.13333334f * x2));
  r = ax < 0.03125f ? small : r;
  return copysignf(r, x);
}

// tanh-GELU with the reference's bf16 op-by-op rounding (diffusion_rs_common/src/core/op.rs:539-578):
//   0.5*v*(1 + tanh(c*v*(1 + 0.044715*v*v)))   evaluated left to right, every product/sum rounded to bf16.
__device__ __forceinline__ bf162 gelu_bf16x2(bf162 v) {
  const bf162 kHalf = bf162_const(0.5f), kOne = bf162_const(1.0f);
  const bf162 kC = bf162_const(0.79788456080286535587989211986876373f), kK = bf162_const(0.044715f);
  const bf162 a = __hmul2_rn(kHalf, v);
  const bf162 p = __hadd2_rn(kOne, __hmul2_rn(__hmul2_rn(kK, v), v));
  const bf162 q = __hmul2_rn(__hmul2_rn(kC, v), p);
  const float2 qf = __bfloat1622float2(q);
  const bf162 t = __floats2bfloat162_rn(tanh_f32(qf.x), tanh_f32(qf.y));
  return __hmul2_rn(a, __hadd2_rn(kOne, t));
}
// scalar (slow-path) version, same arithmetic
__device__ __forceinline__ float gelu_bf16_steps(float v) {
  const bf162 r = gelu_bf16x2(__floats2bfloat162_rn(v, v));
  return __low2float(r);
}

enum { EV_PLAIN = 0, EV_GELU = 1, EV_RES = 2, EV_GATE_RES = 3, EV_ALPHA = 4, EV_KINDS = 5, EV_QKROPE = 5 };  // x bias mode (3)

// 16 consecutive output columns of one row: acc (fp32, from TMEM) -> bf16, fully unrolled, compile-time variant.
// The residual (the only operand that is unique per element, i.e. a real L2/HBM round trip) arrives in registers: it
// is prefetched one 32-column chunk ahead by epi_drain, the first chunk even before the accumulator is complete.
template <int BIAS, int EV>
__device__ __forceinline__ void epi16(const uint32_t* acc, bf16* outp, const bf16* biasp, const bf16* gatep,
                                      const uint4* resr, bf162 alpha2) {
  uint32_t b[8], g[8], r[8], o[8];
  if (BIAS != BIAS_NONE) {
    *reinterpret_cast<uint4*>(&b[0]) = *reinterpret_cast<const uint4*>(biasp);
    *reinterpret_cast<uint4*>(&b[4]) = *reinterpret_cast<const uint4*>(biasp + 8);
  }
  if (EV == EV_GATE_RES) {
    *reinterpret_cast<uint4*>(&g[0]) = *reinterpret_cast<const uint4*>(gatep);
    *reinterpret_cast<uint4*>(&g[4]) = *reinterpret_cast<const uint4*>(gatep + 8);
  }
  if (EV == EV_GATE_RES || EV == EV_RES) {
    *reinterpret_cast<uint4*>(&r[0]) = resr[0];
    *reinterpret_cast<uint4*>(&r[4]) = resr[1];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a0 = __uint_as_float(acc[2 * i]), a1 = __uint_as_float(acc[2 * i + 1]);
    bf162 v;
    if (BIAS == BIAS_FUSED) {  // bias enters the fp32 accumulator (cuBLASLt C operand, beta = 1): one rounding
      v = __floats2bfloat162_rn(a0 + bf_lo(b[i]), a1 + bf_hi(b[i]));
    } else {
      v = __floats2bfloat162_rn(a0, a1);
      if (BIAS == BIAS_AFTER_ROUND) v = __hadd2_rn(v, as_bf162(b[i]));  // separate bf16 broadcast_add
    }
    if (EV == EV_ALPHA) v = __hmul2_rn(v, alpha2);
    if (EV == EV_GELU) v = gelu_bf16x2(v);
    if (EV == EV_GATE_RES) v = __hmul2_rn(as_bf162(g[i]), v);
    if (EV == EV_GATE_RES || EV == EV_RES) v = __hadd2_rn(as_bf162(r[i]), v);
    o[i] = as_u32(v);
  }
  *reinterpret_cast<uint4*>(outp) = *reinterpret_cast<uint4*>(&o[0]);
  *reinterpret_cast<uint4*>(outp + 8) = *reinterpret_cast<uint4*>(&o[4]);
}

// generic per-element fallback for ragged N (final proj N=64 tail-free, conv_out N=3, ...): runtime flags
__device__ __noinline__ void epi_slow(const uint32_t* acc, int ncols, bf16* outp, const bf16* biasp, int bias_mode,
                                      int ev, const bf16* gatep, const bf16* resp, float alpha) {
  for (int e = 0; e < ncols; ++e) {
    float v = __uint_as_float(acc[e]);
    const float bv = biasp ? __bfloat162float(biasp[e]) : 0.f;
    if (bias_mode == BIAS_FUSED) {
      v = rbf(v + bv);
    } else {
      v = rbf(v);
      if (bias_mode == BIAS_AFTER_ROUND) v = rbf(v + bv);
    }
    if (ev == EV_ALPHA) v = rbf(v * alpha);
    if (ev == EV_GELU) v = gelu_bf16_steps(v);
    if (ev == EV_GATE_RES) v = rbf(__bfloat162float(gatep[e]) * v);
    if (ev == EV_GATE_RES || ev == EV_RES) v = rbf(__bfloat162float(resp[e]) + v);
    outp[e] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused q|k|v epilogue: one thread owns one token row x one head (128 accumulator columns in TMEM).
//   x = bf16(acc + bias);  q,k: y = bf16(bf16(x / sqrt(mean(x^2) + eps)) * w);  RoPE on interleaved pairs:
//   out0 = bf16(bf16(c*y0) + bf16(-s*y1)), out1 = bf16(bf16(s*y0) + bf16(c*y1));  v: copied.  Written to [B,H,L,128].
// Same rounding points as qknorm_rope_kernel (elementwise.cu) / the reference (model.rs:86-95, layer_norm.rs:136-153).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_qkrope(const GemmProblemDev& p, uint32_t t_half, int n_half0, long long grow,
                                           bool valid) {
  const int D = p.qk_H * 128;
  const int which = n_half0 / D;  // 0 q, 1 k, 2 v
  const int h = (n_half0 - which * D) >> 7;
  uint32_t x[64];  // the head's 128 values, packed bf16x2
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // 16 columns at a time keeps the live register set small
    uint32_t a[16];
    tmem_ld16(t_half + c * 16, a);
    tc_wait_ld();
    uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
    if (p.bias_mode != BIAS_NONE) {
      const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n_half0 + c * 16);
      b0 = bp[0], b1 = bp[1];
    }
    const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a0 = __uint_as_float(a[2 * i]), a1 = __uint_as_float(a[2 * i + 1]);
      bf162 v;
      if (p.bias_mode == BIAS_FUSED) {
        v = __floats2bfloat162_rn(a0 + bf_lo(b[i]), a1 + bf_hi(b[i]));
      } else {
        v = __floats2bfloat162_rn(a0, a1);
        if (p.bias_mode == BIAS_AFTER_ROUND) v = __hadd2_rn(v, as_bf162(b[i]));
      }
      x[c * 8 + i] = as_u32(v);
      const float2 f = __bfloat1622float2(v);
      ss = fmaf(f.x, f.x, ss);
      ss = fmaf(f.y, f.y, ss);
    }
  }
  if (!valid) return;
  const long long b_idx = p.rows_per_batch > 0 ? grow / p.rows_per_batch : 0;
  const int l = p.qk_loff + static_cast<int>(grow - b_idx * p.rows_per_batch);
  bf16* dst = (which == 0 ? p.qk_Q : (which == 1 ? p.qk_K : p.qk_V)) + ((b_idx * p.qk_H + h) * p.qk_L + l) * 128;
  if (which < 2) {
    const float denom = sqrtf(ss / 128.0f + p.qk_eps);
    const float inv = 1.0f / denom;
    const uint4* wp = reinterpret_cast<const uint4*>(which == 0 ? p.qk_wq : p.qk_wk);
    // pe2 is laid out [batch][pair][token]: the 32 lanes of a warp (consecutive tokens) read 256 contiguous bytes
    const uint2* pe = p.qk_pe2 + b_idx * p.qk_pe_bstride + l;
#pragma unroll
    for (int i4 = 0; i4 < 16; ++i4) {
      const uint4 w4 = wp[i4];  // 4 pairs of norm weights
      const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i4 * 4 + e;
        const float2 f = __bfloat1622float2(as_bf162(x[i]));
        // x / denom, correctly rounded via one Newton step on the reciprocal estimate (3 FMAs instead of a div.rn)
        float q0 = f.x * inv, q1 = f.y * inv;
        q0 = fmaf(fmaf(-q0, denom, f.x), inv, q0);
        q1 = fmaf(fmaf(-q1, denom, f.y), inv, q1);
        const bf162 y = __hmul2(__floats2bfloat162_rn(q0, q1), as_bf162(wv[e]));
        const uint32_t yu = as_u32(y);
        const uint32_t y00 = __byte_perm(yu, yu, 0x1010), y11 = __byte_perm(yu, yu, 0x3232);
        const uint2 cs = pe[static_cast<long long>(i) * p.qk_L];  // cs.x = (cos, sin), cs.y = (-sin, cos)
        x[i] = as_u32(__hadd2_rn(__hmul2_rn(as_bf162(cs.x), as_bf162(y00)), __hmul2_rn(as_bf162(cs.y), as_bf162(y11))));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

struct TileCoord {
  int prob, m_t, n_t;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& P, int t) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < MAX_PROBLEMS; ++i)
    if (i < P.count && t >= P.p[i].tile_begin) pi = i;
  const GemmProblemDev& p = P.p[pi];
  int lt = t - p.tile_begin;
  int group_size = p.group_m * p.tiles_n;
  int g = lt / group_size;
  int first_m = g * p.group_m;
  int gm = min(p.sched_m - first_m, p.group_m);
  int in_g = lt - g * group_size;
  TileCoord c;
  c.prob = pi;
  c.m_t = first_m + in_g % gm;
  c.n_t = in_g / gm;
  return c;
}

// ------------------------------------------------------------------------------------------------
// Fused de-quantisation producer.  Thread `row` owns row `row` of this CTA's B tile (128 rows in pair mode): per
// k-block it reads that row's 64 packed weights from HBM, expands them with the reference's arithmetic
//   NF4/FP4 : bf16(code[nibble] * absmax[block])           (kernels/bitsandbytes/dequant.cu:94-170)
//   Q4_K    : bf16(f16(d*sc*q - dmin*m))                    (k_quants.rs:1568-1599, gguf/mod.rs:29-31)
//   int8    : bf16(float(w) * SCB[row] / 127)               (dequant.cu:205-214)
// and writes the 128-byte row into the K-major 128B-swizzled smem tile the MMA consumes (chunk c of row r lives at
// r*128 + ((c ^ (r & 7)) << 4)), then fence.proxy.async + one mbarrier arrival per warp.
// ------------------------------------------------------------------------------------------------
__constant__ float c_nf4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};
__constant__ float c_fp4[16] = {0.0f,  5.208333333e-03f,  0.66666667f,  1.0f,  0.33333333f,  0.5f,  0.16666667f,  0.25f,
                                -0.0f, -5.208333333e-03f, -0.66666667f, -1.0f, -0.33333333f, -0.5f, -0.16666667f, -0.25f};

// packed data of one k-block (64 weights) of one W row, per format
template <int KIND> struct DqRegs;
template <> struct DqRegs<QB_NF4> { uint4 a, b; float am; };          // 32 nibble bytes + absmax (also FP4)
template <> struct DqRegs<QB_Q4K> { uint4 hdr, a, b; };                // d, dmin, 12 scale bytes + 32 q bytes
template <> struct DqRegs<QB_INT8> { uint4 a, b, c, d; };              // 64 int8

// read this thread's row of the stage's TMA-staged packed data (shared memory, conflict-free 16-byte reads)
template <int KIND>
__device__ __forceinline__ void dq_read(const uint8_t* pk, int row, DqRegs<KIND>& r, int aux_sel) {
  if constexpr (KIND == QB_NF4) {
    const uint4* src = reinterpret_cast<const uint4*>(pk + row * 32);
    r.a = src[0], r.b = src[1];
    r.am = *reinterpret_cast<const float*>(pk + PK_AUX_OFF + row * 16 + aux_sel * 4);  // aligned group of 4 absmax
  } else if constexpr (KIND == QB_Q4K) {
    const uint4* src = reinterpret_cast<const uint4*>(pk + row * 32);
    r.a = src[0], r.b = src[1];
    r.hdr = *reinterpret_cast<const uint4*>(pk + PK_AUX_OFF + row * 16);
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(pk + row * 64);
    r.a = src[0], r.b = src[1], r.c = src[2], r.d = src[3];
  }
}

// expand one k-block of one row to 64 bf16 (32 packed words)
template <int KIND>
__device__ __forceinline__ void dq_decode(const DqRegs<KIND>& r, uint32_t* o, const float2* lut, int j, float row_scale) {
  if constexpr (KIND == QB_NF4) {
    const uint32_t w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float2 cp = lut[(w[i >> 2] >> (8 * (i & 3))) & 0xffu];
      o[i] = pack_bf16(cp.x * r.am, cp.y * r.am);
    }
  } else if constexpr (KIND == QB_Q4K) {
    const __half2 dm = *reinterpret_cast<const __half2*>(&r.hdr.x);
    const float d = __low2float(dm), dmin = __high2float(dm);
    // 12 scale bytes = hdr.y, hdr.z, hdr.w
    auto sbyte = [&](int idx) -> uint32_t {
      const uint32_t word = idx < 4 ? r.hdr.y : (idx < 8 ? r.hdr.z : r.hdr.w);
      return (word >> (8 * (idx & 3))) & 0xffu;
    };
    float dd[2], mm[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int is = 2 * j + h;
      uint32_t s6, m6;
      if (is < 4) {
        s6 = sbyte(is) & 63, m6 = sbyte(is + 4) & 63;
      } else {
        s6 = (sbyte(is + 4) & 0xF) | ((sbyte(is - 4) >> 6) << 4);
        m6 = (sbyte(is + 4) >> 4) | ((sbyte(is) >> 6) << 4);
      }
      dd[h] = __fmul_rn(d, static_cast<float>(s6));
      mm[h] = __fmul_rn(dmin, static_cast<float>(m6));
    }
    const uint32_t w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) {  // bytes 2i, 2i+1: low nibbles -> elements 2i, 2i+1; high nibbles -> 32 + 2i, 33 + 2i
      const uint32_t b0 = (w[i >> 1] >> (16 * (i & 1))) & 0xffu, b1 = (w[i >> 1] >> (16 * (i & 1) + 8)) & 0xffu;
      const float l0 = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(dd[0], static_cast<float>(b0 & 15)), mm[0])));
      const float l1 = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(dd[0], static_cast<float>(b1 & 15)), mm[0])));
      const float h0 = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(dd[1], static_cast<float>(b0 >> 4)), mm[1])));
      const float h1 = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(dd[1], static_cast<float>(b1 >> 4)), mm[1])));
      o[i] = pack_bf16(l0, l1);
      o[16 + i] = pack_bf16(h0, h1);
    }
  } else {
    const uint32_t w[16] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w,
                            r.c.x, r.c.y, r.c.z, r.c.w, r.d.x, r.d.y, r.d.z, r.d.w};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int v0 = static_cast<int8_t>((w[i >> 1] >> (16 * (i & 1))) & 0xffu);
      const int v1 = static_cast<int8_t>((w[i >> 1] >> (16 * (i & 1) + 8)) & 0xffu);
      o[i] = pack_bf16((static_cast<float>(v0) * row_scale) / 127.f, (static_cast<float>(v1) * row_scale) / 127.f);
    }
  }
}

// all k-blocks of one tile for one row, format fixed at compile time.  The packed bytes arrive in the stage's PK
// area by TMA (issued by warp 0 together with the A tile), so these threads never have a global load in flight when
// they execute fence.proxy.async (which would otherwise drain it and serialise every k-block on HBM latency).
template <bool CTA2, int KIND>
__device__ __forceinline__ void dq_tile(const QuantMember& qm, bool in_range, int num_kb, int n_local, uint8_t* smem,
                                        uint64_t* full_bar, uint64_t* pk_full, int row, const float2* lut2, int STAGES,
                                        int STAGE_BYTES, int& stage, uint32_t& phase) {
  const int lane = threadIdx.x & 31;
  const float2* lut = lut2 + (qm.kind == QB_FP4 ? 256 : 0);
  const float row_scale = (KIND == QB_INT8 && in_range) ? __ldg(qm.scb + n_local) : 0.f;
  const int bs_shift = 31 - __clz(qm.blocksize);  // blocksizes are powers of two (bitsandbytes/mod.rs:14)
  for (int kb = 0; kb < num_kb; ++kb) {
    uint8_t* sa = smem + stage * STAGE_BYTES;
    uint8_t* sb = sa + A_BYTES + row * 128;
    const uint8_t* pk = sa + STAGE_BYTES - PK_BYTES;
    // pk_full completes only after the TMA producer has seen empty_bar[stage]: the B tile slot is free as well
    mbar_wait(&pk_full[stage], phase);
    DqRegs<KIND> r;
    dq_read<KIND>(pk, row, r, ((kb * BLOCK_K) >> bs_shift) & 3);
    uint32_t o[32];  // 64 bf16
    dq_decode<KIND>(r, o, lut, kb & 3, row_scale);
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(sb + ((c ^ sw) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) {
      if (CTA2) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&full_bar[stage]), 0));
      else      mbar_arrive(&full_bar[stage]);
    }
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
}

template <bool CTA2>
__device__ __forceinline__ void dequant_producer(const GemmParams& P, uint8_t* smem, uint64_t* full_bar,
                                                 uint64_t* pk_full, int unit_id, int num_units, uint32_t cta_rank,
                                                 int row, const float2* lut2, int STAGES, int STAGE_BYTES, int B_ROWS) {
  int stage = 0;
  uint32_t phase = 0;
  for (int t = unit_id; t < P.total_tiles; t += num_units) {
    TileCoord tc = decode_tile(P, t);
    const GemmProblemDev& p = P.p[tc.prob];
    const int n = tc.n_t * BLOCK_N + (CTA2 ? static_cast<int>(cta_rank) * B_ROWS : 0) + row;  // row of W
    const bool in_range = n < p.N;
    int mi = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (i < p.qcount && n >= p.qm[i].row_begin) mi = i;
    const QuantMember& qm = p.qm[mi];
    const int n_local = n - qm.row_begin;
    switch (qm.kind) {  // uniform across the CTA: a tile never straddles two members
      case QB_Q4K:
        dq_tile<CTA2, QB_Q4K>(qm, in_range, p.num_kb, n_local, smem, full_bar, pk_full, row, lut2, STAGES, STAGE_BYTES,
                              stage, phase);
        break;
      case QB_INT8:
        dq_tile<CTA2, QB_INT8>(qm, in_range, p.num_kb, n_local, smem, full_bar, pk_full, row, lut2, STAGES, STAGE_BYTES,
                               stage, phase);
        break;
      default:
        dq_tile<CTA2, QB_NF4>(qm, in_range, p.num_kb, n_local, smem, full_bar, pk_full, row, lut2, STAGES, STAGE_BYTES,
                              stage, phase);
        break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128-row x 256-column accumulator (TMEM columns [t_acc, t_acc + 256)) by the calling warp: lane quarter
// q = warp % 4 (hardware rule), column half chosen by the warp's index (8 epilogue warps) or both halves in turn (4).
// Waits for the accumulator itself (`full_bar` / `parity`) AFTER it has issued the loads of the first residual chunk;
// the caller releases the accumulator afterwards.
// ------------------------------------------------------------------------------------------------
template <int EPI_WARPS, bool QKROPE = true>
__device__ __forceinline__ void epi_drain(const GemmProblemDev& p, int n_t, int m_t, uint32_t t_acc, int warp, int lane,
                                          uint64_t* full_bar, uint32_t parity) {
  const int q = warp & 3;
  const int r = q * 32 + lane;  // row inside the tile == TMEM lane
  long long grow;
  bool valid;
  if (p.conv) {
    int per_img = p.tiles_h * p.tiles_w;
    int cn = m_t / per_img;
    int rem = m_t - cn * per_img;
    int h = (rem / p.tiles_w) * CONV_TH + r / CONV_TW;
    int w = (rem % p.tiles_w) * CONV_TW + r % CONV_TW;
    valid = (m_t < p.tiles_m) && (h < p.cH) && (w < p.cW);
    grow = (static_cast<long long>(cn) * p.cH + h) * p.cW + w;
  } else {
    grow = static_cast<long long>(m_t) * BLOCK_M + r;
    valid = grow < p.M;
  }
  const long long gate_off =
      (p.gate != nullptr) ? (p.rows_per_batch > 0 ? grow / p.rows_per_batch : 0) * p.gate_bstride +
                                (p.step_ptr ? *p.step_ptr * p.gate_step_stride : 0)
                          : 0;
  const bf162 alpha2 = __float2bfloat162_rn(p.alpha);
  const int bias_mode = p.bias_mode;
  const bool res_ev = valid && (p.ev0 == EV_RES || p.ev0 == EV_GATE_RES);  // only segment 0 can carry a residual

  // 8 epilogue warps: each drains one 128-column half; 4 epilogue warps (quantised-B kernel): both halves in turn
#pragma unroll 1
  for (int hh = 0; hh < (EPI_WARPS == 4 ? 2 : 1); ++hh) {
  const int chalf = EPI_WARPS == 4 ? hh : ((warp - 2) >> 2);
  const uint32_t t_row = t_acc + chalf * 128 + (static_cast<uint32_t>(q * 32) << 16);
  const int n_half0 = n_t * BLOCK_N + chalf * 128;
  const bool half_seg1 = (p.n_split > 0) && (n_half0 >= p.n_split);
  // residual of the NEXT 32-column chunk of this thread's row (full chunks of segment 0 only)
  uint4 rn[4];
  auto prefetch_res = [&](int chunk) {
    const int n0 = n_half0 + chunk * 32;
    if (res_ev && chunk < 4 && n0 + 32 <= p.N && !((p.n_split > 0) && (n0 >= p.n_split))) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + grow * p.ld0 + n0);
#pragma unroll
      for (int i = 0; i < 4; ++i) rn[i] = rp[i];
    }
  };
  prefetch_res(0);
  if (hh == 0) {
    mbar_wait(full_bar, parity);
    tc_fence_after();
  }
  if (QKROPE && p.ev0 == EV_QKROPE && !half_seg1 && n_half0 < p.N) {
    epi_qkrope(p, t_row, n_half0, grow, valid);
  } else
#pragma unroll 1
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int n0 = n_t * BLOCK_N + chalf * 128 + chunk * 32;
    if (n0 >= p.N) break;  // warp-uniform
    const uint4 rc[4] = {rn[0], rn[1], rn[2], rn[3]};
    prefetch_res(chunk + 1);
    uint32_t acc_r[32];
    tmem_ld32(t_row + chunk * 32, acc_r);
    tc_wait_ld();
    if (!valid) continue;
    const bool seg1 = (p.n_split > 0) && (n0 >= p.n_split);
    bf16* outp = seg1 ? p.out1 + grow * p.ld1 + (n0 - p.n_split + p.col_off1) : p.out0 + grow * p.ld0 + n0;
    const int ev = seg1 ? p.ev1 : p.ev0;
    const bf16* resp = (ev == EV_RES || ev == EV_GATE_RES) ? p.res + grow * p.ld0 + n0 : nullptr;
    const bf16* gatep = (ev == EV_GATE_RES) ? p.gate + gate_off + n0 : nullptr;
    const bf16* biasp = (bias_mode != BIAS_NONE) ? p.bias + n0 : nullptr;
    if (n0 + 32 <= p.N) {
      // warp-uniform dispatch to a fully specialised 2 x 16-column body
#define FB_EPI_CASE(B, E)                                                                  \
  case (B) * EV_KINDS + (E):                                                               \
epi16<B, E>(acc_r, outp, biasp, gatep, rc, alpha2);                                    \
epi16<B, E>(acc_r + 16, outp + 16, biasp ? biasp + 16 : nullptr, gatep ? gatep + 16 : nullptr, \
            rc + 2, alpha2);                                                           \
break;
#define FB_EPI_BIAS(B) \
  FB_EPI_CASE(B, EV_PLAIN) FB_EPI_CASE(B, EV_GELU) FB_EPI_CASE(B, EV_RES) FB_EPI_CASE(B, EV_GATE_RES) FB_EPI_CASE(B, EV_ALPHA)
      switch (bias_mode * EV_KINDS + ev) {
        FB_EPI_BIAS(BIAS_NONE)
        FB_EPI_BIAS(BIAS_FUSED)
        FB_EPI_BIAS(BIAS_AFTER_ROUND)
        default:
          break;
      }
#undef FB_EPI_BIAS
#undef FB_EPI_CASE
    } else {
      epi_slow(acc_r, p.N - n0, outp, biasp, bias_mode, ev, gatep, resp, p.alpha);
    }
  }
  }  // hh
}

template <bool CTA2, bool QB, bool CL4 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams P) {
  static_assert(!CL4 || (CTA2 && !QB), "CL4 is a pair-of-pairs build for dense weights");
  using Cfg = GemmCfg<CTA2, QB>;
  constexpr int EPI_WARPS = EpiCfg<QB>::WARPS;
  constexpr int DQ_WARPS = QB ? 4 : 0;  // arrivals on a full barrier: 1 (TMA expect_tx) + dequant warps (of both CTAs)
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  const uint32_t rank_cl = CTA2 ? cluster_ctarank() : 0;        // rank inside the cluster (0..1, CL4: 0..3)
  const uint32_t cta_rank = rank_cl & 1;                        // rank inside the CTA pair
  const uint32_t pair_id = CL4 ? (rank_cl >> 1) : 0;            // CL4: which of the two pairs of the cluster
  const uint32_t leader_cl = rank_cl & ~1u;                     // cluster rank of this pair's leader CTA
  constexpr int UNIT_CTAS = CL4 ? 4 : (CTA2 ? 2 : 1);
  const int unit_id = blockIdx.x / UNIT_CTAS;                   // tile-scheduling unit (CTA, CTA pair or cluster of 4)
  const int num_units = gridDim.x / UNIT_CTAS;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* pk_full = tmem_empty + 2;        // [STAGES] packed weights of the stage have landed (QB only)
  uint64_t* peer_full = pk_full + STAGES;    // [STAGES] CL4: the non-leader CTA's stage has landed (relayed arrival)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(peer_full + STAGES);
  float2* lut2 = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 256);  // [2][256] byte -> (hi, lo) code pairs

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.count; ++i) {
      tma_prefetch_desc(&P.p[i].tmap_a);
      tma_prefetch_desc(&P.p[i].tmap_b);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + (CTA2 ? 2 : 1) * DQ_WARPS);
      mbar_init(&empty_bar[s], CL4 ? 2 : 1);  // CL4: both pairs' MMAs must have consumed the stage
      mbar_init(&pk_full[s], 1);
      mbar_init(&peer_full[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CTA2 ? 2 * EPI_WARPS : EPI_WARPS);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc_2sm(tmem_base_slot, 512); else tmem_alloc(tmem_base_slot, 512);
  }
  if (QB) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      const int b = i & 255;
      const float* cb = (i < 256) ? c_nf4 : c_fp4;
      lut2[i] = make_float2(cb[b >> 4], cb[b & 15]);  // high nibble is the first weight (dequant.cu:155-156)
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();  // peer barriers must be initialised before remote arrivals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // everything above overlapped the previous kernel's tail; its outputs (our A operand, residual, ...) are valid from here
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit_id; t < P.total_tiles; t += num_units) {
        TileCoord tc = decode_tile(P, t);
        const GemmProblemDev& p = P.p[tc.prob];
        const int m_t = UNIT_CTAS * tc.m_t + static_cast<int>(rank_cl);
        const int b_row0 = tc.n_t * BLOCK_N + (CTA2 ? static_cast<int>(cta_rank) * Cfg::B_ROWS : 0);
        int cn = 0, ch0 = 0, cw0 = 0;
        if (p.conv) {
          int per_img = p.tiles_h * p.tiles_w;
          cn = m_t / per_img;
          int rem = m_t - cn * per_img;
          ch0 = (rem / p.tiles_w) * CONV_TH;
          cw0 = (rem % p.tiles_w) * CONV_TW;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          int a_c0, b_c0;
          int tap = 0, kh = 0, kw = 0, pad = 0;
          if (p.conv) {
            tap = kb / p.c_chunks;
            const int cc = kb - tap * p.c_chunks;
            kh = tap / p.ksize, kw = tap - kh * p.ksize;
            pad = p.ksize >> 1;
            a_c0 = cc * BLOCK_K;
            b_c0 = tap * p.cC + cc * BLOCK_K;
          } else {
            a_c0 = b_c0 = kb * BLOCK_K;
          }
          if (QB) {
            // packed weights of this CTA's B rows for this k-block -> the stage's PK area (local barrier)
            int mi = 0;
#pragma unroll
            for (int i = 1; i < 4; ++i)
              if (i < p.qcount && b_row0 >= p.qm[i].row_begin) mi = i;
            const int kind = p.qm[mi].kind;
            const int r0 = b_row0 - p.qm[mi].row_begin;
            uint8_t* pk = sb + Cfg::B_BYTES;
            if (kind == QB_INT8) {
              mbar_arrive_expect_tx(&pk_full[stage], Cfg::B_ROWS * 64);
              tma_load_2d(pk, &p.tmap_pk[mi], &pk_full[stage], kb * 64, r0);
            } else if (kind == QB_Q4K) {
              mbar_arrive_expect_tx(&pk_full[stage], Cfg::B_ROWS * 48);
              tma_load_2d(pk, &p.tmap_pk[mi], &pk_full[stage], (kb >> 2) * 144 + 16 + (kb & 3) * 32, r0);
              tma_load_2d(pk + PK_AUX_OFF, &p.tmap_aux[mi], &pk_full[stage], (kb >> 2) * 144, r0);
            } else {
              const int bs_shift = 31 - __clz(p.qm[mi].blocksize);
              mbar_arrive_expect_tx(&pk_full[stage], Cfg::B_ROWS * 48);
              tma_load_2d(pk, &p.tmap_pk[mi], &pk_full[stage], kb * 32, r0);
              // TMA box starts must be 16-byte aligned: fetch the aligned group of 4 absmax values, the consumer picks
              tma_load_2d(pk + PK_AUX_OFF, &p.tmap_aux[mi], &pk_full[stage], (((kb * BLOCK_K) >> bs_shift) & ~3) * 4, r0);
            }
          }
          if (CL4) {
            // own A tile + own quarter of W (multicast to the sibling CTA of the other pair, which sends us its quarter)
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(sa, &p.tmap_a, &full_bar[stage], a_c0, m_t * BLOCK_M);
            const uint16_t mask = static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2)));
            tma_load_2d_mcast(sb + pair_id * (Cfg::B_BYTES / 2), &p.tmap_b4, &full_bar[stage], b_c0,
                              b_row0 + static_cast<int>(pair_id) * (Cfg::B_ROWS / 2), mask);
          } else if (CTA2) {
            // both CTAs' bytes are credited to the leader's barrier; only the leader arms it
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (QB ? A_BYTES : STAGE_BYTES));
            if (p.conv) tma_load_4d_2sm(sa, &p.tmap_a, bar, a_c0, cw0 + kw - pad, ch0 + kh - pad, cn);
            else        tma_load_2d_2sm(sa, &p.tmap_a, bar, a_c0, m_t * BLOCK_M);
            if (!QB) tma_load_2d_2sm(sb, &p.tmap_b, bar, b_c0, b_row0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], QB ? A_BYTES : STAGE_BYTES);
            if (p.conv) tma_load_4d(sa, &p.tmap_a, &full_bar[stage], a_c0, cw0 + kw - pad, ch0 + kh - pad, cn);
            else        tma_load_2d(sa, &p.tmap_a, &full_bar[stage], a_c0, m_t * BLOCK_M);
            if (!QB) tma_load_2d(sb, &p.tmap_b, &full_bar[stage], b_c0, b_row0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (single thread; in pair mode only the leader CTA issues) =================
    if (CL4 && lane == 0 && cta_rank == 1) {
      // relay: tell the pair's leader when this CTA's stage (A tile + its half of W) has landed
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit_id; t < P.total_tiles; t += num_units) {
        const GemmProblemDev& p = P.p[decode_tile(P, t).prob];
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          // relaxed: the payload was written by the async proxy (TMA) into THIS CTA's shared memory and is read there by
          // the tensor core; a release.cluster arrive per k-block would drain this thread's loads and flush L1 (~1 us)
          mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&peer_full[stage]), leader_cl));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
#if FB_GEMM_TRACE
      long long* const trace = (unit_id == 0) ? P.trace : nullptr;
#else
      constexpr long long* trace = nullptr;  // tracing is compiled out of production builds (FLUXB200_GEMM_TRACE=1 python -m ...build)
#endif
      int tile_no = 0;
      for (int t = unit_id; t < P.total_tiles; t += num_units, ++tile_no) {
        TileCoord tc = decode_tile(P, t);
        const GemmProblemDev& p = P.p[tc.prob];
        const long long c0 = trace ? clock64() : 0;
        if (CTA2) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1); else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const long long c1 = trace ? clock64() : 0;
        long long full_wait = 0;
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const long long w0 = trace ? clock64() : 0;
          if (CL4) {
            mbar_wait(&full_bar[stage], phase);
            mbar_wait_cluster(&peer_full[stage], phase);
          } else if (CTA2) {
            // CTA-scope wait: the stage was written by TMA (async proxy, both CTAs' bytes credited to this barrier) and
            // is read by tcgen05.mma (async proxy); the barrier's completion orders the two.  An acquire.cluster
            // try_wait is a cluster-scope fence on EVERY k-block and showed up as "MMA thread waits for TMA data".
            if (FB_FULL_WAIT_CLUSTER) mbar_wait_cluster(&full_bar[stage], phase); else mbar_wait(&full_bar[stage], phase);
          } else {
            mbar_wait(&full_bar[stage], phase);
          }
          tc_fence_after();
          if (trace) full_wait += clock64() - w0;
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = umma_smem_desc_sw128(sa + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            if (CTA2) umma_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else      umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs) once these MMAs have read it
          if (CL4) tc_commit_2sm(&empty_bar[stage], 0xF);
          else if (CTA2) tc_commit_2sm(&empty_bar[stage], 3);
          else tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CTA2) tc_commit_2sm(&tmem_full[acc], static_cast<uint16_t>(3u << leader_cl)); else tc_commit(&tmem_full[acc]);
        if (trace && tile_no < 64) {
          trace[tile_no * 4 + 0] = c0;                 // tile start
          trace[tile_no * 4 + 1] = c1 - c0;            // waited for the epilogue to free the accumulator
          trace[tile_no * 4 + 2] = full_wait;          // waited for TMA data, summed over the k-blocks
          trace[tile_no * 4 + 3] = clock64() - c0;     // tile total (issue side)
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (QB && warp >= 2 + EPI_WARPS) {
    // ================= dequant producer warps (quantised B): one W row per thread, 64 weights per k-block =========
    dequant_producer<CTA2>(P, smem, full_bar, pk_full, unit_id, num_units, cta_rank, (warp - 2 - EPI_WARPS) * 32 + lane,
                           lut2, STAGES, STAGE_BYTES, Cfg::B_ROWS);
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;              // TMEM lane quarter this warp may access (hardware: warp_id % 4)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = unit_id; t < P.total_tiles; t += num_units) {
      TileCoord tc = decode_tile(P, t);
      const GemmProblemDev& p = P.p[tc.prob];
      const int m_t = UNIT_CTAS * tc.m_t + static_cast<int>(rank_cl);
      epi_drain<EPI_WARPS>(p, tc.n_t, m_t, tmem_base + acc * BLOCK_N, warp, lane, &tmem_full[acc], acc_phase);
      // release the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), leader_cl));  // the leader's MMA thread waits on it
        else      mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// BIG tiles for the long-K GEMMs (MLP-down K = 12288, single-block linear2 K = 15360).
//
// Measured on the 256x256-per-pair kernel above (scripts/gemm_trace.py): the MMA thread waits for TMA data 16-24 % of
// every tile; an SM has to ingest 32 KB per 512 MMA clocks (64 B/clk) and gets ~51.  This kernel halves the W bytes per
// FLOP: a CTA pair computes a 512x256 tile as TWO M = 256 MMAs per k-step that share one W tile, i.e. per SM and
// k-block 2 x 16 KB of A + 16 KB of W for 1024 MMA clocks (48 B/clk).  Both 128x256 fp32 accumulators of the tile fill
// the CTA's 512 TMEM columns, so the epilogue can no longer hide behind the next tile's main loop; with K >= 8192 the
// main loop is 200-250 k clocks per tile and the exposed epilogue (a few k clocks, partly overlapped: accumulator 0
// is released and refilled while accumulator 1 drains) is small against the TMA wait it removes.
//
// Wave quantisation: 4608x3072 is 108 such tiles on 74 CTA pairs.  The work list is therefore hybrid: as many FULL
// waves of 512x256 tiles as fit (positions [0, n_big)), and the remaining positions as their two 256x256 halves
// ("small" items: one accumulator, alternating so that they double-buffer like the kernel above).  Items are dealt
// round-robin: for linear2, every pair computes one BIG tile and 68 of the 74 pairs one small tile.
// Dense weights, no conv mode; any epilogue variant (it reuses epi_drain).
// ------------------------------------------------------------------------------------------------
struct BigCfg {
  static constexpr int STAGES = 4;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;     // this CTA's half of the 256 W rows: 16 KB
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;       // [A sub-tile 0][A sub-tile 1][W half] = 48 KB
  static constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

struct BigItem {
  int prob, nsub;      // nsub = 0: nothing to do (second half of an odd edge)
  int blk[2], nt[2];   // sub-tile s = 256-row block blk[s] x 256-column tile nt[s]
};

__device__ __forceinline__ BigItem decode_big_item(const GemmParams& P, int w) {
  int pos, sub_sel = -1;
  if (w < P.n_big) {
    pos = w;
  } else {
    const int r = w - P.n_big;
    pos = P.n_big + (r >> 1), sub_sel = r & 1;
  }
  const TileCoord tc = decode_tile(P, pos);
  const GemmProblemDev& p = P.p[tc.prob];
  BigItem it;
  it.prob = tc.prob;
  int first, count;  // index of the item's first sub-tile along the doubled direction, number of sub-tiles there
  if (P.big_ndir) {
    first = 2 * tc.n_t, count = p.real_tiles_n;
  } else {
    first = 2 * tc.m_t, count = (p.tiles_m + 1) / 2;  // 256-row blocks of the problem
  }
  const bool have2 = first + 1 < count;
  int i0;
  if (sub_sel < 0) {
    i0 = first, it.nsub = have2 ? 2 : 1;
  } else {
    i0 = first + sub_sel, it.nsub = (sub_sel == 0 || have2) ? 1 : 0;
  }
  if (P.big_ndir) {
    it.blk[0] = it.blk[1] = tc.m_t, it.nt[0] = i0, it.nt[1] = i0 + 1;
  } else {
    it.blk[0] = i0, it.blk[1] = i0 + 1, it.nt[0] = it.nt[1] = tc.n_t;
  }
  return it;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_big_kernel(const __grid_constant__ GemmParams P) {
  constexpr int STAGES = BigCfg::STAGES;
  constexpr int STAGE_BYTES = BigCfg::STAGE_BYTES;
  constexpr int EPI_WARPS = 8;
  const uint32_t cta_rank = cluster_ctarank();
  const int unit_id = blockIdx.x / 2;
  const int num_units = gridDim.x / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;  // [2] one per accumulator
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.count; ++i) {
      tma_prefetch_desc(&P.p[i].tmap_a);
      tma_prefetch_desc(&P.p[i].tmap_b);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_base_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bar_base = mapa_u32(smem_u32(full_bar), 0);  // the leader's barriers collect both CTAs' bytes
      for (int w = unit_id; w < P.total_items; w += num_units) {
        const BigItem it = decode_big_item(P, w);
        if (it.nsub == 0) continue;
        const GemmProblemDev& p = P.p[it.prob];
        const bool nd = P.big_ndir != 0;
        // stage = three 16 KB slots: M direction [A sub 0][A sub 1][W]; N direction [A][W sub 0][W sub 1]
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          const uint32_t bar = bar_base + stage * 8;
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (it.nsub * A_BYTES + BigCfg::B_BYTES));
          for (int s = 0; s < (nd ? 1 : it.nsub); ++s)
            tma_load_2d_2sm(sa + s * A_BYTES, &p.tmap_a, bar, kb * BLOCK_K,
                            (2 * it.blk[s] + static_cast<int>(cta_rank)) * BLOCK_M);
          for (int s = 0; s < (nd ? it.nsub : 1); ++s)
            tma_load_2d_2sm(sa + (nd ? 1 + s : 2) * A_BYTES, &p.tmap_b, bar, kb * BLOCK_K,
                            it.nt[s] * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / 2));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA) =================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_phase = 0;  // bit a: phase of accumulator a
      int small_count = 0;
#if FB_GEMM_TRACE
      long long* const trace = (unit_id == 0) ? P.trace : nullptr;
#else
      constexpr long long* trace = nullptr;
#endif
      int item_no = 0;
      for (int w = unit_id; w < P.total_items; w += num_units) {
        const BigItem it = decode_big_item(P, w);
        if (it.nsub == 0) continue;
        const GemmProblemDev& p = P.p[it.prob];
        int acc0 = 0;
        if (it.nsub == 2) small_count = 0; else acc0 = small_count++ & 1;
        const long long c0 = trace ? clock64() : 0;
        long long full_wait = 0, acc_wait = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const long long w0 = trace ? clock64() : 0;
          if (FB_FULL_WAIT_CLUSTER) mbar_wait_cluster(&full_bar[stage], phase); else mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (trace) full_wait += clock64() - w0;
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const bool nd = P.big_ndir != 0;
          for (int s = 0; s < it.nsub; ++s) {
            const int acc = acc0 + s;
            if (kb == 0) {  // the epilogue (of both CTAs) must have drained this accumulator
              const long long a0 = trace ? clock64() : 0;
              mbar_wait_cluster(&tmem_empty[acc], ((acc_phase >> acc) & 1u) ^ 1u);
              tc_fence_after();
              if (trace) acc_wait += clock64() - a0;
            }
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            const uint64_t da = umma_smem_desc_sw128(sa + (nd ? 0 : s) * A_BYTES, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(sa + (nd ? 1 + s : 2) * A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) umma_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            if (kb == p.num_kb - 1) {  // this accumulator is complete once the MMAs issued so far retire
              tc_commit_2sm(&tmem_full[acc], 3);
              acc_phase ^= 1u << acc;
            }
          }
          tc_commit_2sm(&empty_bar[stage], 3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (trace && item_no < 64) {
          trace[item_no * 4 + 0] = c0;                      // item start
          trace[item_no * 4 + 1] = acc_wait;                // waited for the epilogue to free the accumulator(s)
          trace[item_no * 4 + 2] = full_wait;               // waited for TMA data, summed over the k-blocks
          trace[item_no * 4 + 3] = (clock64() - c0) * 4 + it.nsub;  // item total (issue side) x4 + sub-tile count
        }
        ++item_no;
      }
    }
  } else {
    // ================= epilogue warps =================
    uint32_t ephase = 0;  // bit a: phase of accumulator a
    int small_count = 0;
    for (int w = unit_id; w < P.total_items; w += num_units) {
      const BigItem it = decode_big_item(P, w);
      if (it.nsub == 0) continue;
      const GemmProblemDev& p = P.p[it.prob];
      int acc0 = 0;
      if (it.nsub == 2) small_count = 0; else acc0 = small_count++ & 1;
      for (int s = 0; s < it.nsub; ++s) {
        const int acc = acc0 + s;
        const int m_t = 2 * it.blk[s] + static_cast<int>(cta_rank);
        epi_drain<EPI_WARPS, false>(p, it.nt[s], m_t, tmem_base + acc * BLOCK_N, warp, lane, &tmem_full[acc],
                                    (ephase >> acc) & 1u);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        ephase ^= 1u << acc;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static long long* g_gemm_trace = nullptr;
int set_gemm_trace(long long* buf) {
  if (buf != nullptr && !FB_GEMM_TRACE)
    return fail("this libfluxb200.so was built without the GEMM trace (rebuild with FLUXB200_GEMM_TRACE=1)");
  g_gemm_trace = buf;
  return 0;
}

int gemm_init_device() {
  static DeviceOnce attr_once;
  if (attr_once.need()) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<false>::SMEM)));
    FB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<true>::SMEM)));
    FB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<true, true>::SMEM)));
    FB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<true>::SMEM)));
    FB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(BigCfg::SMEM)));
    attr_once.done();
  }
  return 0;
}

int launch_gemm(const GemmDesc* descs, int count, cudaStream_t stream) {
  FB_REQUIRE(count >= 1 && count <= MAX_PROBLEMS, "launch_gemm: 1..4 problems per launch");
  static_assert(sizeof(GemmParams) < 32000, "kernel parameter block too large");
  if (int rc = gemm_init_device()) return rc;
  bool use_pair = true;
  use_pair = get_flag("gemm_pair") != 0;  // A/B switch (env FLUXB200_GEMM_SINGLE_CTA=1 or fluxb200_set_flag)
  bool quant_b = false;
  for (int i = 0; i < count; ++i) quant_b |= (descs[i].qb != nullptr);
  if (quant_b) {
    for (int i = 0; i < count; ++i)
      FB_REQUIRE(descs[i].qb != nullptr && !descs[i].conv, "launch_gemm: quantised and dense problems cannot share a launch");
    use_pair = true;  // the fused-dequant producer exists for the CTA-pair kernel only
  }
  // CL4 (pairs of pairs with W multicast): dense, non-conv problems whose M fills 512-row units without much padding
  bool use_cl4 = use_pair && !quant_b && get_flag("gemm_cl4") != 0;
  for (int i = 0; i < count && use_cl4; ++i) {
    const int tiles_m = (descs[i].M + BLOCK_M - 1) / BLOCK_M;
    const int padded = (tiles_m + 3) / 4 * 4;
    if (descs[i].conv || (padded - tiles_m) * 32 > tiles_m) use_cl4 = false;
  }
  static int max_clusters4 = -1;  // co-resident clusters of 4 CTAs of this kernel (33 on a 148-SM B200)
  if (use_cl4 && max_clusters4 < 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(4 * (num_sms() / 4)), cfg.blockDim = dim3(GEMM_THREADS), cfg.dynamicSmemBytes = GemmCfg<true>::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<true, false, true>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 0;
    }
    max_clusters4 = n;
  }
  if (use_cl4 && max_clusters4 < 1) use_cl4 = false;
  // BIG tiles (512x256 per pair, see gemm_tcgen05_big_kernel): dense long-K problems with at least one full wave
  // ("gemm_big" = 1: K >= 8192; 2: every eligible GEMM, for experiments)
  const int big_flag = get_flag("gemm_big");  // 1/2: sub-tiles stacked along M (share W); 3/4: side by side along N (share A)
  const bool big_ndir = big_flag >= 3;
  const bool big_all = big_flag == 2 || big_flag == 4;
  bool use_big = use_pair && !quant_b && !use_cl4 && big_flag > 0;
  {
    long long positions = 0;
    for (int i = 0; i < count && use_big; ++i) {
      const GemmDesc& d = descs[i];
      if (d.conv || d.qkrope || d.n_split || d.K % BLOCK_K != 0 || (!big_all && d.K < 8192)) use_big = false;
      const int blocks = (d.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), ntiles = (d.N + BLOCK_N - 1) / BLOCK_N;
      positions += big_ndir ? static_cast<long long>(blocks) * ((ntiles + 1) / 2)
                            : static_cast<long long>((blocks + 1) / 2) * ntiles;
    }
    if (positions < num_sms() / 2) use_big = false;
  }
  const int b_box_rows = use_pair ? GemmCfg<true>::B_ROWS : GemmCfg<false>::B_ROWS;
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.count = count;
  int tile = 0;
  for (int i = 0; i < count; ++i) {
    const GemmDesc& d = descs[i];
    GemmProblemDev& p = P.p[i];
    FB_REQUIRE(d.a && (d.w || d.qb) && (d.out0 || d.qkrope), "launch_gemm: null operand");
    FB_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "launch_gemm: empty problem");
    FB_REQUIRE((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.w) & 15) == 0,
               "launch_gemm: operands must be 16-byte aligned (TMA)");
    p.qcount = 0;
    if (d.qb) {
      FB_REQUIRE(d.qb->count >= 1 && d.qb->count <= 4, "launch_gemm: 1..4 quantised members");
      FB_REQUIRE(d.K % 64 == 0, "launch_gemm: quantised B needs K % 64 == 0");
      p.qcount = d.qb->count;
      for (int qi = 0; qi < d.qb->count; ++qi) {
        const QuantMember& qmem = d.qb->m[qi];
        FB_REQUIRE(qmem.packed && (reinterpret_cast<uintptr_t>(qmem.packed) & 15) == 0, "launch_gemm: packed weights must be 16-byte aligned");
        FB_REQUIRE(qmem.row_begin % 128 == 0, "launch_gemm: quantised members must start on a multiple of 128 rows");
        if (qmem.kind == QB_NF4 || qmem.kind == QB_FP4)
          FB_REQUIRE(qmem.absmax && qmem.blocksize % 64 == 0, "launch_gemm: 4-bit members need absmax and blocksize % 64 == 0");
        if (qmem.kind == QB_Q4K) FB_REQUIRE(d.K % 256 == 0, "launch_gemm: Q4_K needs K % 256 == 0");
        if (qmem.kind == QB_INT8) FB_REQUIRE(qmem.scb != nullptr, "launch_gemm: int8 members need SCB");
        p.qm[qi] = qmem;
        // tensor maps over this member's packed bytes (+ aux): rows = the member's output rows
        const uint64_t rows = static_cast<uint64_t>((qi + 1 < d.qb->count ? d.qb->m[qi + 1].row_begin : d.N) - qmem.row_begin);
        int rc = 0;
        if (qmem.kind == QB_INT8) {
          rc = encode_tmap_2d_raw(&p.tmap_pk[qi], qmem.packed, 1, d.K, rows, d.K, 64, 128);
        } else if (qmem.kind == QB_Q4K) {
          const uint64_t rb = static_cast<uint64_t>(d.K / 256) * 144;
          rc = encode_tmap_2d_raw(&p.tmap_pk[qi], qmem.packed, 1, rb, rows, rb, 32, 128);
          if (!rc) rc = encode_tmap_2d_raw(&p.tmap_aux[qi], qmem.packed, 1, rb, rows, rb, 16, 128);
        } else {
          const uint64_t nabs = static_cast<uint64_t>(d.K / qmem.blocksize);
          FB_REQUIRE(d.K % qmem.blocksize == 0 && (nabs * 4) % 16 == 0 && nabs >= 4,
                     "launch_gemm: fused 4-bit dequant needs (K / blocksize) % 4 == 0 (TMA row pitch of absmax)");
          FB_REQUIRE((reinterpret_cast<uintptr_t>(qmem.absmax) & 15) == 0, "launch_gemm: absmax must be 16-byte aligned");
          rc = encode_tmap_2d_raw(&p.tmap_pk[qi], qmem.packed, 1, d.K / 2, rows, d.K / 2, 32, 128);
          // absmax is mapped as raw bytes (an f32-typed map with a 4-element box faults on sm_100): 16-byte boxes
          if (!rc) rc = encode_tmap_2d_raw(&p.tmap_aux[qi], qmem.absmax, 1, nabs * 4, rows, nabs * 4, 16, 128);
        }
        if (rc) return rc;
      }
    }
    FB_REQUIRE((reinterpret_cast<uintptr_t>(d.out0) & 15) == 0 || (d.N % 8) != 0,
               "launch_gemm: output must be 16-byte aligned");
    FB_REQUIRE(d.n_split % BLOCK_N == 0, "launch_gemm: n_split must be a multiple of 256");
    if (d.n_split > 0) FB_REQUIRE(d.out1 != nullptr, "launch_gemm: n_split without out1");
    if (d.bias_mode != BIAS_NONE) FB_REQUIRE(d.bias != nullptr, "launch_gemm: bias_mode without bias");
    p.M = d.M, p.N = d.N, p.K = d.K;
    p.conv = d.conv;
    if (d.conv) {
      FB_REQUIRE(d.ksize == 1 || d.ksize == 3, "launch_gemm: conv ksize must be 1 or 3");
      FB_REQUIRE(d.cC % 8 == 0, "launch_gemm: conv channels must be a multiple of 8");
      FB_REQUIRE(d.K == d.ksize * d.ksize * d.cC, "launch_gemm: conv K != taps*C");
      FB_REQUIRE(static_cast<long long>(d.cN) * d.cH * d.cW == d.M, "launch_gemm: conv M != N*H*W");
      p.cH = d.cH, p.cW = d.cW, p.cC = d.cC, p.ksize = d.ksize;
      p.c_chunks = (d.cC + BLOCK_K - 1) / BLOCK_K;
      p.tiles_h = (d.cH + CONV_TH - 1) / CONV_TH;
      p.tiles_w = (d.cW + CONV_TW - 1) / CONV_TW;
      p.tiles_m = d.cN * p.tiles_h * p.tiles_w;
      p.num_kb = d.ksize * d.ksize * p.c_chunks;
      int rc = encode_tmap_4d(&p.tmap_a, d.a, d.cC, d.cW, d.cH, d.cN, static_cast<uint64_t>(d.cC) * 2,
                              static_cast<uint64_t>(d.cW) * d.cC * 2, static_cast<uint64_t>(d.cH) * d.cW * d.cC * 2,
                              BLOCK_K, CONV_TW, CONV_TH, 1);
      if (rc) return rc;
    } else {
      FB_REQUIRE(d.K % 8 == 0 && d.lda % 8 == 0, "launch_gemm: K and lda must be multiples of 8");
      p.tiles_m = (d.M + BLOCK_M - 1) / BLOCK_M;
      p.num_kb = (d.K + BLOCK_K - 1) / BLOCK_K;
      int rc = encode_tmap_2d(&p.tmap_a, d.a, d.K, d.M, static_cast<uint64_t>(d.lda) * 2, BLOCK_K, BLOCK_M);
      if (rc) return rc;
    }
    if (!d.qb) {
      FB_REQUIRE(d.ldb % 8 == 0, "launch_gemm: ldb must be a multiple of 8");
      int rc = encode_tmap_2d(&p.tmap_b, d.w, d.K, d.N, static_cast<uint64_t>(d.ldb) * 2, BLOCK_K, b_box_rows);
      if (rc) return rc;
      if (use_cl4) {
        rc = encode_tmap_2d(&p.tmap_b4, d.w, d.K, d.N, static_cast<uint64_t>(d.ldb) * 2, BLOCK_K, GemmCfg<true>::B_ROWS / 2);
        if (rc) return rc;
      }
    }
    p.tiles_n = (d.N + BLOCK_N - 1) / BLOCK_N;
    p.real_tiles_n = p.tiles_n;
    p.sched_m = (use_cl4 || (use_big && !big_ndir)) ? (p.tiles_m + 3) / 4 : (use_pair ? (p.tiles_m + 1) / 2 : p.tiles_m);
    if (use_big && big_ndir) p.tiles_n = (p.tiles_n + 1) / 2;  // positions are 256 rows x 512 columns
    // L2-aware rasterisation: when the whole A operand fits comfortably in the 126 MB L2 (activations of one DiT
    // block: 28 MB), walk all of M for a few N tiles at a time so that every weight panel is fetched from HBM once
    // (ncu: 650 MB -> ~algorithmic 360 MB of DRAM traffic for the 4608x21504x3072 launch); otherwise groups of 8.
    {
      const double a_bytes = 2.0 * static_cast<double>(d.M) * d.K;
      p.group_m = (a_bytes <= 48e6) ? p.sched_m : GROUP_M;
      if (p.group_m < 1) p.group_m = 1;
    }
    p.tile_begin = tile;
    tile += p.sched_m * p.tiles_n;
    p.tile_end = tile;
    p.out0 = d.out0, p.ld0 = d.ld0, p.out1 = d.out1, p.ld1 = d.ld1;
    p.n_split = d.n_split, p.col_off1 = d.col_off1;
    {
      const bool has_alpha = d.alpha != 1.0f;
      FB_REQUIRE(!(d.gate && !d.res), "launch_gemm: gate without residual is not supported");
      FB_REQUIRE(!(d.act0 == ACT_GELU && (d.res || has_alpha)), "launch_gemm: GELU cannot be combined with residual/alpha");
      FB_REQUIRE(!(has_alpha && d.res), "launch_gemm: alpha cannot be combined with a residual");
      p.ev0 = d.act0 == ACT_GELU ? EV_GELU : (d.gate ? EV_GATE_RES : (d.res ? EV_RES : (has_alpha ? EV_ALPHA : EV_PLAIN)));
      if (d.qkrope) {
        FB_REQUIRE(p.ev0 == EV_PLAIN, "launch_gemm: qkrope cannot be combined with another epilogue on segment 0");
        const int nq = 3 * d.qk_H * 128;
        FB_REQUIRE(d.qk_H > 0 && (d.n_split == nq || (d.n_split == 0 && d.N == nq)),
                   "launch_gemm: qkrope needs the q|k|v projection (3*H*128 columns) as segment 0");
        FB_REQUIRE(d.qk_wq && d.qk_wk && d.qk_pe2 && d.qk_Q && d.qk_K && d.qk_V && d.rows_per_batch > 0,
                   "launch_gemm: qkrope operands missing");
        p.ev0 = EV_QKROPE;
        p.qk_wq = d.qk_wq, p.qk_wk = d.qk_wk, p.qk_pe2 = d.qk_pe2, p.qk_pe_bstride = d.qk_pe_bstride;
        p.qk_Q = d.qk_Q, p.qk_K = d.qk_K, p.qk_V = d.qk_V;
        p.qk_H = d.qk_H, p.qk_L = d.qk_L, p.qk_loff = d.qk_loff, p.qk_eps = d.qk_eps;
      }
      p.ev1 = d.act1 == ACT_GELU ? EV_GELU : EV_PLAIN;
      if (p.ev0 == EV_GATE_RES || p.ev0 == EV_RES) {
        FB_REQUIRE((reinterpret_cast<uintptr_t>(d.res) & 15) == 0 && d.ld0 % 8 == 0, "launch_gemm: residual alignment");
        if (d.gate)
          FB_REQUIRE((reinterpret_cast<uintptr_t>(d.gate) & 15) == 0 && d.gate_bstride % 8 == 0, "launch_gemm: gate alignment");
      }
      if (d.bias_mode != BIAS_NONE)
        FB_REQUIRE((reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "launch_gemm: bias must be 16-byte aligned");
    }
    p.bias = d.bias, p.bias_mode = d.bias_mode;
    p.gate = d.gate, p.gate_bstride = d.gate_bstride, p.rows_per_batch = d.rows_per_batch, p.res = d.res;
    p.step_ptr = d.step_ptr, p.gate_step_stride = d.gate_step_stride;
    if (d.step_ptr) FB_REQUIRE(d.gate_step_stride % 8 == 0, "launch_gemm: gate step stride alignment");
    p.alpha = d.alpha, p.has_alpha = (d.alpha != 1.0f);
    if (d.N % 8 == 0) FB_REQUIRE(d.ld0 % 8 == 0, "launch_gemm: ldo must be a multiple of 8");
  }
  P.total_tiles = tile;
  P.trace = g_gemm_trace;
  P.big_ndir = big_ndir ? 1 : 0;
  if (use_big) {
    const int units = num_sms() / 2;
    P.n_big = (tile / units) * units;  // full waves of 512x256 tiles; the rest as 256x256 halves
    P.total_items = P.n_big + 2 * (tile - P.n_big);
  }
  double flops = 0, bytes = 0;
  for (int i = 0; i < count; ++i) {
    flops += 2.0 * descs[i].M * static_cast<double>(descs[i].N) * descs[i].K;
    bytes += 2.0 * (static_cast<double>(descs[i].M) * descs[i].K + static_cast<double>(descs[i].N) * descs[i].K +
                    static_cast<double>(descs[i].M) * descs[i].N);
  }
  ProfScope _ps(KK_GEMM, flops, bytes, stream);
  count_launch(KK_GEMM);
  const bool pdl = get_flag("pdl") != 0;
  if (use_big) {
    FB_CHECK_CUDA(launch_ex(gemm_tcgen05_big_kernel, dim3(2 * std::min(P.total_items, num_sms() / 2)), dim3(GEMM_THREADS),
                            BigCfg::SMEM, stream, 2, pdl, P));
  } else if (use_cl4) {
    FB_CHECK_CUDA(launch_ex(gemm_tcgen05_kernel<true, false, true>, dim3(4 * std::min(tile, max_clusters4)),
                            dim3(GEMM_THREADS), GemmCfg<true>::SMEM, stream, 4, pdl, P));
  } else if (use_pair) {
    static int max_units = -1;  // experiment knob: FLUXB200_GEMM_MAX_UNITS=37 runs the pair kernel on half of the SMs
    if (max_units < 0) {
      const char* e = getenv("FLUXB200_GEMM_MAX_UNITS");
      max_units = e ? std::max(1, atoi(e)) : num_sms() / 2;
    }
    const dim3 grid(2 * std::min(tile, std::min(max_units, num_sms() / 2)));
    if (quant_b)
      FB_CHECK_CUDA(launch_ex(gemm_tcgen05_kernel<true, true>, grid, dim3(GEMM_THREADS), GemmCfg<true, true>::SMEM, stream,
                              2, pdl, P));
    else
      FB_CHECK_CUDA(launch_ex(gemm_tcgen05_kernel<true, false>, grid, dim3(GEMM_THREADS), GemmCfg<true, false>::SMEM,
                              stream, 2, pdl, P));
  } else {
    FB_CHECK_CUDA(launch_ex(gemm_tcgen05_kernel<false, false>, dim3(std::min(tile, num_sms())), dim3(GEMM_THREADS),
                            GemmCfg<false>::SMEM, stream, 1, pdl, P));
  }
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
