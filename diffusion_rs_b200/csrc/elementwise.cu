// fluxb200 — HBM-bound warp-reduction / elementwise kernels of the DiT step.
// Each kernel fuses what the reference runs as a chain of separate bf16 tensor ops and keeps the
// reference's rounding points (f32 op -> round-to-nearest-even bf16 after every tensor op).
#include <algorithm>

#include "internal.h"
#include "kernels.h"
#include "ptx.cuh"

namespace fb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf_lo(u.x), f[1] = bf_hi(u.x), f[2] = bf_lo(u.y), f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z), f[5] = bf_hi(u.z), f[6] = bf_lo(u.w), f[7] = bf_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]), u.y = pack_bf16(f[2], f[3]), u.z = pack_bf16(f[4], f[5]), u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm(eps, no affine) + AdaLN modulate:  out = ((LN(x) -> bf16) * (1 + scale -> bf16) -> bf16) + shift -> bf16
// reference: nn::LayerNorm fast path (nn/ops.rs:1021-1043 / reduce.cu:73-131) then ModulationOut::scale_shift
// (models/flux/model.rs:217-221). One warp per row, the row stays in registers between the two passes.
// ------------------------------------------------------------------------------------------------
// Up to two row segments per launch (the img and the txt stream of a double block share it): segment s covers rows
// [row_begin_s, row_begin_{s+1}) of the launch, reads x_s and its own shift/scale vectors; the output rows of the launch
// are contiguous.
struct LnSegment {
  const bf16* x;
  const bf16* shift;
  const bf16* scale;
  long long in_bstride_rows;
  int in_row_off, rows_per_batch, row_begin, pad_;
};
struct LnParams {
  LnSegment seg[2];
  int nseg, total_rows;
  long long mod_bstride;
  bf16* out;
  float eps;
  const int* step_ptr;
  long long step_stride;
};

// One warp per row, the row stays in registers between the two passes; a warp walks rows with the stride of the whole
// grid, which is sized to exactly fill the machine (4 blocks of 4 warps per SM): 4608 rows over 2368 warps is two
// balanced rounds, where one row per warp left a 15 %-full second wave.
// REREAD: the row is NOT kept in registers between the two passes but read again (it sits in L1: 32 warps x 6 KB per SM),
// which halves the register footprint and doubles the warps in flight per SM (8 blocks instead of 4).
template <int D, bool REREAD>
__global__ void __launch_bounds__(128, REREAD ? 8 : 4) ln_modulate_kernel(const __grid_constant__ LnParams P) {
  constexpr int VEC = D / 256;  // uint4 (8 bf16) per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();  // x is the previous kernel's output
  // denoising loop: the modulation vectors of ALL steps were projected before the loop ([step][batch][...]); the
  // current step comes from a device counter so that one captured CUDA graph serves every step
  const long long step_off = P.step_ptr ? *P.step_ptr * P.step_stride : 0;
  const __nv_bfloat162 one2 = __float2bfloat162_rn(1.0f);
  for (int row = blockIdx.x * 4 + warp; row < P.total_rows; row += gridDim.x * 4) {
    const LnSegment& sg = P.seg[(P.nseg > 1 && row >= P.seg[1].row_begin) ? 1 : 0];
    const int lr = row - sg.row_begin;
    const int b = lr / sg.rows_per_batch;
    const int i = lr - b * sg.rows_per_batch;
    const bf16* xr = sg.x + (static_cast<long long>(b) * sg.in_bstride_rows + sg.in_row_off + i) * D;
    uint4 u[REREAD ? 1 : VEC];  // !REREAD: the row stays packed in registers (48 regs)
    if (!REREAD) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) u[k] = *reinterpret_cast<const uint4*>(xr + (k * 32 + lane) * 8);
    }
    // The kernel was instruction-bound (~15 instructions per element), not memory-bound: both passes run on packed
    // pairs - f32x2 add/fma for the statistics and the normalisation, then the modulate chain in bf16x2 (an
    // HMUL2/HADD2.BF16 is exactly "f32 op, round to nearest even", the reference's per-op rounding).
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const uint4 uk = REREAD ? *reinterpret_cast<const uint4*>(xr + (k * 32 + lane) * 8) : u[REREAD ? 0 : k];
      const uint32_t w[4] = {uk.x, uk.y, uk.z, uk.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = bf_lo(w[e]), hi = bf_hi(w[e]);
        asm("{\n\t.reg .b64 a, v;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 v, {%2, %3};\n\tadd.rn.f32x2 a, a, v;\n\t"
            "mov.b64 {%0, %1}, a;\n\t}"
            : "+f"(s0), "+f"(s1)
            : "f"(lo), "f"(hi));
        asm("{\n\t.reg .b64 a, v;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 v, {%2, %3};\n\tfma.rn.f32x2 a, v, v, a;\n\t"
            "mov.b64 {%0, %1}, a;\n\t}"
            : "+f"(q0), "+f"(q1)
            : "f"(lo), "f"(hi));
      }
    }
    const float s = warp_sum(s0 + s1);
    const float s2 = warp_sum(q0 + q1);
    const float mean = s / D;
    const float var = s2 / D - mean * mean;
    const float inv_std = 1.0f / sqrtf(var + P.eps);
    const float nmean = -mean;
    const long long mod_off = static_cast<long long>(b) * P.mod_bstride + step_off;
    const bf16* sh = sg.shift + mod_off;
    const bf16* sc = sg.scale + mod_off;
    bf16* orow = P.out + static_cast<long long>(row) * D;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (k * 32 + lane) * 8;
      const uint4 fs = *reinterpret_cast<const uint4*>(sh + c);
      const uint4 fc = *reinterpret_cast<const uint4*>(sc + c);
      const uint4 uk = REREAD ? *reinterpret_cast<const uint4*>(xr + c) : u[REREAD ? 0 : k];
      const uint32_t w[4] = {uk.x, uk.y, uk.z, uk.w};
      const uint32_t ws[4] = {fs.x, fs.y, fs.z, fs.w};
      const uint32_t wc[4] = {fc.x, fc.y, fc.z, fc.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float y0 = bf_lo(w[e]), y1 = bf_hi(w[e]);
        // (v - mean) * inv_std in f32, two roundings like the reference's fused kernel, on both lanes at once
        asm("{\n\t.reg .b64 v, m, r;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 m, {%2, %2};\n\tmov.b64 r, {%3, %3};\n\t"
            "add.rn.f32x2 v, v, m;\n\tmul.rn.f32x2 v, v, r;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(y0), "+f"(y1)
            : "f"(nmean), "f"(inv_std));
        const __nv_bfloat162 n2 = __floats2bfloat162_rn(y0, y1);                               // LN -> bf16
        const __nv_bfloat162 sc1 = __hadd2_rn(*reinterpret_cast<const __nv_bfloat162*>(&wc[e]), one2);  // scale + 1 -> bf16
        const __nv_bfloat162 m2 = __hmul2_rn(n2, sc1);                                           // * -> bf16
        const __nv_bfloat162 o2 = __hadd2_rn(m2, *reinterpret_cast<const __nv_bfloat162*>(&ws[e]));  // + shift -> bf16
        o[e] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      *reinterpret_cast<uint4*>(orow + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int launch_ln_modulate2(const LnInput* in, int nseg, long long mod_bstride, bf16* out, int D, float eps,
                        cudaStream_t stream, const int* step_ptr, long long step_stride) {
  FB_REQUIRE(D == 3072, "ln_modulate: hidden size must be 3072 (HIDDEN_SIZE, model.rs:17)");
  FB_REQUIRE(nseg == 1 || nseg == 2, "ln_modulate: 1 or 2 segments");
  LnParams P{};
  int rows = 0;
  for (int s = 0; s < nseg; ++s) {
    LnSegment& g = P.seg[s];
    g.x = in[s].x, g.shift = in[s].shift, g.scale = in[s].scale;
    g.in_bstride_rows = in[s].in_bstride_rows, g.in_row_off = in[s].in_row_off, g.rows_per_batch = in[s].rows_per_batch;
    g.row_begin = rows;
    rows += in[s].rows_per_batch * in[s].batch;
  }
  P.nseg = nseg, P.total_rows = rows, P.mod_bstride = mod_bstride, P.out = out, P.eps = eps;
  P.step_ptr = step_ptr, P.step_stride = step_stride;
  ProfScope _ps(KK_LN_MOD, 0, 4.0 * rows * D, stream);
  count_launch(KK_LN_MOD, 1);
  if (get_flag("ln_reread")) {
    const int grid = std::min((rows + 3) / 4, 8 * num_sms());
    FB_CHECK_CUDA(launch_ex(ln_modulate_kernel<3072, true>, dim3(grid), dim3(128), 0, stream, 1, get_flag("pdl") != 0, P));
  } else {
    const int grid = std::min((rows + 3) / 4, 4 * num_sms());
    FB_CHECK_CUDA(launch_ex(ln_modulate_kernel<3072, false>, dim3(grid), dim3(128), 0, stream, 1, get_flag("pdl") != 0, P));
  }
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_ln_modulate(const bf16* x, long long in_bstride_rows, int in_row_off, int rows_per_batch, int batch,
                       const bf16* shift, const bf16* scale, long long mod_bstride, bf16* out, int D, float eps,
                       cudaStream_t stream, const int* step_ptr, long long step_stride) {
  LnInput in{x, shift, scale, in_bstride_rows, in_row_off, rows_per_batch, batch};
  return launch_ln_modulate2(&in, 1, mod_bstride, out, D, eps, stream, step_ptr, step_stride);
}

// ------------------------------------------------------------------------------------------------
// QK RMS-norm + RoPE + head-major relayout.
//   in : qkv rows [rows, ld] with q | k | v at column offsets 0, D, 2D (bf16)
//   out: Q, K, V [B, H, L, 128]; the stream's tokens land at sequence offset l_off (txt first, then img)
// reference: SelfAttention::qkv (model.rs:399-427), RmsNorm slow path (nn/layer_norm.rs:136-153),
//            apply_rope (model.rs:86-95): out0 = cos*x0 + (-sin)*x1, out1 = sin*x0 + cos*x1, every op rounded to bf16.
// One block per token; 16 lanes per head (8 elements = 16 B each); every thread owns the same 8 columns of q, k and v
// of its head, so the three 16-byte loads are in flight together and the RoPE factors are loaded once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) qknorm_rope_kernel(const bf16* __restrict__ qkv, long long ld,
                                                          int rows_per_batch, int H, int L, int l_off,
                                                          const bf16* __restrict__ wq, const bf16* __restrict__ wk,
                                                          const bf16* __restrict__ pe_cos,
                                                          const bf16* __restrict__ pe_sin, long long pe_bstride,
                                                          bf16* __restrict__ Q,
                                                          bf16* __restrict__ K, bf16* __restrict__ V, float eps) {
  const int row = blockIdx.x;
  const int h = threadIdx.x >> 4;   // head
  const int sub = threadIdx.x & 15; // 8-column group inside the head
  if (h >= H) return;
  const int b = row / rows_per_batch;
  const int l = l_off + (row - b * rows_per_batch);
  const int D = H * 128;
  const bf16* base = qkv + static_cast<long long>(row) * ld + h * 128 + sub * 8;
  const uint4 uq = *reinterpret_cast<const uint4*>(base);
  const uint4 uk = *reinterpret_cast<const uint4*>(base + D);
  const uint4 uv = *reinterpret_cast<const uint4*>(base + 2 * D);
  const long long pe_off = b * pe_bstride + static_cast<long long>(l) * 64 + sub * 4;
  const uint2 cu = *reinterpret_cast<const uint2*>(pe_cos + pe_off);
  const uint2 su = *reinterpret_cast<const uint2*>(pe_sin + pe_off);
  const uint4 wqu = *reinterpret_cast<const uint4*>(wq + sub * 8);
  const uint4 wku = *reinterpret_cast<const uint4*>(wk + sub * 8);
  const long long dst_off = ((static_cast<long long>(b) * H + h) * L + l) * 128 + sub * 8;
  *reinterpret_cast<uint4*>(V + dst_off) = uv;
  const float c[4] = {bf_lo(cu.x), bf_hi(cu.x), bf_lo(cu.y), bf_hi(cu.y)};
  const float sn[4] = {bf_lo(su.x), bf_hi(su.x), bf_lo(su.y), bf_hi(su.y)};
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    float x[8], wf[8], o[8];
    unpack8(which == 0 ? uq : uk, x);
    unpack8(which == 0 ? wqu : wku, wf);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += x[e] * x[e];
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);  // 16-lane groups
    const float denom = sqrtf(ss / 128.0f + eps);
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = rbf(rbf(x[e] / denom) * wf[e]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {  // pairs (2i, 2i+1), i = sub*4 + p
      const float x0 = y[2 * p], x1 = y[2 * p + 1];
      o[2 * p] = rbf(rbf(c[p] * x0) + rbf(-sn[p] * x1));
      o[2 * p + 1] = rbf(rbf(sn[p] * x0) + rbf(c[p] * x1));
    }
    *reinterpret_cast<uint4*>((which == 0 ? Q : K) + dst_off) = pack8(o);
  }
}

int launch_qknorm_rope(const bf16* qkv, long long ld, int rows_per_batch, int batch, int H, int L, int l_off,
                       const bf16* wq, const bf16* wk, const bf16* pe_cos, const bf16* pe_sin, long long pe_bstride,
                       bf16* Q, bf16* K, bf16* V, float eps, cudaStream_t stream) {
  const int rows = rows_per_batch * batch;
  ProfScope _ps(KK_QKNORM_ROPE, 0, 12.0 * rows * H * 128, stream);
  count_launch(KK_QKNORM_ROPE, 1);
  FB_REQUIRE(H >= 1 && H <= 32, "qknorm_rope: 1..32 heads");
  qknorm_rope_kernel<<<rows, H * 16, 0, stream>>>(qkv, ld, rows_per_batch, H, L, l_off, wq, wk, pe_cos, pe_sin,
                                               pe_bstride, Q, K, V, eps);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// small kernels of the per-image prologue (vec_ of every step) and of the step graph
// ------------------------------------------------------------------------------------------------
// SiLU in bf16 steps: v / (1 + exp(-v))  (core/op.rs:699-706)
__global__ void silu_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  const float e = rbf(expf(-v));
  const float d = rbf(1.0f + e);
  y[i] = __float2bfloat16_rn(v / d);
}
int launch_silu(const bf16* x, bf16* y, long long n, cudaStream_t stream) {
  count_launch(KK_MISC);
  silu_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, y, n);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// timestep_embedding (model.rs:104-122): t*1000; freqs = exp(k * (-ln(1e4)/half)); [cos | sin] (f32) -> bf16
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int B, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float tt = t[b] * 1000.0f;
  const float coef = static_cast<float>(-9.210340371976184 / static_cast<double>(half));  // -ln(10000)/half
  const float freq = expf(static_cast<float>(k) * coef);
  const float arg = tt * freq;
  out[b * dim + k] = __float2bfloat16_rn(cosf(arg));
  out[b * dim + half + k] = __float2bfloat16_rn(sinf(arg));
}
int launch_timestep_embedding(const float* t, bf16* out, int B, int dim, cudaStream_t stream) {
  const int n = B * dim / 2;
  count_launch(KK_MISC);
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, stream>>>(t, out, B, dim);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// vec_ = (a [+ g]) + y, each add rounded to bf16 (model.rs:813-820)
// a: [rows, D] (one row per (step, batch element)); g, y: [B, D] (they do not depend on the step): row r uses g/y row r % B
__global__ void vec_combine_kernel(const bf16* __restrict__ a, const bf16* __restrict__ g, const bf16* __restrict__ y,
                                   bf16* __restrict__ out, int rows, int B, int Dm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Dm) return;
  const int r = i / Dm, c = i - r * Dm;
  const int j = (r % B) * Dm + c;
  float v = __bfloat162float(a[i]);
  if (g) v = rbf(v + __bfloat162float(g[j]));
  v = rbf(v + __bfloat162float(y[j]));
  out[i] = __float2bfloat16_rn(v);
}
int launch_vec_combine(const bf16* a, const bf16* g, const bf16* y, bf16* out, int rows, int B, int Dm,
                       cudaStream_t stream) {
  count_launch(KK_MISC);
  vec_combine_kernel<<<(rows * Dm + 255) / 256, 256, 0, stream>>>(a, g, y, out, rows, B, Dm);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Per-step scalars of the denoising loop, passed BY VALUE in the kernel parameters (no pageable/pinned staging buffer,
// no synchronisation, capturable): t_all[s*B + b] = t_curr(s), g_all[s*B + b] = guidance, dt_tab[s][c] = bf16(t_prev - t_curr)
// (Sampler::sample, pipelines/sampling.rs:37-44; guidance = Tensor::full, pipelines/flux/mod.rs:300-304).
__global__ void step_scalars_kernel(const StepScalars v, int s0, int n, int B, int C, float guidance,
                                    float* __restrict__ t_all, float* __restrict__ g_all, bf16* __restrict__ dt_tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int s = i / C, c = i - s * C;
  dt_tab[static_cast<long long>(s0 + s) * C + c] = __float2bfloat16_rn(v.dt[s]);
  if (c < B) {
    t_all[(s0 + s) * B + c] = v.t[s];
    g_all[(s0 + s) * B + c] = guidance;
  }
}
int launch_step_scalars(const StepScalars& v, int s0, int n, int B, int C, float guidance, float* t_all, float* g_all,
                        bf16* dt_tab, cudaStream_t stream) {
  FB_REQUIRE(n >= 1 && n <= StepScalars::N && B <= C, "step_scalars: bad chunk");
  count_launch(KK_MISC);
  step_scalars_kernel<<<(n * C + 255) / 256, 256, 0, stream>>>(v, s0, n, B, C, guidance, t_all, g_all, dt_tab);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// end of a denoising step: advance the device-side step counter the step graph's kernels index their per-step data with
__global__ void step_advance_kernel(int* step) { *step += 1; }
int launch_step_advance(int* step, cudaStream_t stream) {
  count_launch(KK_MISC);
  step_advance_kernel<<<1, 1, 0, stream>>>(step);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dst[b][0 .. n16) = src[b][0 .. n16) in 16-byte units with independent batch strides (txt/img -> joint stream "cat",
// latent staging).  A kernel instead of cudaMemcpyAsync so that the step graph consists of kernel nodes only.
__global__ void copy_rows_kernel(uint4* __restrict__ dst, long long dst_bstride16, const uint4* __restrict__ src,
                                 long long src_bstride16, long long n16) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n16) return;
  dst[blockIdx.y * dst_bstride16 + i] = src[blockIdx.y * src_bstride16 + i];
}
int launch_copy_rows(void* dst, long long dst_bstride_bytes, const void* src, long long src_bstride_bytes,
                     long long bytes_per_batch, int batch, cudaStream_t stream) {
  FB_REQUIRE(bytes_per_batch % 16 == 0 && dst_bstride_bytes % 16 == 0 && src_bstride_bytes % 16 == 0 &&
                 (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
             "copy_rows: 16-byte granularity");
  const long long n16 = bytes_per_batch / 16;
  count_launch(KK_MISC);
  copy_rows_kernel<<<dim3(static_cast<unsigned>((n16 + 255) / 256), batch), 256, 0, stream>>>(
      static_cast<uint4*>(dst), dst_bstride_bytes / 16, static_cast<const uint4*>(src), src_bstride_bytes / 16, n16);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Euler update (pipelines/sampling.rs:43): img = img + pred * dt ; dt rounded to bf16 by the affine op, two roundings
__global__ void euler_kernel(bf16* __restrict__ img, const bf16* __restrict__ pred, float dt_bf16, long long n) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float p = rbf(__bfloat162float(pred[i]) * dt_bf16);
  img[i] = __float2bfloat16_rn(__bfloat162float(img[i]) + p);
}
int launch_euler(bf16* img, const bf16* pred, float dt, long long n, cudaStream_t stream) {
  const float dtb = __bfloat162float(__float2bfloat16_rn(dt));
  count_launch(KK_MISC);
  euler_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(img, pred, dtb, n);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out = x * mul + add with bf16 scalars and two roundings (Tensor::affine on bf16, cuda_kernels/affine.cu:33)
__global__ void affine_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, float mul_b, float add_b, long long n) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  y[i] = __float2bfloat16_rn(rbf(__bfloat162float(x[i]) * mul_b) + add_b);
}
int launch_affine(const bf16* x, bf16* y, float mul, float add, long long n, cudaStream_t stream) {
  const float mb = __bfloat162float(__float2bfloat16_rn(mul));
  const float ab = __bfloat162float(__float2bfloat16_rn(add));
  count_launch(KK_MISC);
  affine_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, y, mb, ab, n);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
