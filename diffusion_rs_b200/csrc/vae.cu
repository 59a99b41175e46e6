// fluxb200 — VAE decode kernels (NHWC internally).  The 3x3 / 1x1 convolutions and the mid-block attention GEMMs run
// on the tcgen05 GEMM (gemm.cu, implicit-GEMM conv mode: TMA boxes over the NHWC image, no im2col buffer); this file
// holds the HBM-bound pieces: GroupNorm(+SiLU), nearest 2x upsample, bf16 row softmax, layout changes, pre/post
// processing.  Reference: diffusion_rs_core/src/models/vaes/vae.rs, diffusion_rs_common/src/nn/group_norm.rs:39-74.
#include "internal.h"
#include "kernels.h"
#include "ptx.cuh"

namespace fb {

__device__ __forceinline__ void unpack8v(const uint4& u, float* f) {
  f[0] = bf_lo(u.x), f[1] = bf_hi(u.x), f[2] = bf_lo(u.y), f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z), f[5] = bf_hi(u.z), f[6] = bf_lo(u.w), f[7] = bf_hi(u.w);
}

__device__ __forceinline__ float silu_steps(float v) {  // core/op.rs:703-705, bf16 op by op
  const float e = rbf(expf(-v));
  const float d = rbf(1.0f + e);
  return rbf(v / d);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(32 groups) + optional SiLU on NHWC: ONE statistics pass + one apply pass (the reference runs two
// statistics passes in f32: mean, then the centred sum of squares; nn/group_norm.rs:39-74).
// The single pass accumulates SHIFTED moments  S1 = sum(x - c), S2 = sum((x - c)^2)  with the pilot c = the group's
// first value of the image (so the sums are of O(sigma) terms and  var = (S2 - S1^2/n)/n  has no cancellation
// problem), f32 per thread (<= 256 terms), f64 across threads / blocks (the result does not depend on the unordered
// atomic arrival order beyond f64 rounding).  mean and sqrt(var + eps) then agree with the reference's two-pass f32
// values to ~1e-7 relative, far below the bf16 rounding of the normalised value.
// x: [N, HW, C] bf16; stats: double [N, groups, 2] = {S1, S2}.  Thread t owns the 8-channel vector t % (C/8).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gn_pilot(const bf16* __restrict__ x, int n, int HW, int C, int cpg, int g) {
  return __bfloat162float(x[static_cast<long long>(n) * HW * C + g * cpg]);
}

__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ stats, int HW,
                                                       int C, int groups, int pix_per_block) {
  __shared__ double part[64];  // [group][S1, S2]
  const int n = blockIdx.y;
  const int cv = C / 8;                 // vectors per pixel
  const int rows = blockDim.x / cv;     // pixels processed per iteration
  const int vec = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int cpg = C / groups;
  if (threadIdx.x < 64) part[threadIdx.x] = 0.0;
  __syncthreads();
  const int g0 = (vec * 8) / cpg;
  const int g1 = (vec * 8 + 4) / cpg;  // differs from g0 only when cpg == 4
  const float c0 = gn_pilot(x, n, HW, C, cpg, g0), c1 = gn_pilot(x, n, HW, C, cpg, g1);
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  if (prow < rows) {
    const bf16* base = x + static_cast<long long>(n) * HW * C + vec * 8;
    int p = p_begin + prow;
    // four independent 16-byte loads in flight per thread
    for (; p + 3 * rows < p_end; p += 4 * rows) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(base + static_cast<long long>(p + k * rows) * C);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8v(u[k], f);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d0 = f[e] - c0, d1 = f[4 + e] - c1;
          s0 += d0, q0 = fmaf(d0, d0, q0);
          s1 += d1, q1 = fmaf(d1, d1, q1);
        }
      }
    }
    for (; p < p_end; p += rows) {
      float f[8];
      unpack8v(*reinterpret_cast<const uint4*>(base + static_cast<long long>(p) * C), f);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = f[e] - c0, d1 = f[4 + e] - c1;
        s0 += d0, q0 = fmaf(d0, d0, q0);
        s1 += d1, q1 = fmaf(d1, d1, q1);
      }
    }
  }
  if (g0 == g1) {
    atomicAdd(&part[g0 * 2], static_cast<double>(s0) + static_cast<double>(s1));
    atomicAdd(&part[g0 * 2 + 1], static_cast<double>(q0) + static_cast<double>(q1));
  } else {
    atomicAdd(&part[g0 * 2], static_cast<double>(s0));
    atomicAdd(&part[g0 * 2 + 1], static_cast<double>(q0));
    atomicAdd(&part[g1 * 2], static_cast<double>(s1));
    atomicAdd(&part[g1 * 2 + 1], static_cast<double>(q1));
  }
  __syncthreads();
  if (threadIdx.x < 2 * groups) {
    const double v = part[threadIdx.x];
    if (v != 0.0) atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + threadIdx.x], v);
  }
}

// y = silu?( bf16( bf16( bf16((x - mean) / sqrt(var + eps)) * w ) + b ) ); grid = (vector blocks per image, N)
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, const double* __restrict__ stats,
                                                       const bf16* __restrict__ w, const bf16* __restrict__ b,
                                                       bf16* __restrict__ y, int HW, int C, int groups, float eps,
                                                       int apply_silu, long long vec_per_image) {
  __shared__ float2 gs[32];  // per group: mean, sqrt(var + eps)
  const int n = blockIdx.y;
  const int cpg = C / groups;
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const double cnt = static_cast<double>(HW) * cpg;
    const double S1 = stats[(static_cast<long long>(n) * groups + g) * 2];
    const double S2 = stats[(static_cast<long long>(n) * groups + g) * 2 + 1];
    const double dm = S1 / cnt;
    const double var = fmax((S2 - S1 * dm) / cnt, 0.0);
    const float mean = static_cast<float>(static_cast<double>(gn_pilot(x, n, HW, C, cpg, g)) + dm);
    gs[g] = make_float2(mean, sqrtf(static_cast<float>(var) + eps));
  }
  __syncthreads();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= vec_per_image) return;
  const int cv = C / 8;
  const int vec = static_cast<int>(i % cv);
  const long long off = (static_cast<long long>(n) * vec_per_image + i) * 8;
  const float2 st0 = gs[(vec * 8) / cpg], st1 = gs[(vec * 8 + 4) / cpg];
  // The kernel was instruction-bound (1.1-1.3 TB/s: two f32 divisions and an expf per element).  Now: the division by
  // sqrt(var + eps) is a reciprocal + one Newton step (correctly rounded quotient, 3 FMAs); * w and + b run as packed
  // bf16x2 ops (an HMUL2/HADD2.BF16 is exactly "f32 op, round to nearest even", the reference's per-op rounding); SiLU
  // keeps its three bf16 roundings with ex2.approx / a Newton-corrected reciprocal in f32.
  const float inv0 = 1.0f / st0.y, inv1 = 1.0f / st1.y;
  const uint4 xu = *reinterpret_cast<const uint4*>(x + off);
  const uint4 wu = *reinterpret_cast<const uint4*>(w + vec * 8);
  const uint4 bu = *reinterpret_cast<const uint4*>(b + vec * 8);
  const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w}, ww[4] = {wu.x, wu.y, wu.z, wu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
  uint32_t ow[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float mean = e < 2 ? st0.x : st1.x, den = e < 2 ? st0.y : st1.y, inv = e < 2 ? inv0 : inv1;
    const float d0 = bf_lo(xw[e]) - mean, d1 = bf_hi(xw[e]) - mean;
    float q0 = d0 * inv, q1 = d1 * inv;
    q0 = fmaf(fmaf(-q0, den, d0), inv, q0);
    q1 = fmaf(fmaf(-q1, den, d1), inv, q1);
    __nv_bfloat162 v = __floats2bfloat162_rn(q0, q1);                                          // normalised -> bf16
    v = __hmul2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&ww[e]));                       // * w -> bf16
    v = __hadd2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&bw[e]));                       // + b -> bf16
    if (apply_silu) {  // v / (1 + exp(-v)), every op rounded to bf16 (core/op.rs:703-705)
      const float2 f = __bfloat1622float2(v);
      const __nv_bfloat162 en = __floats2bfloat162_rn(ex2_approx(-f.x * 1.4426950408889634f),
                                                      ex2_approx(-f.y * 1.4426950408889634f));
      const float2 dd = __bfloat1622float2(__hadd2_rn(en, __float2bfloat162_rn(1.0f)));
      float r0 = __frcp_rn(dd.x), r1 = __frcp_rn(dd.y);
      float s0 = f.x * r0, s1 = f.y * r1;
      // exp(-v) overflows to inf for v < -88.7: v / inf = -0, and the Newton step would turn it into 0 * inf = NaN
      s0 = dd.x > 3.0e38f ? f.x * 0.0f : fmaf(fmaf(-s0, dd.x, f.x), r0, s0);
      s1 = dd.y > 3.0e38f ? f.y * 0.0f : fmaf(fmaf(-s1, dd.y, f.y), r1, s1);
      v = __floats2bfloat162_rn(s0, s1);
    }
    ow[e] = *reinterpret_cast<const uint32_t*>(&v);
  }
  *reinterpret_cast<uint4*>(y + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

int launch_groupnorm_silu(const bf16* x, const bf16* w, const bf16* b, bf16* y, int N, int HW, int C, int groups,
                          float eps, int apply_silu, double* stats, cudaStream_t stream) {
  FB_REQUIRE(groups == 32 && C % (8 * 1) == 0 && (C / groups) >= 4 && 256 % (C / 8) == 0,
             "groupnorm: needs 32 groups, >= 4 channels per group and C/8 dividing 256");
  FB_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * N * groups * 2, stream));
  const int pix_per_block = 1024;
  ProfScope _ps(KK_GROUPNORM, 0, 6.0 * N * HW * C, stream);  // algorithmic: x read twice (stats, apply), y written once
  count_launch(KK_GROUPNORM, 2);
  dim3 grid((HW + pix_per_block - 1) / pix_per_block, N);
  gn_stats_kernel<<<grid, 256, 0, stream>>>(x, stats, HW, C, groups, pix_per_block);
  const long long vec_per_image = static_cast<long long>(HW) * (C / 8);
  gn_apply_kernel<<<dim3(static_cast<unsigned>((vec_per_image + 255) / 256), N), 256, 0, stream>>>(
      x, stats, w, b, y, HW, C, groups, eps, apply_silu, vec_per_image);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// nearest 2x upsample, NHWC (Upsample::forward vae.rs:224-228; upsample_nearest2d conv.cu:501-540)
__global__ void upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int H, int W, int C,
                                  long long total_vec) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total_vec) return;
  const int cv = C / 8;
  const int v = static_cast<int>(i % cv);
  long long p = i / cv;
  const int ox = static_cast<int>(p % (2 * W));
  p /= (2 * W);
  const int oy = static_cast<int>(p % (2 * H));
  const long long n = p / (2 * H);
  const uint4 u = *reinterpret_cast<const uint4*>(x + ((n * H + oy / 2) * W + ox / 2) * C + v * 8);
  *reinterpret_cast<uint4*>(y + i * 8) = u;
}
int launch_upsample2x_nhwc(const bf16* x, bf16* y, int N, int H, int W, int C, cudaStream_t stream) {
  const long long total_vec = static_cast<long long>(N) * 4 * H * W * (C / 8);
  count_launch(KK_MISC);
  upsample2x_kernel<<<static_cast<unsigned>((total_vec + 255) / 256), 256, 0, stream>>>(x, y, H, W, C, total_vec);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Row softmax on bf16 scores, in place (vae.rs:28-33 -> softmax_last_dim on the model dtype; CUDA softmax_bf16
// reduce.cu:180-219, 576: max-subtract and exp in bf16, row sum accumulated in f32).
__global__ void __launch_bounds__(256) softmax_rows_kernel(bf16* __restrict__ x, int cols) {
  __shared__ float red[8];
  bf16* row = x + static_cast<long long>(blockIdx.x) * cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x * 8; c < cols; c += 256 * 8) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, f[e]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x * 8; c < cols; c += 256 * 8) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f[e] = rbf(expf(rbf(f[e] - m)));
      s += f[e];
    }
    uint4 u;
    u.x = pack_bf16(f[0], f[1]), u.y = pack_bf16(f[2], f[3]), u.z = pack_bf16(f[4], f[5]), u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(row + c) = u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  s = rbf(s);
  for (int c = threadIdx.x * 8; c < cols; c += 256 * 8) {
    float f[8];
    unpack8v(*reinterpret_cast<const uint4*>(row + c), f);
    uint4 u;
    u.x = pack_bf16(f[0] / s, f[1] / s), u.y = pack_bf16(f[2] / s, f[3] / s);
    u.z = pack_bf16(f[4] / s, f[5] / s), u.w = pack_bf16(f[6] / s, f[7] / s);
    *reinterpret_cast<uint4*>(row + c) = u;
  }
}
int launch_softmax_rows_bf16(bf16* x, long long rows, int cols, cudaStream_t stream) {
  FB_REQUIRE(cols % 8 == 0, "softmax_rows: cols must be a multiple of 8");
  count_launch(KK_MISC);
  softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(x, cols);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// [rows, cols] -> [cols, rows] through a 32x32 smem tile
__global__ void transpose_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int rows, int cols) {
  __shared__ bf16 tile[32][33];
  const long long zoff = static_cast<long long>(blockIdx.z) * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = x[zoff + static_cast<long long>(r) * cols + c];
  }
  __syncthreads();
  int r = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int cc = blockIdx.x * 32 + j;
    if (r < rows && cc < cols) y[zoff + static_cast<long long>(cc) * rows + r] = tile[threadIdx.x][j];
  }
}
static int launch_transpose_batched(const bf16* x, bf16* y, int batch, int rows, int cols, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  count_launch(KK_MISC);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, y, rows, cols);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_transpose_2d(const bf16* x, bf16* y, int rows, int cols, cudaStream_t stream) {
  return launch_transpose_batched(x, y, 1, rows, cols, stream);
}
// NCHW [N, C, HW] <-> NHWC [N, HW, C] are batched 2-D transposes
int launch_nchw_to_nhwc(const bf16* x, bf16* y, int N, int C, int H, int W, cudaStream_t stream) {
  return launch_transpose_batched(x, y, N, C, H * W, stream);
}
int launch_nhwc_to_nchw(const bf16* x, bf16* y, int N, int C, int H, int W, cudaStream_t stream) {
  return launch_transpose_batched(x, y, N, H * W, C, stream);
}

// Conv weight repack [Cout, Cin, k, k] -> [Cout, k, k, Cin] so K = tap*Cin + c is contiguous (done once at load)
__global__ void repack_conv_weight_kernel(const bf16* __restrict__ w, bf16* __restrict__ out, int Cin, int taps,
                                          long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % Cin);
  const int t = static_cast<int>((i / Cin) % taps);
  const long long o = i / (static_cast<long long>(Cin) * taps);
  out[i] = w[(o * Cin + c) * taps + t];
}
int launch_repack_conv_weight(const bf16* w, bf16* out, int Cout, int Cin, int taps, cudaStream_t stream) {
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  count_launch(KK_MISC);
  repack_conv_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(w, out, Cin, taps, total);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// unpack (flux/sampling.rs:61-68) + `z / scaling_factor + shift_factor` (flux/mod.rs:329; two bf16 affine ops) -> NHWC
//   packed [N, h2*w2, 64] with feature index c*4 + ph*2 + pw  ->  z [N, 2*h2, 2*w2, 16]
__global__ void unpack_latents_kernel(const bf16* __restrict__ packed, bf16* __restrict__ nhwc, int h2, int w2,
                                      float inv_scale_b, float shift_b, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % 16);
  long long p = i / 16;
  const int x = static_cast<int>(p % (2 * w2));
  p /= (2 * w2);
  const int y = static_cast<int>(p % (2 * h2));
  const long long n = p / (2 * h2);
  const float v = __bfloat162float(packed[(n * h2 * w2 + (y / 2) * w2 + (x / 2)) * 64 + c * 4 + (y & 1) * 2 + (x & 1)]);
  // affine(1/s, 0) then affine(1, shift): x*mul -> bf16, +add -> bf16 each
  const float a = rbf(rbf(v * inv_scale_b) + 0.0f);
  nhwc[i] = __float2bfloat16_rn(rbf(a * 1.0f) + shift_b);
}
int launch_unpack_latents(const bf16* packed, bf16* nhwc, int N, int h2, int w2, float inv_scale, float shift,
                          cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * 4 * h2 * w2 * 16;
  const float isb = __bfloat162float(__float2bfloat16_rn(inv_scale));
  const float shb = __bfloat162float(__float2bfloat16_rn(shift));
  count_launch(KK_MISC);
  unpack_latents_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(packed, nhwc, h2, w2, isb, shb,
                                                                                        total);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// post-process (flux/mod.rs:332): clamp(-1,1) -> +1 -> *127.5 (bf16 ops) -> u8 (truncating, saturating cast).
// in: NHWC bf16 [N,H,W,C] ; out: u8 in NHWC (image layout) or NCHW (the reference's tensor layout)
__global__ void postprocess_u8_kernel(const bf16* __restrict__ x, uint8_t* __restrict__ out, int C, long long HW,
                                      int to_nchw, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  float v = __bfloat162float(x[i]);
  v = fminf(fmaxf(v, -1.0f), 1.0f);
  v = rbf(rbf(v * 1.0f) + 1.0f);
  v = rbf(rbf(v * 127.5f) + 0.0f);
  const uint8_t q = static_cast<uint8_t>(fminf(fmaxf(truncf(v), 0.f), 255.f));
  if (!to_nchw) {
    out[i] = q;
  } else {
    const int c = static_cast<int>(i % C);
    const long long p = (i / C) % HW;
    const long long n = i / (C * HW);
    out[(n * C + c) * HW + p] = q;
  }
}
int launch_postprocess_u8(const bf16* nhwc, uint8_t* out, int N, int C, int H, int W, int to_nchw,
                          cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * H * W * C;
  count_launch(KK_MISC);
  postprocess_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      nhwc, out, C, static_cast<long long>(H) * W, to_nchw, total);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
