// fluxb200 — the two text encoders FLUX conditions on, behind the C ABI.
//
//   T5 encoder : diffusion_rs_core/src/models/t5/mod.rs  (T5EncoderModel::forward :659; once per prompt,
//                pipelines/flux/mod.rs:236-250).  24 x [RMS-norm -> q|k|v -> attention + relative-position bias -> o
//                -> +x -> RMS-norm -> NewGelu(wi_0 x) * wi_1 x -> wo -> +x], final RMS-norm.
//   CLIP text  : diffusion_rs_core/src/models/clip/text.rs (ClipTextTransformer::forward :304-316): token + position
//                embedding, 12 pre-LN layers with causal attention and quick-GELU MLP, final LayerNorm, EOS pooling.
//
// Every Linear is one launch of the tcgen05 GEMM (q|k|v and wi_0|wi_1 fused along N, GELU / bias / residual in the
// epilogue).  The attention here is small (L <= 512, head dim 64, 5e9..1e11 FLOP per prompt against 3.7e15 per image)
// and needs an additive bias / causal mask plus the reference's bf16 rounding points, so it is a plain CUDA-core
// kernel: one warp per query row, keys across lanes, fp32 accumulation.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "fluxb200.h"
#include "internal.h"
#include "kernels.h"
#include "ptx.cuh"

namespace fb {

static constexpr int TE_HEAD_DIM = 64;
static constexpr int TE_MAX_L = 512;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// out[r, :] = table[ids[r], :] (+ pos[r % L, :], one bf16 rounding)
__global__ void te_embed_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ table,
                                const bf16* __restrict__ pos, bf16* __restrict__ out, int rows, int L, int D, int vocab) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  int id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const bf16* src = table + static_cast<long long>(id) * D;
  const bf16* p = pos ? pos + static_cast<long long>(r % L) * D : nullptr;
  bf16* dst = out + static_cast<long long>(r) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = __bfloat162float(src[c]);
    if (p) v = rbf(v + __bfloat162float(p[c]));
    dst[c] = __float2bfloat16_rn(v);
  }
}

// block-wide sum of two values (blockDim.x <= 1024, multiple of 32)
FB_DEVICE void block_sum2(float& a, float& b) {
  __shared__ float sa[32], sb[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sa[w] = a, sb[w] = b;
  __syncthreads();
  a = l < nw ? sa[l] : 0.f;
  b = l < nw ? sb[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

// T5LayerNorm (t5/mod.rs:111-121): y = bf16(x / sqrt(mean(x^2) + eps)); out = bf16(y * w)
__global__ void te_rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out, int D,
                                  float eps) {
  const bf16* xr = x + static_cast<long long>(blockIdx.x) * D;
  bf16* orow = out + static_cast<long long>(blockIdx.x) * D;
  float s2 = 0.f, dummy = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = __bfloat162float(xr[c]);
    s2 += v * v;
  }
  block_sum2(s2, dummy);
  const float denom = sqrtf(s2 / D + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float y = rbf(__bfloat162float(xr[c]) / denom);
    orow[c] = __float2bfloat16_rn(y * __bfloat162float(w[c]));
  }
}

// nn::LayerNorm fast path with affine parameters (nn/ops.rs:1021-1043): f32 (x - mean) * rstd * w + b, one rounding
__global__ void te_layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                    bf16* __restrict__ out, int D, float eps) {
  const bf16* xr = x + static_cast<long long>(blockIdx.x) * D;
  bf16* orow = out + static_cast<long long>(blockIdx.x) * D;
  float s = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = __bfloat162float(xr[c]);
    s += v;
    s2 += v * v;
  }
  block_sum2(s, s2);
  const float mean = s / D;
  const float inv_std = 1.0f / sqrtf(s2 / D - mean * mean + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float y = (__bfloat162float(xr[c]) - mean) * inv_std * __bfloat162float(w[c]) + __bfloat162float(b[c]);
    orow[c] = __float2bfloat16_rn(y);
  }
}

// out = bf16(a * b)   (hidden_gelu.broadcast_mul(hidden_linear), t5/mod.rs:193)
__global__ void te_mul_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out,
                              long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) * __bfloat162float(b[i]));
}

// quick_gelu: xs * sigmoid(xs * 1.702), every op rounded to bf16 (clip/text.rs:15-19; sigmoid = recip(1 + exp(-v)))
__global__ void te_quick_gelu_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long n, float k_bf16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  const float t = rbf(rbf(v * k_bf16) + 0.0f);  // Tensor::affine(1.702, 0)
  const float e = rbf(expf(-t));
  const float sg = rbf(1.0f / rbf(1.0f + e));
  out[i] = __float2bfloat16_rn(v * sg);
}

// Attention over a fused [rows, 3*inner] q|k|v buffer, head dim 64, L <= 512.  One warp per (batch, head, query).
//   MODE 0 (T5, t5/mod.rs:306-387): s = bf16(q.k); s = bf16(s + bias[bucket(j - i)][h]); bf16 softmax_last_dim
//           (max-subtract, exp, f32 row sum, divide: each rounded); out = bf16(p.v).  No 1/sqrt(d) scaling.
//   MODE 1 (CLIP, clip/text.rs:117-146): q = bf16(q * bf16(scale)); then f32: s = q.k + (j > i ? f32::MIN : 0);
//           f32 softmax; out = bf16(p.v).
template <int MODE>
__global__ void __launch_bounds__(128) te_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ ctx, int B,
                                                           int H, int L, const bf16* __restrict__ rel_emb,
                                                           const int32_t* __restrict__ rel_bucket, float scale_bf16) {
  __shared__ float sq[4][TE_HEAD_DIM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * 4 + warp;  // (b, h, i)
  const long long total = static_cast<long long>(B) * H * L;
  const bool active = gw < total;
  const int i = active ? static_cast<int>(gw % L) : 0;
  const int h = active ? static_cast<int>((gw / L) % H) : 0;
  const int b = active ? static_cast<int>(gw / (static_cast<long long>(L) * H)) : 0;
  const int inner = H * TE_HEAD_DIM;
  const long long ld = 3LL * inner;
  const bf16* base = qkv + static_cast<long long>(b) * L * ld + h * TE_HEAD_DIM;
  {
    const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(base + static_cast<long long>(i) * ld + 2 * lane);
    float q0 = __bfloat162float(q2.x), q1 = __bfloat162float(q2.y);
    if (MODE == 1) q0 = rbf(rbf(q0 * scale_bf16) + 0.0f), q1 = rbf(rbf(q1 * scale_bf16) + 0.0f);
    sq[warp][2 * lane] = q0;
    sq[warp][2 * lane + 1] = q1;
  }
  __syncwarp();
  constexpr int SLOTS = TE_MAX_L / 32;
  float s[SLOTS];
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    const int j = t * 32 + lane;
    s[t] = -INFINITY;
    if (t * 32 < L && j < L) {
      const uint4* kr = reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * ld + inner);
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < TE_HEAD_DIM / 8; ++c) {
        const uint4 u = kr[c];
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dot = fmaf(sq[warp][c * 8 + 2 * e], bf_lo(w4[e]), dot);
          dot = fmaf(sq[warp][c * 8 + 2 * e + 1], bf_hi(w4[e]), dot);
        }
      }
      if (MODE == 0) {
        const float bias = __bfloat162float(rel_emb[rel_bucket[j - i + L - 1] * H + h]);
        s[t] = rbf(rbf(dot) + bias);
      } else {
        s[t] = dot + (j > i ? -3.4028234663852886e38f : 0.0f);
      }
      m = fmaxf(m, s[t]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    const int j = t * 32 + lane;
    if (t * 32 < L && j < L) {
      s[t] = MODE == 0 ? rbf(expf(rbf(s[t] - m))) : expf(s[t] - m);
      sum += s[t];
    } else {
      s[t] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (MODE == 0) sum = rbf(sum);
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) s[t] = MODE == 0 ? rbf(s[t] / sum) : s[t] / sum;
  // out[d] = sum_j p_j v[j][d]; this lane owns d = 2*lane, 2*lane+1
  float o0 = 0.f, o1 = 0.f;
  const bf16* vbase = base + 2 * inner + 2 * lane;
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    if (t * 32 < L) {
      const int jn = min(32, L - t * 32);
      for (int jj = 0; jj < jn; ++jj) {
        const float p = __shfl_sync(0xffffffffu, s[t], jj);
        const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vbase + static_cast<long long>(t * 32 + jj) * ld);
        o0 = fmaf(p, __bfloat162float(v2.x), o0);
        o1 = fmaf(p, __bfloat162float(v2.y), o1);
      }
    }
  }
  if (active) {
    bf16* dst = ctx + (static_cast<long long>(b) * L + i) * inner + h * TE_HEAD_DIM + 2 * lane;
    *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
  }
}

// pooled[b, :] = hidden[b, argmax_l ids[b, l], :]   (clip/text.rs:306-315)
__global__ void te_pool_argmax_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ hidden,
                                      bf16* __restrict__ pooled, int L, int D) {
  const int b = blockIdx.x;
  __shared__ int s_idx;
  if (threadIdx.x == 0) {
    int best = 0, bv = ids[static_cast<long long>(b) * L];
    for (int l = 1; l < L; ++l) {
      const int v = ids[static_cast<long long>(b) * L + l];
      if (v > bv) bv = v, best = l;  // first maximum, like Tensor::argmax
    }
    s_idx = best;
  }
  __syncthreads();
  const bf16* src = hidden + (static_cast<long long>(b) * L + s_idx) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) pooled[static_cast<long long>(b) * D + c] = src[c];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static float te_host_rbf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

struct TeTensor {
  bf16* dev = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct TeStore {
  std::map<std::string, TeTensor> raw;
  std::vector<void*> owned;  // fused copies
  ~TeStore() {
    for (auto& kv : raw) cudaFree(kv.second.dev);
    for (void* p : owned) cudaFree(p);
  }
  int load(const char* name, const void* data, int32_t dtype, const int64_t* shape, int32_t rank, int32_t is_device,
           cudaStream_t st) {
    FB_REQUIRE(name && data && shape, "load_weight: null argument");
    FB_REQUIRE(dtype == FLUXB200_DT_BF16, std::string("text encoders take bf16 tensors only (") + name + ")");
    TeTensor t;
    t.shape.assign(shape, shape + rank);
    const size_t bytes = static_cast<size_t>(t.numel()) * 2;
    FB_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.dev), bytes ? bytes : 16));
    FB_CHECK_CUDA(cudaMemcpyAsync(t.dev, data, bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    auto it = raw.find(name);
    if (it != raw.end()) {
      cudaFree(it->second.dev);
      raw.erase(it);
    }
    raw[name] = t;
    return 0;
  }
  int get(const std::string& name, std::vector<int64_t> shape, const bf16** out) const {
    auto it = raw.find(name);
    FB_REQUIRE(it != raw.end(), "missing tensor " + name);
    FB_REQUIRE(it->second.shape == shape, "tensor " + name + " has an unexpected shape");
    *out = it->second.dev;
    return 0;
  }
  // concatenate [N_i, K] weights along N into one owned [sum N_i, K] buffer
  int fuse(const std::vector<std::string>& names, const std::vector<int64_t>& Ns, int64_t K, const bf16** out,
           cudaStream_t st) {
    int64_t N = 0;
    for (auto n : Ns) N += n;
    bf16* buf = nullptr;
    FB_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&buf), static_cast<size_t>(N) * K * 2));
    owned.push_back(buf);
    int64_t off = 0;
    for (size_t i = 0; i < names.size(); ++i) {
      const bf16* src = nullptr;
      const bool vec = K == 1;
      int rc = vec ? get(names[i], {Ns[i]}, &src) : get(names[i], {Ns[i], K}, &src);
      if (rc) return rc;
      FB_CHECK_CUDA(cudaMemcpyAsync(buf + off * K, src, static_cast<size_t>(Ns[i]) * K * 2, cudaMemcpyDeviceToDevice, st));
      off += Ns[i];
    }
    *out = buf;
    return 0;
  }
};

static int te_gemm(const bf16* a, const bf16* w, const bf16* bias, bf16* out, int M, int N, int K, const bf16* res,
                   int n_split, bf16* out1, int act0, cudaStream_t st) {
  GemmDesc d;
  d.a = a, d.lda = K, d.w = w, d.ldb = K, d.M = M, d.N = N, d.K = K;
  d.out0 = out, d.ld0 = n_split > 0 ? n_split : N;
  if (n_split > 0) d.out1 = out1, d.ld1 = N - n_split, d.n_split = n_split;
  d.act0 = act0;
  d.bias = bias, d.bias_mode = bias ? BIAS_AFTER_ROUND : BIAS_NONE;  // nn::Linear / matmul + broadcast_add
  d.res = res;
  return launch_gemm(&d, 1, st);
}

static size_t te_align(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace fb

using namespace fb;

#define TE_TRY(x)        \
  do {                   \
    int _rc = (x);       \
    if (_rc) return _rc; \
  } while (0)

// ================================================================================================
// T5
// ================================================================================================
struct T5Layer {
  const bf16 *ln0, *qkv, *o, *ln1, *wi, *wo;
};
struct fluxb200_t5 {
  fluxb200_t5_config cfg{};
  bool finalized = false;
  TeStore store;
  const bf16 *shared = nullptr, *rel_emb = nullptr, *final_ln = nullptr;
  std::vector<T5Layer> layers;
};

extern "C" {

int fluxb200_t5_create(const fluxb200_t5_config* cfg, fluxb200_t5** out) {
  FB_REQUIRE(cfg && out, "t5_create: null argument");
  FB_REQUIRE(cfg->d_kv == TE_HEAD_DIM, "t5: d_kv must be 64");
  FB_REQUIRE(cfg->d_model % 8 == 0 && cfg->d_ff % 256 == 0, "t5: d_model % 8 == 0 and d_ff % 256 == 0 required");
  FB_REQUIRE(cfg->num_layers > 0 && cfg->num_heads > 0 && cfg->vocab_size > 0, "t5: bad config");
  FB_REQUIRE(cfg->relative_attention_num_buckets >= 4, "t5: bad relative_attention_num_buckets");
  auto* m = new fluxb200_t5();
  m->cfg = *cfg;
  *out = m;
  return 0;
}
void fluxb200_t5_destroy(fluxb200_t5* m) { delete m; }

int fluxb200_t5_load_weight(fluxb200_t5* m, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                            int32_t rank, int32_t is_device, fluxb200_stream_t stream) {
  FB_REQUIRE(m && !m->finalized, "t5_load_weight: null model or already finalized");
  return m->store.load(name, data, dtype, shape, rank, is_device, static_cast<cudaStream_t>(stream));
}

int fluxb200_t5_finalize(fluxb200_t5* m, fluxb200_stream_t stream) {
  FB_REQUIRE(m && !m->finalized, "t5_finalize: null model or called twice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto& c = m->cfg;
  const int64_t D = c.d_model, inner = static_cast<int64_t>(c.num_heads) * c.d_kv, F = c.d_ff;
  // T5EncoderModel::new (t5/mod.rs:645-657): shared | decoder.embed_tokens | encoder.embed_tokens
  const char* emb_names[3] = {"shared.weight", "decoder.embed_tokens.weight", "encoder.embed_tokens.weight"};
  for (const char* n : emb_names)
    if (!m->shared && m->store.raw.count(n)) TE_TRY(m->store.get(n, {c.vocab_size, D}, &m->shared));
  FB_REQUIRE(m->shared, "missing tensor shared.weight");
  TE_TRY(m->store.get("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                      {c.relative_attention_num_buckets, c.num_heads}, &m->rel_emb));
  TE_TRY(m->store.get("encoder.final_layer_norm.weight", {D}, &m->final_ln));
  m->layers.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "encoder.block." + std::to_string(i) + ".layer.";
    T5Layer& l = m->layers[i];
    TE_TRY(m->store.get(p + "0.layer_norm.weight", {D}, &l.ln0));
    TE_TRY(m->store.get(p + "1.layer_norm.weight", {D}, &l.ln1));
    TE_TRY(m->store.fuse({p + "0.SelfAttention.q.weight", p + "0.SelfAttention.k.weight", p + "0.SelfAttention.v.weight"},
                         {inner, inner, inner}, D, &l.qkv, st));
    TE_TRY(m->store.get(p + "0.SelfAttention.o.weight", {D, inner}, &l.o));
    TE_TRY(m->store.fuse({p + "1.DenseReluDense.wi_0.weight", p + "1.DenseReluDense.wi_1.weight"}, {F, F}, D, &l.wi, st));
    TE_TRY(m->store.get(p + "1.DenseReluDense.wo.weight", {D, F}, &l.wo));
  }
  FB_CHECK_CUDA(cudaStreamSynchronize(st));
  // the un-fused q/k/v/wi copies are no longer needed
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "encoder.block." + std::to_string(i) + ".layer.";
    for (const char* s : {"0.SelfAttention.q.weight", "0.SelfAttention.k.weight", "0.SelfAttention.v.weight",
                          "1.DenseReluDense.wi_0.weight", "1.DenseReluDense.wi_1.weight"}) {
      auto it = m->store.raw.find(p + s);
      if (it != m->store.raw.end()) {
        cudaFree(it->second.dev);
        m->store.raw.erase(it);
      }
    }
  }
  m->finalized = true;
  return 0;
}

int fluxb200_t5_workspace_size(const fluxb200_t5* m, int32_t batch, int32_t L, uint64_t* bytes) {
  FB_REQUIRE(m && bytes, "t5_workspace_size: null argument");
  FB_REQUIRE(batch > 0 && L > 0 && L <= TE_MAX_L, "t5: sequence length must be in 1..512");
  const auto& c = m->cfg;
  const size_t rows = static_cast<size_t>(batch) * L;
  const size_t inner = static_cast<size_t>(c.num_heads) * c.d_kv;
  size_t t = 0;
  t += te_align(rows * c.d_model * 2) * 2;  // x, normed
  t += te_align(rows * 3 * inner * 2);      // qkv
  t += te_align(rows * inner * 2);          // ctx
  t += te_align(rows * c.d_ff * 2) * 3;     // gelu(wi_0), wi_1, product
  t += te_align((2 * static_cast<size_t>(L) - 1) * 4);
  *bytes = t;
  return 0;
}

int fluxb200_t5_forward(fluxb200_t5* m, const int32_t* ids, void* out, int32_t batch, int32_t L, void* workspace,
                        uint64_t workspace_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(m && m->finalized, "t5_forward: model not finalized");
  FB_REQUIRE(ids && out && workspace, "t5_forward: null argument");
  uint64_t need = 0;
  TE_TRY(fluxb200_t5_workspace_size(m, batch, L, &need));
  FB_REQUIRE(workspace_bytes >= need, "t5_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto& c = m->cfg;
  const int rows = batch * L, D = c.d_model, H = c.num_heads, inner = H * c.d_kv, F = c.d_ff;
  uint8_t* p = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) {
    void* r = p;
    p += te_align(bytes);
    return r;
  };
  bf16* x = static_cast<bf16*>(take(static_cast<size_t>(rows) * D * 2));
  bf16* nrm = static_cast<bf16*>(take(static_cast<size_t>(rows) * D * 2));
  bf16* qkv = static_cast<bf16*>(take(static_cast<size_t>(rows) * 3 * inner * 2));
  bf16* ctx = static_cast<bf16*>(take(static_cast<size_t>(rows) * inner * 2));
  bf16* g = static_cast<bf16*>(take(static_cast<size_t>(rows) * F * 2));
  bf16* hl = static_cast<bf16*>(take(static_cast<size_t>(rows) * F * 2));
  bf16* prod = static_cast<bf16*>(take(static_cast<size_t>(rows) * F * 2));
  int32_t* bucket = static_cast<int32_t*>(take((2 * static_cast<size_t>(L) - 1) * 4));

  // relative-position buckets for j - i in [-(L-1), L-1], exactly as written at t5/mod.rs:334-372 (host logf)
  {
    std::vector<int32_t> hb(2 * L - 1);
    const int nb_total = c.relative_attention_num_buckets, nb = nb_total / 2, max_exact = nb / 2;
    const float base = static_cast<float>(c.relative_attention_max_distance) / max_exact;
    for (int rel = -(L - 1); rel <= L - 1; ++rel) {
      int v;
      if (rel > 0) {  // i < j
        if (rel < max_exact) v = rel + nb;
        else {
          const float bb = logf(static_cast<float>(rel) / max_exact) / logf(base) * (nb - max_exact);
          v = std::min(max_exact + nb + static_cast<int>(bb), nb_total - 1);
        }
      } else {
        const int d = -rel;
        if (d < max_exact) v = d;
        else {
          const float bb = logf(static_cast<float>(d) / max_exact) / logf(base) * (nb - max_exact);
          v = std::min(max_exact + static_cast<int>(bb), nb - 1);
        }
      }
      hb[rel + L - 1] = v;
    }
    FB_CHECK_CUDA(cudaMemcpyAsync(bucket, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice, st));
    FB_CHECK_CUDA(cudaStreamSynchronize(st));  // hb is a stack temporary
  }

  count_launch(KK_MISC);
  te_embed_kernel<<<rows, 256, 0, st>>>(ids, m->shared, nullptr, x, rows, L, D, c.vocab_size);
  const long long att_warps = static_cast<long long>(batch) * H * L;
  for (int i = 0; i < c.num_layers; ++i) {
    const T5Layer& l = m->layers[i];
    count_launch(KK_MISC, 4);
    te_rmsnorm_kernel<<<rows, 256, 0, st>>>(x, l.ln0, nrm, D, c.layer_norm_epsilon);
    TE_TRY(te_gemm(nrm, l.qkv, nullptr, qkv, rows, 3 * inner, D, nullptr, 0, nullptr, ACT_NONE, st));
    te_attention_kernel<0><<<static_cast<unsigned>((att_warps + 3) / 4), 128, 0, st>>>(qkv, ctx, batch, H, L, m->rel_emb,
                                                                                      bucket, 1.0f);
    TE_TRY(te_gemm(ctx, l.o, nullptr, x, rows, D, inner, x, 0, nullptr, ACT_NONE, st));  // x = bf16(x + bf16(ctx.o^T))
    te_rmsnorm_kernel<<<rows, 256, 0, st>>>(x, l.ln1, nrm, D, c.layer_norm_epsilon);
    // [NewGelu(wi_0 x) | wi_1 x] in one GEMM, GELU on the first d_ff columns
    TE_TRY(te_gemm(nrm, l.wi, nullptr, g, rows, 2 * F, D, nullptr, F, hl, ACT_GELU, st));
    const long long n = static_cast<long long>(rows) * F;
    te_mul_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(g, hl, prod, n);
    TE_TRY(te_gemm(prod, l.wo, nullptr, x, rows, D, F, x, 0, nullptr, ACT_NONE, st));
  }
  count_launch(KK_MISC);
  te_rmsnorm_kernel<<<rows, 256, 0, st>>>(x, m->final_ln, static_cast<bf16*>(out), D, c.layer_norm_epsilon);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ================================================================================================
// CLIP
// ================================================================================================
struct ClipLayer {
  const bf16 *ln1w, *ln1b, *qkv_w, *qkv_b, *out_w, *out_b, *ln2w, *ln2b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};
struct fluxb200_clip {
  fluxb200_clip_config cfg{};
  bool finalized = false;
  TeStore store;
  const bf16 *tok = nullptr, *pos = nullptr, *fln_w = nullptr, *fln_b = nullptr;
  std::vector<ClipLayer> layers;
};

extern "C" {

int fluxb200_clip_create(const fluxb200_clip_config* cfg, fluxb200_clip** out) {
  FB_REQUIRE(cfg && out, "clip_create: null argument");
  FB_REQUIRE(cfg->num_attention_heads > 0 && cfg->projection_dim == cfg->num_attention_heads * TE_HEAD_DIM,
             "clip: head dim must be 64");
  FB_REQUIRE(cfg->intermediate_size % 8 == 0 && cfg->num_hidden_layers > 0 && cfg->vocab_size > 0, "clip: bad config");
  FB_REQUIRE(cfg->max_position_embeddings > 0 && cfg->max_position_embeddings <= TE_MAX_L, "clip: bad max positions");
  auto* m = new fluxb200_clip();
  m->cfg = *cfg;
  *out = m;
  return 0;
}
void fluxb200_clip_destroy(fluxb200_clip* m) { delete m; }

int fluxb200_clip_load_weight(fluxb200_clip* m, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                              int32_t rank, int32_t is_device, fluxb200_stream_t stream) {
  FB_REQUIRE(m && !m->finalized, "clip_load_weight: null model or already finalized");
  return m->store.load(name, data, dtype, shape, rank, is_device, static_cast<cudaStream_t>(stream));
}

int fluxb200_clip_finalize(fluxb200_clip* m, fluxb200_stream_t stream) {
  FB_REQUIRE(m && !m->finalized, "clip_finalize: null model or called twice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto& c = m->cfg;
  const int64_t D = c.projection_dim, I = c.intermediate_size;
  TE_TRY(m->store.get("embeddings.token_embedding.weight", {c.vocab_size, D}, &m->tok));
  TE_TRY(m->store.get("embeddings.position_embedding.weight", {c.max_position_embeddings, D}, &m->pos));
  TE_TRY(m->store.get("final_layer_norm.weight", {D}, &m->fln_w));
  TE_TRY(m->store.get("final_layer_norm.bias", {D}, &m->fln_b));
  m->layers.resize(c.num_hidden_layers);
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const std::string p = "encoder.layers." + std::to_string(i) + ".";
    ClipLayer& l = m->layers[i];
    TE_TRY(m->store.get(p + "layer_norm1.weight", {D}, &l.ln1w));
    TE_TRY(m->store.get(p + "layer_norm1.bias", {D}, &l.ln1b));
    TE_TRY(m->store.get(p + "layer_norm2.weight", {D}, &l.ln2w));
    TE_TRY(m->store.get(p + "layer_norm2.bias", {D}, &l.ln2b));
    TE_TRY(m->store.fuse({p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight", p + "self_attn.v_proj.weight"},
                         {D, D, D}, D, &l.qkv_w, st));
    TE_TRY(m->store.fuse({p + "self_attn.q_proj.bias", p + "self_attn.k_proj.bias", p + "self_attn.v_proj.bias"},
                         {D, D, D}, 1, &l.qkv_b, st));
    TE_TRY(m->store.get(p + "self_attn.out_proj.weight", {D, D}, &l.out_w));
    TE_TRY(m->store.get(p + "self_attn.out_proj.bias", {D}, &l.out_b));
    TE_TRY(m->store.get(p + "mlp.fc1.weight", {I, D}, &l.fc1_w));
    TE_TRY(m->store.get(p + "mlp.fc1.bias", {I}, &l.fc1_b));
    TE_TRY(m->store.get(p + "mlp.fc2.weight", {D, I}, &l.fc2_w));
    TE_TRY(m->store.get(p + "mlp.fc2.bias", {D}, &l.fc2_b));
  }
  FB_CHECK_CUDA(cudaStreamSynchronize(st));
  m->finalized = true;
  return 0;
}

int fluxb200_clip_workspace_size(const fluxb200_clip* m, int32_t batch, int32_t L, uint64_t* bytes) {
  FB_REQUIRE(m && bytes, "clip_workspace_size: null argument");
  FB_REQUIRE(batch > 0 && L > 0 && L <= m->cfg.max_position_embeddings, "clip: sequence longer than max_position_embeddings");
  const auto& c = m->cfg;
  const size_t rows = static_cast<size_t>(batch) * L;
  size_t t = 0;
  t += te_align(rows * c.projection_dim * 2) * 3;  // x, normed, ctx
  t += te_align(rows * 3 * c.projection_dim * 2);  // qkv
  t += te_align(rows * c.intermediate_size * 2) * 2;
  *bytes = t;
  return 0;
}

int fluxb200_clip_forward(fluxb200_clip* m, const int32_t* ids, void* hidden_out, void* pooled_out, int32_t batch,
                          int32_t L, void* workspace, uint64_t workspace_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(m && m->finalized, "clip_forward: model not finalized");
  FB_REQUIRE(ids && workspace && (hidden_out || pooled_out), "clip_forward: null argument");
  uint64_t need = 0;
  TE_TRY(fluxb200_clip_workspace_size(m, batch, L, &need));
  FB_REQUIRE(workspace_bytes >= need, "clip_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto& c = m->cfg;
  const int rows = batch * L, D = c.projection_dim, H = c.num_attention_heads, I = c.intermediate_size;
  uint8_t* p = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) {
    void* r = p;
    p += te_align(bytes);
    return r;
  };
  bf16* x = static_cast<bf16*>(take(static_cast<size_t>(rows) * D * 2));
  bf16* nrm = static_cast<bf16*>(take(static_cast<size_t>(rows) * D * 2));
  bf16* ctx = static_cast<bf16*>(take(static_cast<size_t>(rows) * D * 2));
  bf16* qkv = static_cast<bf16*>(take(static_cast<size_t>(rows) * 3 * D * 2));
  bf16* h1 = static_cast<bf16*>(take(static_cast<size_t>(rows) * I * 2));
  bf16* h2 = static_cast<bf16*>(take(static_cast<size_t>(rows) * I * 2));
  const float scale_bf16 = te_host_rbf(1.0f / sqrtf(static_cast<float>(TE_HEAD_DIM)));
  const float k_bf16 = te_host_rbf(1.702f);
  const float eps = 1e-5f;  // clip/text.rs:193-198

  count_launch(KK_MISC);
  te_embed_kernel<<<rows, 256, 0, st>>>(ids, m->tok, m->pos, x, rows, L, D, c.vocab_size);
  const long long att_warps = static_cast<long long>(batch) * H * L;
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const ClipLayer& l = m->layers[i];
    count_launch(KK_MISC, 4);
    te_layernorm_kernel<<<rows, 256, 0, st>>>(x, l.ln1w, l.ln1b, nrm, D, eps);
    TE_TRY(te_gemm(nrm, l.qkv_w, l.qkv_b, qkv, rows, 3 * D, D, nullptr, 0, nullptr, ACT_NONE, st));
    te_attention_kernel<1><<<static_cast<unsigned>((att_warps + 3) / 4), 128, 0, st>>>(qkv, ctx, batch, H, L, nullptr,
                                                                                      nullptr, scale_bf16);
    TE_TRY(te_gemm(ctx, l.out_w, l.out_b, x, rows, D, D, x, 0, nullptr, ACT_NONE, st));
    te_layernorm_kernel<<<rows, 256, 0, st>>>(x, l.ln2w, l.ln2b, nrm, D, eps);
    TE_TRY(te_gemm(nrm, l.fc1_w, l.fc1_b, h1, rows, I, D, nullptr, 0, nullptr, ACT_NONE, st));
    const long long n = static_cast<long long>(rows) * I;
    te_quick_gelu_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(h1, h2, n, k_bf16);
    TE_TRY(te_gemm(h2, l.fc2_w, l.fc2_b, x, rows, D, I, x, 0, nullptr, ACT_NONE, st));
  }
  count_launch(KK_MISC, 2);
  bf16* hid = hidden_out ? static_cast<bf16*>(hidden_out) : nrm;
  te_layernorm_kernel<<<rows, 256, 0, st>>>(x, m->fln_w, m->fln_b, hid, D, eps);
  if (pooled_out) te_pool_argmax_kernel<<<batch, 256, 0, st>>>(ids, hid, static_cast<bf16*>(pooled_out), L, D);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
