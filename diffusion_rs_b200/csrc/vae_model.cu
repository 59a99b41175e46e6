// fluxb200 — VAE decoder driver behind the C ABI (AutoEncoderKl::decode / Decoder::forward,
// diffusion_rs_core/src/models/vaes/vae.rs:371-455).  Activations are NHWC internally so that every convolution
// is an implicit GEMM fed by TMA boxes and every 1x1 conv / attention projection is a plain GEMM.
#include <math.h>

#include <map>
#include <string>
#include <vector>

#include "fluxb200.h"
#include "internal.h"
#include "kernels.h"

namespace fb {

struct ConvW {
  bf16* w = nullptr;  // [Cout, k, k, Cin]
  bf16* bias = nullptr;
  int Cin = 0, Cout = 0, k = 0;
};
struct NormW {
  bf16 *w = nullptr, *b = nullptr;
  int C = 0;
};
struct Resnet {
  NormW n1, n2;
  ConvW c1, c2, sc;
  bool has_sc = false;
};
struct VaeRaw {
  void* dev = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

}  // namespace fb
using namespace fb;

struct fluxb200_vae {
  fluxb200_vae_config cfg{};
  bool finalized = false;
  std::map<std::string, VaeRaw> raw;
  std::vector<void*> owned;
  ConvW conv_in, conv_out;
  Resnet mid1, mid2;
  bool has_attn = false;
  NormW attn_norm;
  ConvW attn_q, attn_k, attn_v, attn_o;  // Linear [C,C] used as 1x1 convs (vae.rs:46-82)
  std::vector<std::vector<Resnet>> up;
  std::vector<ConvW> upsamplers;
  NormW norm_out;
};

namespace fb {

static int conv_run(const ConvW& c, const bf16* x, bf16* out, const bf16* res, int N, int H, int W, cudaStream_t st) {
  GemmDesc d;
  d.a = x, d.w = c.w, d.ldb = static_cast<int64_t>(c.k) * c.k * c.Cin;
  d.M = N * H * W, d.N = c.Cout, d.K = c.k * c.k * c.Cin;
  d.out0 = out, d.ld0 = c.Cout;
  d.bias = c.bias, d.bias_mode = c.bias ? BIAS_AFTER_ROUND : BIAS_NONE;  // Conv2d::forward adds the bias as a bf16 op
  d.res = res;
  if (c.k == 1 && c.Cin % 8 == 0) {
    d.lda = c.Cin;  // 1x1 conv on NHWC == plain GEMM over pixels
  } else {
    d.conv = 1, d.cN = N, d.cH = H, d.cW = W, d.cC = c.Cin, d.ksize = c.k;
  }
  return launch_gemm(&d, 1, st);
}

struct VaeWs {
  bf16 *X, *T1, *T2, *scores, *q, *k, *v, *vt;
  double* stats;
  size_t total = 0;
};

static size_t au(size_t v) { return (v + 1023) / 1024 * 1024; }

static VaeWs vae_carve(const fluxb200_vae* v, void* base, int B, int h, int w) {
  VaeWs s{};
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += au(bytes);
    return r;
  };
  // largest activation: walk the decoder geometry
  const int* ch = v->cfg.block_out_channels;
  size_t maxel = static_cast<size_t>(h) * w * ch[3];
  int H = h, W = w;
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int cout = ch[3 - lvl];
    const int cin = lvl == 0 ? ch[3] : ch[3 - lvl + 1];
    maxel = std::max(maxel, static_cast<size_t>(H) * W * std::max(cin, cout));
    if (lvl != 3) {
      H *= 2, W *= 2;
      maxel = std::max(maxel, static_cast<size_t>(H) * W * cout);
    }
  }
  maxel *= B;
  s.stats = static_cast<double*>(take(sizeof(double) * B * 32 * 2));
  s.X = static_cast<bf16*>(take(maxel * 2));
  s.T1 = static_cast<bf16*>(take(maxel * 2));
  s.T2 = static_cast<bf16*>(take(maxel * 2));
  const size_t hw = static_cast<size_t>(h) * w;
  const int cm = ch[3];
  if (v->has_attn) {
    s.scores = static_cast<bf16*>(take(hw * hw * 2));  // one batch element at a time
    s.q = static_cast<bf16*>(take(B * hw * cm * 2));
    s.k = static_cast<bf16*>(take(B * hw * cm * 2));
    s.v = static_cast<bf16*>(take(B * hw * cm * 2));
    s.vt = static_cast<bf16*>(take(hw * cm * 2));
  }
  s.total = off;
  return s;
}

static int resnet_run(const fluxb200_vae* v, const Resnet& r, VaeWs& s, int N, int H, int W, cudaStream_t st) {
  // ResnetBlock::forward vae.rs:158-171
  const int HW = H * W;
  const float eps = 1e-6f;
  int rc;
  if ((rc = launch_groupnorm_silu(s.X, r.n1.w, r.n1.b, s.T1, N, HW, r.n1.C, v->cfg.norm_num_groups, eps, 1, s.stats, st)))
    return rc;
  if ((rc = conv_run(r.c1, s.T1, s.T2, nullptr, N, H, W, st))) return rc;
  if ((rc = launch_groupnorm_silu(s.T2, r.n2.w, r.n2.b, s.T1, N, HW, r.n2.C, v->cfg.norm_num_groups, eps, 1, s.stats, st)))
    return rc;
  if (r.has_sc) {
    if ((rc = conv_run(r.sc, s.X, s.T2, nullptr, N, H, W, st))) return rc;  // xs.apply(conv_shortcut)
    if ((rc = conv_run(r.c2, s.T1, s.T2, s.T2, N, H, W, st))) return rc;    // + h, in place
    std::swap(s.X, s.T2);
  } else {
    if ((rc = conv_run(r.c2, s.T1, s.X, s.X, N, H, W, st))) return rc;  // xs + h, in place
  }
  return 0;
}

static int attn_run(const fluxb200_vae* v, VaeWs& s, int N, int H, int W, cudaStream_t st) {
  // AttnBlock::forward vae.rs:96-110; single head, scores / softmax in the model dtype (vae.rs:28-33)
  const int HW = H * W, C = v->attn_q.Cin;
  int rc;
  if ((rc = launch_groupnorm_silu(s.X, v->attn_norm.w, v->attn_norm.b, s.T1, N, HW, C, v->cfg.norm_num_groups, 1e-6f, 0,
                                  s.stats, st)))
    return rc;
  if ((rc = conv_run(v->attn_q, s.T1, s.q, nullptr, N, H, W, st))) return rc;
  if ((rc = conv_run(v->attn_k, s.T1, s.k, nullptr, N, H, W, st))) return rc;
  if ((rc = conv_run(v->attn_v, s.T1, s.v, nullptr, N, H, W, st))) return rc;
  const float scale = static_cast<float>(1.0 / sqrt(static_cast<double>(C)));
  const float scale_b = __bfloat162float(__float2bfloat16_rn(scale));  // `* scale_factor` is a bf16 affine op
  for (int n = 0; n < N; ++n) {
    const size_t o = static_cast<size_t>(n) * HW * C;
    GemmDesc d;  // scores = bf16(bf16(q.k^T) * scale)
    d.a = s.q + o, d.lda = C, d.w = s.k + o, d.ldb = C, d.M = HW, d.N = HW, d.K = C;
    d.out0 = s.scores, d.ld0 = HW, d.alpha = scale_b;
    if ((rc = launch_gemm(&d, 1, st))) return rc;
    if ((rc = launch_softmax_rows_bf16(s.scores, HW, HW, st))) return rc;
    if ((rc = launch_transpose_2d(s.v + o, s.vt, HW, C, st))) return rc;  // V^T [C, HW]: K-major B operand
    GemmDesc e;  // out = bf16(P.V) written into T1's slot for this batch element
    e.a = s.scores, e.lda = HW, e.w = s.vt, e.ldb = HW, e.M = HW, e.N = C, e.K = HW;
    e.out0 = s.T1 + o, e.ld0 = C;
    if ((rc = launch_gemm(&e, 1, st))) return rc;
  }
  // xs.apply(out) + init_xs, in place on X
  return conv_run(v->attn_o, s.T1, s.X, s.X, N, H, W, st);
}

// z (NHWC [B,h,w,16]) must already sit in s.T1; result NHWC [B,8h,8w,3] is left in s.T2
static int decode_core(const fluxb200_vae* v, VaeWs& s, int B, int h, int w, cudaStream_t st) {
  int rc;
  int H = h, W = w;
  if ((rc = conv_run(v->conv_in, s.T1, s.X, nullptr, B, H, W, st))) return rc;
  if ((rc = resnet_run(v, v->mid1, s, B, H, W, st))) return rc;
  if (v->has_attn)
    if ((rc = attn_run(v, s, B, H, W, st))) return rc;
  if ((rc = resnet_run(v, v->mid2, s, B, H, W, st))) return rc;
  for (int lvl = 0; lvl < 4; ++lvl) {
    for (auto& r : v->up[lvl])
      if ((rc = resnet_run(v, r, s, B, H, W, st))) return rc;
    if (lvl != 3) {
      const ConvW& c = v->upsamplers[lvl];
      if ((rc = launch_upsample2x_nhwc(s.X, s.T1, B, H, W, c.Cin, st))) return rc;
      H *= 2, W *= 2;
      if ((rc = conv_run(c, s.T1, s.T2, nullptr, B, H, W, st))) return rc;
      std::swap(s.X, s.T2);
    }
  }
  if ((rc = launch_groupnorm_silu(s.X, v->norm_out.w, v->norm_out.b, s.T1, B, H * W, v->norm_out.C,
                                  v->cfg.norm_num_groups, 1e-6f, 1, s.stats, st)))
    return rc;
  return conv_run(v->conv_out, s.T1, s.T2, nullptr, B, H, W, st);
}

static int vae_check_ws(fluxb200_vae* v, int B, int h, int w, void* ws, uint64_t ws_bytes, VaeWs* out) {
  FB_REQUIRE(v && v->finalized, "vae not finalized");
  FB_REQUIRE(B >= 1 && h > 0 && w > 0, "vae: bad geometry");
  FB_REQUIRE(ws != nullptr, "vae: null workspace");
  uint8_t* base = reinterpret_cast<uint8_t*>(au(reinterpret_cast<uintptr_t>(ws)));
  VaeWs s = vae_carve(v, base, B, h, w);
  FB_REQUIRE(s.total + (base - static_cast<uint8_t*>(ws)) <= ws_bytes,
             "vae workspace too small: need " + std::to_string(s.total + 1024) + " bytes");
  *out = s;
  return 0;
}

}  // namespace fb

extern "C" {

int fluxb200_vae_create(const fluxb200_vae_config* cfg, fluxb200_vae** out) {
  FB_REQUIRE(cfg && out, "vae_create: null argument");
  FB_REQUIRE(cfg->norm_num_groups == 32, "vae: norm_num_groups must be 32");
  FB_REQUIRE(cfg->latent_channels % 8 == 0, "vae: latent_channels must be a multiple of 8");
  int dev = 0, major = 0;
  FB_CHECK_CUDA(cudaGetDevice(&dev));
  FB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FB_REQUIRE(major == 10, "fluxb200 needs an sm_100a (Blackwell B200) device; there is no fallback path");
  auto* v = new fluxb200_vae();
  v->cfg = *cfg;
  *out = v;
  return 0;
}

void fluxb200_vae_destroy(fluxb200_vae* v) {
  if (!v) return;
  for (auto& kv : v->raw) cudaFree(kv.second.dev);
  for (void* p : v->owned) cudaFree(p);
  delete v;
}

int fluxb200_vae_load_weight(fluxb200_vae* v, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                             int32_t rank, int32_t is_device, fluxb200_stream_t stream) {
  FB_REQUIRE(v && name && data && shape, "vae_load_weight: null argument");
  FB_REQUIRE(!v->finalized, "vae_load_weight after finalize");
  FB_REQUIRE(dtype == FLUXB200_DT_BF16, "vae weights must be bf16 (the reference casts every tensor to the model dtype)");
  VaeRaw t;
  t.shape.assign(shape, shape + rank);
  const size_t bytes = static_cast<size_t>(t.numel()) * 2;
  FB_CHECK_CUDA(cudaMalloc(&t.dev, bytes ? bytes : 16));
  FB_CHECK_CUDA(cudaMemcpyAsync(t.dev, data, bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                static_cast<cudaStream_t>(stream)));
  auto it = v->raw.find(name);
  if (it != v->raw.end()) {
    cudaFree(it->second.dev);
    v->raw.erase(it);
  }
  v->raw[name] = t;
  return 0;
}

int fluxb200_vae_finalize(fluxb200_vae* v, fluxb200_stream_t stream) {
  FB_REQUIRE(v && !v->finalized, "vae_finalize: bad state");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto get = [&](const std::string& name) -> VaeRaw* {
    auto it = v->raw.find(name);
    return it == v->raw.end() ? nullptr : &it->second;
  };
  auto conv = [&](const std::string& p, int cin, int cout, int k, ConvW& c) -> int {
    VaeRaw* w = get(p + ".weight");
    VaeRaw* b = get(p + ".bias");
    FB_REQUIRE(w && b, "missing tensor " + p + ".{weight,bias}");
    FB_REQUIRE(w->numel() == static_cast<int64_t>(cout) * cin * k * k && b->numel() == cout,
               "shape mismatch for " + p);
    c.Cin = cin, c.Cout = cout, c.k = k;
    c.bias = static_cast<bf16*>(b->dev);
    if (k == 1) {
      c.w = static_cast<bf16*>(w->dev);  // [Cout, Cin(,1,1)] is already K-major
    } else {
      void* d = nullptr;
      FB_CHECK_CUDA(cudaMalloc(&d, static_cast<size_t>(w->numel()) * 2));
      v->owned.push_back(d);
      c.w = static_cast<bf16*>(d);
      if (int rc = launch_repack_conv_weight(static_cast<const bf16*>(w->dev), c.w, cout, cin, k * k, st)) return rc;
    }
    return 0;
  };
  auto norm = [&](const std::string& p, int C, NormW& n) -> int {
    VaeRaw* w = get(p + ".weight");
    VaeRaw* b = get(p + ".bias");
    FB_REQUIRE(w && b && w->numel() == C && b->numel() == C, "missing or bad tensor " + p + ".{weight,bias}");
    n.w = static_cast<bf16*>(w->dev), n.b = static_cast<bf16*>(b->dev), n.C = C;
    return 0;
  };
  auto resnet = [&](const std::string& p, int cin, int cout, Resnet& r) -> int {
    int rc;
    if ((rc = norm(p + ".norm1", cin, r.n1))) return rc;
    if ((rc = conv(p + ".conv1", cin, cout, 3, r.c1))) return rc;
    if ((rc = norm(p + ".norm2", cout, r.n2))) return rc;
    if ((rc = conv(p + ".conv2", cout, cout, 3, r.c2))) return rc;
    r.has_sc = cin != cout;
    if (r.has_sc)
      if ((rc = conv(p + ".conv_shortcut", cin, cout, 1, r.sc))) return rc;
    return 0;
  };
#define TRY(x)           \
  do {                   \
    int _rc = (x);       \
    if (_rc) return _rc; \
  } while (0)
  const int* ch = v->cfg.block_out_channels;
  int block_in = ch[3];
  const std::string d = "decoder.";
  TRY(conv(d + "conv_in", v->cfg.latent_channels, block_in, 3, v->conv_in));
  TRY(resnet(d + "mid_block.resnets.0", block_in, block_in, v->mid1));
  v->has_attn = v->cfg.mid_block_add_attention != 0;
  if (v->has_attn) {
    const std::string a = d + "mid_block.attentions.0.";
    TRY(norm(a + "group_norm", block_in, v->attn_norm));
    TRY(conv(a + "to_q", block_in, block_in, 1, v->attn_q));
    TRY(conv(a + "to_k", block_in, block_in, 1, v->attn_k));
    TRY(conv(a + "to_v", block_in, block_in, 1, v->attn_v));
    TRY(conv(a + "to_out.0", block_in, block_in, 1, v->attn_o));
  }
  TRY(resnet(d + "mid_block.resnets.1", block_in, block_in, v->mid2));
  v->up.assign(4, {});
  v->upsamplers.assign(3, {});
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int block_out = ch[3 - lvl];
    for (int j = 0; j <= v->cfg.layers_per_block; ++j) {
      Resnet r;
      TRY(resnet(d + "up_blocks." + std::to_string(lvl) + ".resnets." + std::to_string(j), block_in, block_out, r));
      v->up[lvl].push_back(r);
      block_in = block_out;
    }
    if (lvl != 3)
      TRY(conv(d + "up_blocks." + std::to_string(lvl) + ".upsamplers.0.conv", block_in, block_in, 3, v->upsamplers[lvl]));
  }
  TRY(norm(d + "conv_norm_out", ch[0], v->norm_out));
  TRY(conv(d + "conv_out", ch[0], v->cfg.out_channels, 3, v->conv_out));
#undef TRY
  FB_CHECK_CUDA(cudaStreamSynchronize(st));
  v->finalized = true;
  return 0;
}

int fluxb200_vae_workspace_size(const fluxb200_vae* v, int32_t batch, int32_t h, int32_t w, uint64_t* bytes) {
  FB_REQUIRE(v && bytes && v->finalized, "vae_workspace_size: vae not finalized");
  *bytes = vae_carve(v, nullptr, batch, h, w).total + 1024;
  return 0;
}

int fluxb200_vae_decode(fluxb200_vae* v, const void* z, void* out, int32_t batch, int32_t h, int32_t w,
                        void* workspace, uint64_t workspace_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(z && out, "vae_decode: null tensor");
  VaeWs s;
  if (int rc = vae_check_ws(v, batch, h, w, workspace, workspace_bytes, &s)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = launch_nchw_to_nhwc(static_cast<const bf16*>(z), s.T1, batch, v->cfg.latent_channels, h, w, st))
    return rc;
  if (int rc = decode_core(v, s, batch, h, w, st)) return rc;
  return launch_nhwc_to_nchw(s.T2, static_cast<bf16*>(out), batch, v->cfg.out_channels, 8 * h, 8 * w, st);
}

int fluxb200_vae_decode_packed_u8(fluxb200_vae* v, const void* packed, void* out_u8, int32_t batch, int32_t h2,
                                  int32_t w2, int32_t nchw, void* workspace, uint64_t workspace_bytes,
                                  fluxb200_stream_t stream) {
  FB_REQUIRE(packed && out_u8, "vae_decode_packed_u8: null tensor");
  FB_REQUIRE(v && v->cfg.latent_channels == 16, "packed latents need 16 latent channels");
  const int h = 2 * h2, w = 2 * w2;
  VaeWs s;
  if (int rc = vae_check_ws(v, batch, h, w, workspace, workspace_bytes, &s)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = launch_unpack_latents(static_cast<const bf16*>(packed), s.T1, batch, h2, w2,
                                     static_cast<float>(1.0 / static_cast<double>(v->cfg.scaling_factor)),
                                     v->cfg.shift_factor, st))
    return rc;
  if (int rc = decode_core(v, s, batch, h, w, st)) return rc;
  return launch_postprocess_u8(s.T2, static_cast<uint8_t*>(out_u8), batch, v->cfg.out_channels, 8 * h, 8 * w, nchw, st);
}

int fluxb200_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* res, void* out,
                         int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize,
                         fluxb200_stream_t stream) {
  ConvW c;
  c.w = static_cast<bf16*>(const_cast<void*>(w_packed));
  c.bias = static_cast<bf16*>(const_cast<void*>(bias));
  c.Cin = Cin, c.Cout = Cout, c.k = ksize;
  return conv_run(c, static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(res), N, H, W,
                  static_cast<cudaStream_t>(stream));
}

int fluxb200_repack_conv_weight(const void* w, void* out, int32_t Cout, int32_t Cin, int32_t ksize,
                                fluxb200_stream_t stream) {
  return launch_repack_conv_weight(static_cast<const bf16*>(w), static_cast<bf16*>(out), Cout, Cin, ksize * ksize,
                                   static_cast<cudaStream_t>(stream));
}

int fluxb200_groupnorm_nhwc(const void* x, const void* weight, const void* bias, void* out, int32_t N, int32_t HW,
                            int32_t C, int32_t groups, float eps, int32_t silu, void* stats_scratch,
                            fluxb200_stream_t stream) {
  return launch_groupnorm_silu(static_cast<const bf16*>(x), static_cast<const bf16*>(weight),
                               static_cast<const bf16*>(bias), static_cast<bf16*>(out), N, HW, C, groups, eps, silu,
                               static_cast<double*>(stats_scratch), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
