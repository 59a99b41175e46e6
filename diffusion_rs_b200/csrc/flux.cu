// fluxb200 — host-side FLUX MMDiT driver behind the C ABI (Flux::new / Flux::forward / Sampler::sample).
//
// Mirrors diffusion_rs_core/src/models/flux/model.rs:722-833 block for block, but every Linear is one launch of the
// tcgen05 GEMM (q|k|v and q|k|v|proj_mlp fused along N, txt+img streams grouped in one launch), attention is the
// tcgen05 flash kernel, and LayerNorm/modulate, QK-norm/RoPE, gate/residual, GELU are fused into their neighbours.
#include <math.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "fluxb200.h"
#include "internal.h"
#include "kernels.h"

namespace fb {

static constexpr int D = 3072;       // HIDDEN_SIZE model.rs:17
static constexpr int MLP_D = 4 * D;  // MLP_RATIO model.rs:16
static constexpr int HEAD_DIM = 128;
static constexpr int MAX_STEPS = 1024;

// ------------------------------------------------------------------------------------------------
// RoPE table kernel: EmbedNd / rope() (model.rs:65-84, 142-157) with the reference's bf16 op order:
//   inv_freq (f64 -> f32 -> bf16), freqs = bf16(pos * inv_freq), cos/sin = bf16(cosf/sinf(freqs))
// ------------------------------------------------------------------------------------------------
__constant__ float c_inv_freq[64];
__constant__ int c_freq_axis[64];

__global__ void pe_table_kernel(const bf16* __restrict__ ids, int rows_per_batch, int batch, int L, int l_off,
                                bf16* __restrict__ pe_cos, bf16* __restrict__ pe_sin, uint2* __restrict__ pe2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * rows_per_batch * 64) return;
  const int f = i & 63;
  const int row = i >> 6;
  const int b = row / rows_per_batch, r = row - b * rows_per_batch;
  const float pos = __bfloat162float(ids[static_cast<long long>(row) * 3 + c_freq_axis[f]]);
  const float fr = __bfloat162float(__float2bfloat16_rn(pos * c_inv_freq[f]));
  const long long o = (static_cast<long long>(b) * L + l_off + r) * 64 + f;
  const bf16 c = __float2bfloat16_rn(cosf(fr)), sn = __float2bfloat16_rn(sinf(fr));
  pe_cos[o] = c;
  pe_sin[o] = sn;
  // packed form for the fused GEMM epilogue: {(cos, sin), (-sin, cos)}, laid out [batch][pair][token]
  const uint32_t cu = __bfloat16_as_ushort(c), su = __bfloat16_as_ushort(sn);
  pe2[(static_cast<long long>(b) * 64 + f) * L + l_off + r] = make_uint2(cu | (su << 16), (su ^ 0x8000u) | (cu << 16));
}

static float host_rbf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return x;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

static int upload_rope_constants() {
  float inv[64];
  int axis[64];
  const int axes[3] = {16, 56, 56};  // AXES_DIM model.rs:18
  int k = 0;
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < axes[a]; i += 2) {
      const float f = 1.0f / static_cast<float>(pow(10000.0, static_cast<double>(i) / axes[a]));  // model.rs:73
      inv[k] = host_rbf(f);  // .to_dtype(pos.dtype()) model.rs:77
      axis[k] = a;
      ++k;
    }
  FB_CHECK_CUDA(cudaMemcpyToSymbol(c_inv_freq, inv, sizeof(inv)));
  FB_CHECK_CUDA(cudaMemcpyToSymbol(c_freq_axis, axis, sizeof(axis)));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight store
// ------------------------------------------------------------------------------------------------
struct RawTensor {
  void* dev = nullptr;
  int dtype = 0;
  std::vector<int64_t> shape;
  size_t bytes = 0;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

enum QType { Q_DENSE = 0, Q_NF4 = 1, Q_FP4 = 2, Q_INT8 = 3, Q_Q4K = 4 };

struct Member {  // one reference Linear inside a fused group
  std::string name;
  int N = 0;
  int row_off = 0;
  QType q = Q_DENSE;
  const uint8_t* packed = nullptr;  // nf4/fp4 nibbles, int8 weights or Q4_K blocks (owned by the raw store)
  float* absmax = nullptr;          // f32 absmax after nested de-quantisation (owned, cudaMalloc)
  const float* scb = nullptr;
  int blocksize = 64;
};

struct FusedLinear {
  int N = 0, K = 0;
  bf16* w = nullptr;     // dense [N,K] (owned) or nullptr when quantised
  bf16* bias = nullptr;  // [N] (owned)
  bool quant = false;
  int bias_mode3 = BIAS_FUSED;  // rank-3 call sites: cuBLASLt fused bias (dense) vs separate add (bnb)
  std::vector<Member> members;
  QuantB qb;  // quantised: operand description for the fused-dequant GEMM producer
};

struct DoubleBlock {
  FusedLinear img_mod, txt_mod, img_qkv, txt_qkv, img_proj, txt_proj, img_mlp1, img_mlp2, txt_mlp1, txt_mlp2;
  bf16 *img_nq = nullptr, *img_nk = nullptr, *txt_nq = nullptr, *txt_nk = nullptr;
};
struct SingleBlock {
  FusedLinear mod, lin1 /* q|k|v|proj_mlp */, lin2;
  bf16 *nq = nullptr, *nk = nullptr;
};

struct Workspace {
  uint2* pe2;
  bf16 *pe_cos, *pe_sin, *temb, *gemb, *e1, *e2, *e3, *e4, *vec, *svec, *mod_all;
  bf16 *img, *txt, *x, *xm, *qkv, *Q, *K, *V, *attn_img, *attn_txt, *big, *pred, *txt_cache;
  float* tvals;
  size_t total = 0;
};

}  // namespace fb

using namespace fb;

struct fluxb200_model {
  fluxb200_flux_config cfg{};
  bool finalized = false;
  std::map<std::string, RawTensor> raw;
  FusedLinear img_in, txt_in, time1, time2, vecin1, vecin2, guid1, guid2, final_mod, final_proj;
  std::vector<DoubleBlock> dbl;
  std::vector<SingleBlock> sgl;
  std::vector<void*> owned;  // every cudaMalloc the model owns besides `raw`
  // modulation GEMV job table (dense models): all 2*19 + 38 + 1 projections of silu(vec_) in one launch
  GemvJob* mod_jobs_dev = nullptr;
  int mod_njobs = 0, mod_total_rows = 0;
  std::vector<long long> mod_off;  // offset (elements, per batch row) of each job's output inside mod_all
  long long mod_row_elems = 0;
  bf16* wscratch = nullptr;  // dequantised-weight staging (quantised models only)
  size_t wscratch_elems = 0;
  bool any_quant = false;
  // last forward's geometry, for the parity taps
  int last_B = 0, last_limg = 0, last_ltxt = 0;
  Workspace last_ws{};
};

namespace fb {

static int dtype_size(int dt) {
  switch (dt) {
    case FLUXB200_DT_BF16:
    case FLUXB200_DT_F16:
      return 2;
    case FLUXB200_DT_F32:
      return 4;
    case FLUXB200_DT_U8:
    case FLUXB200_DT_I8:
      return 1;
    default:
      return 0;
  }
}

static int dev_alloc(fluxb200_model* m, void** p, size_t bytes) {
  FB_CHECK_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  m->owned.push_back(*p);
  return 0;
}

static const RawTensor* find(const fluxb200_model* m, const std::string& name) {
  auto it = m->raw.find(name);
  return it == m->raw.end() ? nullptr : &it->second;
}

// tiny JSON number extractor for BnbQuantState (bitsandbytes/mod.rs:43-51)
static bool json_number(const std::string& js, const char* key, double* out) {
  const std::string k = std::string("\"") + key + "\"";
  size_t p = js.find(k);
  if (p == std::string::npos) return false;
  p = js.find(':', p);
  if (p == std::string::npos) return false;
  *out = strtod(js.c_str() + p + 1, nullptr);
  return true;
}

// Build one fused Linear from `names` (concatenated along N).  Mirrors diffusion_rs_backend::linear (lib.rs:223-252):
// `weight.absmax` or `SCB` present -> BnbLinear, else dense; Q4_K tensors -> GgufMatMul semantics.
static int build_linear(fluxb200_model* m, FusedLinear& fl, const std::vector<std::string>& names, int K,
                        const std::vector<int>& Ns, cudaStream_t st) {
  fl.K = K;
  fl.N = 0;
  for (int n : Ns) fl.N += n;
  FB_REQUIRE(K % 8 == 0, "linear in-dim must be a multiple of 8");
  // bias (always a plain tensor)
  if (int rc = dev_alloc(m, reinterpret_cast<void**>(&fl.bias), static_cast<size_t>(fl.N) * 2)) return rc;
  int row = 0;
  bool any_q = false, all_q = true;
  for (size_t i = 0; i < names.size(); ++i) {
    Member mb;
    mb.name = names[i];
    mb.N = Ns[i];
    mb.row_off = row;
    const RawTensor* w = find(m, names[i] + ".weight");
    FB_REQUIRE(w != nullptr, "missing tensor " + names[i] + ".weight");
    const RawTensor* b = find(m, names[i] + ".bias");
    FB_REQUIRE(b != nullptr, "missing tensor " + names[i] + ".bias");
    FB_REQUIRE(b->dtype == FLUXB200_DT_BF16 && b->numel() == Ns[i], "bad bias for " + names[i]);
    FB_CHECK_CUDA(cudaMemcpyAsync(fl.bias + row, b->dev, static_cast<size_t>(Ns[i]) * 2, cudaMemcpyDeviceToDevice, st));
    const int64_t numel = static_cast<int64_t>(Ns[i]) * K;
    if (find(m, names[i] + ".weight.absmax")) {
      // ---- bitsandbytes 4-bit (bitsandbytes/mod.rs:137-222) ----
      const RawTensor* qs_nf4 = find(m, names[i] + ".weight.quant_state.bitsandbytes__nf4");
      const RawTensor* qs_fp4 = find(m, names[i] + ".weight.quant_state.bitsandbytes__fp4");
      FB_REQUIRE(qs_nf4 || qs_fp4, "`BnbLinear` expects fp4/nf4 or int8 layers: " + names[i]);
      const RawTensor* qs = qs_nf4 ? qs_nf4 : qs_fp4;
      std::string js(qs->bytes, '\0');
      FB_CHECK_CUDA(cudaMemcpy(&js[0], qs->dev, qs->bytes, cudaMemcpyDeviceToHost));
      double blocksize = 0, nested_bs = 0, nested_off = 0;
      FB_REQUIRE(json_number(js, "blocksize", &blocksize), "quant_state without blocksize: " + names[i]);
      const int supported[7] = {2048, 4096, 1024, 512, 256, 128, 64};  // SUPPORTED_BLOCKSIZE mod.rs:14
      bool ok = false;
      for (int s : supported) ok |= (s == static_cast<int>(blocksize));
      FB_REQUIRE(ok, "Blocksize of " + std::to_string(static_cast<int>(blocksize)) + " is not supported");
      FB_REQUIRE(w->dtype == FLUXB200_DT_U8 && w->numel() * 2 == numel, "bad packed 4-bit weight for " + names[i]);
      mb.q = qs_nf4 ? Q_NF4 : Q_FP4;
      mb.packed = static_cast<const uint8_t*>(w->dev);
      mb.blocksize = static_cast<int>(blocksize);
      const int64_t nabs = (numel + mb.blocksize - 1) / mb.blocksize;
      const RawTensor* am = find(m, names[i] + ".weight.absmax");
      if (int rc = dev_alloc(m, reinterpret_cast<void**>(&mb.absmax), static_cast<size_t>(nabs) * 4)) return rc;
      const RawTensor* nam = find(m, names[i] + ".weight.nested_absmax");
      if (nam) {
        // double quant: absmax = nested_code[absmax_u8] * nested_absmax[i / nested_blocksize] + nested_offset
        // (BnbLinear::dequantize_4bit mod.rs:230-239 -> blockwise int8 kernel dequant.cu:135-140)
        const RawTensor* ncode = find(m, names[i] + ".weight.nested_quant_map");
        FB_REQUIRE(ncode && ncode->numel() == 256 && ncode->dtype == FLUXB200_DT_F32, "bad nested_quant_map");
        FB_REQUIRE(json_number(js, "nested_blocksize", &nested_bs), "`nested_blocksize` must be present.");
        FB_REQUIRE(json_number(js, "nested_offset", &nested_off), "`offset` must be present.");
        FB_REQUIRE(am->dtype == FLUXB200_DT_U8 && am->numel() == nabs, "bad nested absmax for " + names[i]);
        std::vector<uint8_t> a8(nabs);
        std::vector<float> code(256), nabsmax(nam->numel()), out(nabs);
        FB_CHECK_CUDA(cudaMemcpy(a8.data(), am->dev, nabs, cudaMemcpyDeviceToHost));
        FB_CHECK_CUDA(cudaMemcpy(code.data(), ncode->dev, 1024, cudaMemcpyDeviceToHost));
        FB_CHECK_CUDA(cudaMemcpy(nabsmax.data(), nam->dev, nam->numel() * 4, cudaMemcpyDeviceToHost));
        const int nb = static_cast<int>(nested_bs);
        const float off = static_cast<float>(nested_off);
        for (int64_t j = 0; j < nabs; ++j) out[j] = code[a8[j]] * nabsmax[j / nb] + off;
        FB_CHECK_CUDA(cudaMemcpy(mb.absmax, out.data(), nabs * 4, cudaMemcpyHostToDevice));
      } else {
        FB_REQUIRE(am->dtype == FLUXB200_DT_F32 && am->numel() == nabs, "bad absmax for " + names[i]);
        FB_CHECK_CUDA(cudaMemcpyAsync(mb.absmax, am->dev, nabs * 4, cudaMemcpyDeviceToDevice, st));
      }
      any_q = true;
    } else if (find(m, names[i] + ".SCB")) {
      const RawTensor* scb = find(m, names[i] + ".SCB");
      FB_REQUIRE(w->dtype == FLUXB200_DT_I8 && w->numel() == numel, "bad int8 weight for " + names[i]);
      FB_REQUIRE(scb->dtype == FLUXB200_DT_F32 && scb->numel() == Ns[i], "bad SCB for " + names[i]);
      mb.q = Q_INT8;
      mb.packed = static_cast<const uint8_t*>(w->dev);
      mb.scb = static_cast<const float*>(scb->dev);
      any_q = true;
    } else if (w->dtype == FLUXB200_DT_Q4K) {
      FB_REQUIRE(K % 256 == 0, "Q4_K needs in-dim % 256 == 0: " + names[i]);
      FB_REQUIRE(w->numel() == numel, "bad Q4_K weight shape for " + names[i]);
      mb.q = Q_Q4K;
      mb.packed = static_cast<const uint8_t*>(w->dev);
      any_q = true;
    } else {
      FB_REQUIRE(w->dtype == FLUXB200_DT_BF16 && w->numel() == numel,
                 "shape mismatch for " + names[i] + ".weight, expected [" + std::to_string(Ns[i]) + ", " +
                     std::to_string(K) + "] bf16");
      all_q = false;
    }
    row += Ns[i];
    fl.members.push_back(mb);
  }
  FB_REQUIRE(!(any_q && !all_q), "mixed dense/quantised members in one fused linear: " + names[0]);
  fl.quant = any_q;
  if (!any_q) {
    if (int rc = dev_alloc(m, reinterpret_cast<void**>(&fl.w), static_cast<size_t>(fl.N) * K * 2)) return rc;
    for (auto& mb : fl.members) {
      const RawTensor* w = find(m, mb.name + ".weight");
      FB_CHECK_CUDA(cudaMemcpyAsync(fl.w + static_cast<size_t>(mb.row_off) * K, w->dev,
                                    static_cast<size_t>(mb.N) * K * 2, cudaMemcpyDeviceToDevice, st));
    }
    fl.bias_mode3 = BIAS_FUSED;
  } else {
    m->any_quant = true;
    m->wscratch_elems = std::max(m->wscratch_elems, static_cast<size_t>(fl.N) * K);
    // bnb: separate bf16 add after the matmul (bitsandbytes/mod.rs:301-312); gguf: f32 result + bias, one rounding
    fl.bias_mode3 = (fl.members[0].q == Q_Q4K) ? BIAS_FUSED : BIAS_AFTER_ROUND;
    FB_REQUIRE(fl.members.size() <= 4, "at most 4 fused quantised members");
    fl.qb.count = static_cast<int>(fl.members.size());
    for (size_t i = 0; i < fl.members.size(); ++i) {
      const Member& mb = fl.members[i];
      QuantMember& qm = fl.qb.m[i];
      qm.packed = mb.packed, qm.absmax = mb.absmax, qm.scb = mb.scb, qm.row_begin = mb.row_off, qm.blocksize = mb.blocksize;
      qm.kind = mb.q == Q_NF4 ? QB_NF4 : (mb.q == Q_FP4 ? QB_FP4 : (mb.q == Q_Q4K ? QB_Q4K : QB_INT8));
    }
  }
  return 0;
}

// After the dense copy is made the raw dense tensors are no longer needed.
static void drop_raw_dense(fluxb200_model* m, const FusedLinear& fl) {
  if (fl.quant) return;
  for (auto& mb : fl.members) {
    auto it = m->raw.find(mb.name + ".weight");
    if (it != m->raw.end()) {
      cudaFree(it->second.dev);
      m->raw.erase(it);
    }
  }
}

// Weight operand for the GEMM: dense pointer, or expand the quantised members into the staging buffer.
static int weight_operand(fluxb200_model* m, const FusedLinear& fl, const bf16** w, cudaStream_t st,
                          bool for_gemm = true) {
  if (!fl.quant) {
    *w = fl.w;
    return 0;
  }
  if (for_gemm && get_flag("fused_dequant") && fl.N % 128 == 0 && fl.K % 64 == 0) {
    *w = nullptr;  // gemm_for() switches to the fused-dequant producer
    return 0;
  }
  for (auto& mb : fl.members) {
    bf16* dst = m->wscratch + static_cast<size_t>(mb.row_off) * fl.K;
    const long long n = static_cast<long long>(mb.N) * fl.K;
    int rc = 0;
    switch (mb.q) {
      case Q_NF4:
      case Q_FP4:
        rc = launch_dequant_bnb4(mb.packed, mb.absmax, dst, mb.blocksize, n, mb.q == Q_NF4, st);
        break;
      case Q_INT8:
        rc = launch_dequant_int8(reinterpret_cast<const int8_t*>(mb.packed), mb.scb, dst, fl.K, n, st);
        break;
      case Q_Q4K:
        rc = launch_dequant_q4k(mb.packed, dst, n, st);
        break;
      default:
        return fail("weight_operand: bad quant type");
    }
    if (rc) return rc;
  }
  *w = m->wscratch;
  return 0;
}

// Dense: W read by TMA.  Quantised: the GEMM's producer warps expand the packed weights on the fly (fb::QuantB),
// unless fused de-quantisation is switched off ("fused_dequant" = 0), in which case `w` is the staging buffer.
static GemmDesc gemm_for(const FusedLinear& fl, const bf16* w, const bf16* a, int M, bf16* out, int64_t ldo) {
  GemmDesc d;
  d.a = a, d.lda = fl.K, d.w = w, d.ldb = fl.K;
  if (fl.quant && w == nullptr) d.qb = &fl.qb;
  d.M = M, d.N = fl.N, d.K = fl.K;
  d.out0 = out, d.ld0 = ldo;
  d.bias = fl.bias, d.bias_mode = fl.bias_mode3;
  return d;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void attach_qkrope(GemmDesc& g, const Workspace& w, const bf16* nq, const bf16* nk, int H, int L, int l_off,
                          int rows_per_batch, float eps);

static Workspace carve(const fluxb200_model* m, void* base, int B, int l_img, int l_txt) {
  Workspace w{};
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  const size_t L = static_cast<size_t>(l_img) + l_txt;
  const size_t Mi = static_cast<size_t>(B) * l_img, Mt = static_cast<size_t>(B) * l_txt, Mx = B * L;
  w.tvals = static_cast<float*>(take(MAX_STEPS * 16 * 4));  // per step: t[8] | guidance[8]
  w.pe_cos = static_cast<bf16*>(take(Mx * 64 * 2));
  w.pe_sin = static_cast<bf16*>(take(Mx * 64 * 2));
  w.pe2 = static_cast<uint2*>(take(Mx * 64 * 8));
  w.temb = static_cast<bf16*>(take(static_cast<size_t>(B) * 256 * 2));
  w.gemb = static_cast<bf16*>(take(static_cast<size_t>(B) * 256 * 2));
  w.e1 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.e2 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.e3 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.e4 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.vec = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.svec = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.mod_all = static_cast<bf16*>(take(static_cast<size_t>(B) * m->mod_row_elems * 2));
  w.img = static_cast<bf16*>(take(Mi * D * 2));
  w.txt = static_cast<bf16*>(take(Mt * D * 2));
  w.x = static_cast<bf16*>(take(Mx * D * 2));
  w.xm = static_cast<bf16*>(take(Mx * D * 2));
  w.qkv = static_cast<bf16*>(take(Mx * 3 * D * 2));
  w.Q = static_cast<bf16*>(take(Mx * D * 2));
  w.K = static_cast<bf16*>(take(Mx * D * 2));
  w.V = static_cast<bf16*>(take(Mx * D * 2));
  w.attn_img = static_cast<bf16*>(take(Mi * D * 2));
  w.attn_txt = static_cast<bf16*>(take(Mt * D * 2));
  w.big = static_cast<bf16*>(take(Mx * (D + MLP_D) * 2));  // single: [attn | gelu(mlp)]; double: MLP hidden
  w.pred = static_cast<bf16*>(take(Mi * 64 * 2));
  w.txt_cache = static_cast<bf16*>(take(Mt * D * 2));
  w.total = off;
  return w;
}

static void attach_qkrope(GemmDesc& g, const Workspace& w, const bf16* nq, const bf16* nk, int H, int L, int l_off,
                          int rows_per_batch, float eps) {
  g.qkrope = 1;
  g.qk_wq = nq, g.qk_wk = nk, g.qk_pe2 = w.pe2, g.qk_pe_bstride = static_cast<int64_t>(L) * 64;
  g.qk_Q = w.Q, g.qk_K = w.K, g.qk_V = w.V;
  g.qk_H = H, g.qk_L = L, g.qk_loff = l_off, g.qk_eps = eps;
  g.rows_per_batch = rows_per_batch;
}

// rank-2 Linear on [B, K] (MlpEmbedder / modulation): matmul -> bf16, + bias -> bf16.  `job` indexes the static
// job table built at finalize; `row_base` is that job's row_begin (0 for the embedder jobs).
static int small_linear(fluxb200_model* m, const FusedLinear& fl, int job, int row_base, const bf16* x, bf16* out_base,
                        int B, cudaStream_t st) {
  const bf16* w = nullptr;
  if (int rc = weight_operand(m, fl, &w, st, /*for_gemm=*/false)) return rc;  // quantised: expands into the staging buffer
  return launch_gemv_jobs(m->mod_jobs_dev + job, 1, row_base, fl.N, x, fl.K, B, fl.K, out_base, st);
}

}  // namespace fb

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int fluxb200_dequantize_q4k_bf16(const void* blocks, void* out, int64_t n, fluxb200_stream_t stream) {
  return launch_dequant_q4k(static_cast<const uint8_t*>(blocks), static_cast<bf16*>(out), n,
                            static_cast<cudaStream_t>(stream));
}

int fluxb200_model_create(const fluxb200_flux_config* cfg, fluxb200_model** out) {
  FB_REQUIRE(cfg && out, "model_create: null argument");
  FB_REQUIRE(cfg->num_attention_heads * HEAD_DIM == D, "num_attention_heads * 128 must equal HIDDEN_SIZE 3072");
  FB_REQUIRE(cfg->in_channels % 8 == 0 && cfg->joint_attention_dim % 8 == 0 && cfg->pooled_projection_dim % 8 == 0,
             "channel dims must be multiples of 8");
  int dev = 0, major = 0;
  FB_CHECK_CUDA(cudaGetDevice(&dev));
  FB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FB_REQUIRE(major == 10, "fluxb200 needs an sm_100a (Blackwell B200) device; there is no fallback path");
  if (int rc = upload_rope_constants()) return rc;
  auto* m = new fluxb200_model();
  m->cfg = *cfg;
  *out = m;
  return 0;
}

void fluxb200_model_destroy(fluxb200_model* m) {
  if (!m) return;
  for (auto& kv : m->raw) cudaFree(kv.second.dev);
  for (void* p : m->owned) cudaFree(p);
  if (m->mod_jobs_dev) cudaFree(m->mod_jobs_dev);
  if (m->wscratch) cudaFree(m->wscratch);
  delete m;
}

int fluxb200_model_load_weight(fluxb200_model* m, const char* name, const void* data, int32_t dtype,
                               const int64_t* shape, int32_t rank, int32_t is_device, fluxb200_stream_t stream) {
  FB_REQUIRE(m && name && data && shape, "load_weight: null argument");
  FB_REQUIRE(!m->finalized, "load_weight after finalize");
  RawTensor t;
  t.dtype = dtype;
  t.shape.assign(shape, shape + rank);
  const int64_t n = t.numel();
  if (dtype == FLUXB200_DT_Q4K) {
    FB_REQUIRE(n % 256 == 0, "Q4_K tensor element count must be a multiple of 256");
    t.bytes = static_cast<size_t>(n / 256) * 144;
  } else {
    FB_REQUIRE(dtype_size(dtype) > 0, "load_weight: unknown dtype");
    t.bytes = static_cast<size_t>(n) * dtype_size(dtype);
  }
  FB_CHECK_CUDA(cudaMalloc(&t.dev, t.bytes ? t.bytes : 16));
  cudaError_t e = cudaMemcpyAsync(t.dev, data, t.bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaFree(t.dev);
    return fail(std::string("load_weight copy failed: ") + cudaGetErrorString(e));
  }
  auto it = m->raw.find(name);
  if (it != m->raw.end()) {
    cudaFree(it->second.dev);
    m->raw.erase(it);
  }
  m->raw[name] = t;
  return 0;
}

int fluxb200_model_finalize(fluxb200_model* m, fluxb200_stream_t stream) {
  FB_REQUIRE(m, "finalize: null model");
  FB_REQUIRE(!m->finalized, "finalize called twice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto& c = m->cfg;
  std::vector<FusedLinear*> all;
  auto mk = [&](FusedLinear& fl, std::vector<std::string> names, int K, std::vector<int> Ns) -> int {
    int rc = build_linear(m, fl, names, K, Ns, st);
    if (rc == 0) all.push_back(&fl);
    return rc;
  };
  auto norm_w = [&](const std::string& name, bf16** dst) -> int {
    const RawTensor* t = find(m, name);
    FB_REQUIRE(t && t->dtype == FLUXB200_DT_BF16 && t->numel() == HEAD_DIM, "missing or bad tensor " + name);
    *dst = static_cast<bf16*>(t->dev);
    return 0;
  };
#define TRY(x)      \
  do {              \
    int _rc = (x);  \
    if (_rc) return _rc; \
  } while (0)
  TRY(mk(m->img_in, {"x_embedder"}, c.in_channels, {D}));
  TRY(mk(m->txt_in, {"context_embedder"}, c.joint_attention_dim, {D}));
  TRY(mk(m->time1, {"time_text_embed.timestep_embedder.linear_1"}, 256, {D}));
  TRY(mk(m->time2, {"time_text_embed.timestep_embedder.linear_2"}, D, {D}));
  TRY(mk(m->vecin1, {"time_text_embed.text_embedder.linear_1"}, c.pooled_projection_dim, {D}));
  TRY(mk(m->vecin2, {"time_text_embed.text_embedder.linear_2"}, D, {D}));
  if (c.guidance_embeds) {
    TRY(mk(m->guid1, {"time_text_embed.guidance_embedder.linear_1"}, 256, {D}));
    TRY(mk(m->guid2, {"time_text_embed.guidance_embedder.linear_2"}, D, {D}));
  }
  m->dbl.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    DoubleBlock& b = m->dbl[i];
    const std::string p = "transformer_blocks." + std::to_string(i) + ".";
    TRY(mk(b.img_mod, {p + "norm1.linear"}, D, {6 * D}));
    TRY(mk(b.txt_mod, {p + "norm1_context.linear"}, D, {6 * D}));
    TRY(mk(b.img_qkv, {p + "attn.to_q", p + "attn.to_k", p + "attn.to_v"}, D, {D, D, D}));
    TRY(mk(b.txt_qkv, {p + "attn.add_q_proj", p + "attn.add_k_proj", p + "attn.add_v_proj"}, D, {D, D, D}));
    TRY(mk(b.img_proj, {p + "attn.to_out.0"}, D, {D}));
    TRY(mk(b.txt_proj, {p + "attn.to_add_out"}, D, {D}));
    TRY(mk(b.img_mlp1, {p + "ff.net.0.proj"}, D, {MLP_D}));
    TRY(mk(b.img_mlp2, {p + "ff.net.2"}, MLP_D, {D}));
    TRY(mk(b.txt_mlp1, {p + "ff_context.net.0.proj"}, D, {MLP_D}));
    TRY(mk(b.txt_mlp2, {p + "ff_context.net.2"}, MLP_D, {D}));
    TRY(norm_w(p + "attn.norm_q.weight", &b.img_nq));
    TRY(norm_w(p + "attn.norm_k.weight", &b.img_nk));
    TRY(norm_w(p + "attn.norm_added_q.weight", &b.txt_nq));
    TRY(norm_w(p + "attn.norm_added_k.weight", &b.txt_nk));
  }
  m->sgl.resize(c.num_single_layers);
  for (int i = 0; i < c.num_single_layers; ++i) {
    SingleBlock& b = m->sgl[i];
    const std::string p = "single_transformer_blocks." + std::to_string(i) + ".";
    TRY(mk(b.mod, {p + "norm.linear"}, D, {3 * D}));
    TRY(mk(b.lin1, {p + "attn.to_q", p + "attn.to_k", p + "attn.to_v", p + "proj_mlp"}, D, {D, D, D, MLP_D}));
    TRY(mk(b.lin2, {p + "proj_out"}, D + MLP_D, {D}));
    TRY(norm_w(p + "attn.norm_q.weight", &b.nq));
    TRY(norm_w(p + "attn.norm_k.weight", &b.nk));
  }
  TRY(mk(m->final_mod, {"norm_out.linear"}, D, {2 * D}));
  TRY(mk(m->final_proj, {"proj_out"}, D, {c.in_channels}));
  FB_CHECK_CUDA(cudaStreamSynchronize(st));
  for (FusedLinear* fl : all) drop_raw_dense(m, *fl);

  // modulation outputs: one row of `mod_row_elems` bf16 per batch element, job order = double (img, txt) ..., single, final
  std::vector<const FusedLinear*> mods;
  for (auto& b : m->dbl) {
    mods.push_back(&b.img_mod);
    mods.push_back(&b.txt_mod);
  }
  for (auto& b : m->sgl) mods.push_back(&b.mod);
  mods.push_back(&m->final_mod);
  m->mod_off.clear();
  long long off = 0;
  for (auto* fl : mods) {
    m->mod_off.push_back(off);
    off += fl->N;
  }
  m->mod_row_elems = off;
  m->mod_njobs = static_cast<int>(mods.size());
  m->mod_total_rows = static_cast<int>(off);
  if (m->any_quant) {
    FB_CHECK_CUDA(cudaMalloc(&m->wscratch, m->wscratch_elems * 2));
  }
  // static GEMV job table: [modulation jobs ..., time1, time2, guid1, guid2, vecin1, vecin2]
  {
    std::vector<GemvJob> jobs;
    int row = 0;
    for (size_t i = 0; i < mods.size(); ++i) {
      GemvJob j;
      j.w = mods[i]->quant ? m->wscratch : mods[i]->w;
      j.bias = mods[i]->bias, j.out_off = m->mod_off[i], j.out_ld = m->mod_row_elems;
      j.N = mods[i]->N, j.row_begin = row;
      j.fused_bias = (mods[i]->quant && mods[i]->members[0].q == Q_Q4K) ? 1 : 0, j.pad_ = 0;
      row += mods[i]->N;
      jobs.push_back(j);
    }
    const FusedLinear* emb[6] = {&m->time1, &m->time2, &m->guid1, &m->guid2, &m->vecin1, &m->vecin2};
    for (int e = 0; e < 6; ++e) {
      GemvJob j;
      j.w = emb[e]->quant ? m->wscratch : emb[e]->w;
      j.bias = emb[e]->bias, j.out_off = 0, j.out_ld = D, j.N = emb[e]->N, j.row_begin = 0;
      j.fused_bias = (emb[e]->quant && !emb[e]->members.empty() && emb[e]->members[0].q == Q_Q4K) ? 1 : 0, j.pad_ = 0;
      jobs.push_back(j);
    }
    FB_CHECK_CUDA(cudaMalloc(&m->mod_jobs_dev, sizeof(GemvJob) * jobs.size()));
    FB_CHECK_CUDA(cudaMemcpy(m->mod_jobs_dev, jobs.data(), sizeof(GemvJob) * jobs.size(), cudaMemcpyHostToDevice));
  }
  m->finalized = true;
  return 0;
#undef TRY
}

int fluxb200_model_workspace_size(const fluxb200_model* m, int32_t batch, int32_t l_img, int32_t l_txt,
                                  uint64_t* bytes) {
  FB_REQUIRE(m && bytes && m->finalized, "workspace_size: model not finalized");
  FB_REQUIRE(batch >= 1 && batch <= 8 && l_img > 0 && l_txt > 0, "workspace_size: bad geometry (batch 1..8)");
  *bytes = carve(m, nullptr, batch, l_img, l_txt).total + 1024;
  return 0;
}

}  // extern "C"

namespace fb {

struct StepIO {
  const bf16* img_in;   // [B, l_img, 64]
  const bf16* img_ids;  // [B, l_img, 3]
  const bf16* txt_in;   // [B, l_txt, joint]
  const bf16* txt_ids;  // [B, l_txt, 3]
  const bf16* y;        // [B, pooled]
  bf16* out;            // [B, l_img, 64]
};

// Step-invariant part: RoPE table (model.rs:807-810) and txt_in (model.rs:811)
static int prepare_invariants(fluxb200_model* m, const Workspace& w, const StepIO& io, int B, int l_img, int l_txt,
                              cudaStream_t st) {
  const int L = l_img + l_txt;
  {
    const int n1 = B * l_txt * 64, n2 = B * l_img * 64;
    count_launch(KK_MISC, 2);
    pe_table_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(io.txt_ids, l_txt, B, L, 0, w.pe_cos, w.pe_sin, w.pe2);
    pe_table_kernel<<<(n2 + 255) / 256, 256, 0, st>>>(io.img_ids, l_img, B, L, l_txt, w.pe_cos, w.pe_sin, w.pe2);
    FB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// txt = txt_in(txt) — written into w.txt (double-block stream). Step-invariant but part of Flux::forward.
static int project_txt(fluxb200_model* m, const Workspace& w, const StepIO& io, int B, int l_txt, cudaStream_t st) {
  const bf16* wt = nullptr;
  if (int rc = weight_operand(m, m->txt_in, &wt, st)) return rc;
  GemmDesc d = gemm_for(m->txt_in, wt, io.txt_in, B * l_txt, w.txt, D);
  return launch_gemm(&d, 1, st);
}

// One Flux::forward given prepared pe table. If `txt_cached` is non-null it holds txt_in(txt) already.
static int forward_core(fluxb200_model* m, const Workspace& w, const StepIO& io, const float* t_dev,
                        const float* g_dev, int B, int l_img, int l_txt, const bf16* txt_cached, cudaStream_t st) {
  const auto& c = m->cfg;
  const int L = l_img + l_txt;
  const int Mi = B * l_img, Mt = B * l_txt, Mx = B * L;
  const int H = c.num_attention_heads;
  const float eps = 1e-6f;
  const float scale = 1.0f / sqrtf(static_cast<float>(HEAD_DIM));
  const long long PE_BS = static_cast<long long>(L) * 64;
  const bool fuse_qk = get_flag("qkrope_fusion") != 0;
  const int JE = m->mod_njobs;  // embedder jobs follow the modulation jobs in the static table
  int rc = 0;
#define TRY(x)          \
  do {                  \
    rc = (x);           \
    if (rc) return rc;  \
  } while (0)

  // ---- txt_in / img_in (model.rs:811-812) ----
  if (txt_cached) {
    FB_CHECK_CUDA(cudaMemcpyAsync(w.txt, txt_cached, static_cast<size_t>(Mt) * D * 2, cudaMemcpyDeviceToDevice, st));
  } else {
    TRY(project_txt(m, w, io, B, l_txt, st));
  }
  {
    const bf16* wi = nullptr;
    TRY(weight_operand(m, m->img_in, &wi, st));
    GemmDesc d = gemm_for(m->img_in, wi, io.img_in, Mi, w.img, D);
    TRY(launch_gemm(&d, 1, st));
  }
  // ---- vec_ (model.rs:813-820) ----
  TRY(launch_timestep_embedding(t_dev, w.temb, B, 256, st));
  TRY(small_linear(m, m->time1, JE + 0, 0, w.temb, w.e1, B, st));
  TRY(launch_silu(w.e1, w.e1, static_cast<long long>(B) * D, st));
  TRY(small_linear(m, m->time2, JE + 1, 0, w.e1, w.e2, B, st));
  const bf16* gvec = nullptr;
  if (c.guidance_embeds && g_dev) {
    TRY(launch_timestep_embedding(g_dev, w.gemb, B, 256, st));
    TRY(small_linear(m, m->guid1, JE + 2, 0, w.gemb, w.e1, B, st));
    TRY(launch_silu(w.e1, w.e1, static_cast<long long>(B) * D, st));
    TRY(small_linear(m, m->guid2, JE + 3, 0, w.e1, w.e3, B, st));
    gvec = w.e3;
  }
  TRY(small_linear(m, m->vecin1, JE + 4, 0, io.y, w.e1, B, st));
  TRY(launch_silu(w.e1, w.e1, static_cast<long long>(B) * D, st));
  TRY(small_linear(m, m->vecin2, JE + 5, 0, w.e1, w.e4, B, st));
  TRY(launch_vec_combine(w.e2, gvec, w.e4, w.vec, B * D, st));
  // ---- every modulation projection of silu(vec_) (model.rs:244-299, 694-698) ----
  TRY(launch_silu(w.vec, w.svec, static_cast<long long>(B) * D, st));
  if (!m->any_quant) {
    TRY(launch_gemv_jobs(m->mod_jobs_dev, m->mod_njobs, 0, m->mod_total_rows, w.svec, D, B, D, w.mod_all, st));
  } else {
    std::vector<const FusedLinear*> mods;
    for (auto& b : m->dbl) {
      mods.push_back(&b.img_mod);
      mods.push_back(&b.txt_mod);
    }
    for (auto& b : m->sgl) mods.push_back(&b.mod);
    mods.push_back(&m->final_mod);
    for (size_t i = 0; i < mods.size(); ++i)
      TRY(small_linear(m, *mods[i], static_cast<int>(i), static_cast<int>(m->mod_off[i]), w.svec, w.mod_all, B, st));
  }
  const long long mstride = m->mod_row_elems;
  auto modp = [&](int job, int chunk) { return w.mod_all + m->mod_off[job] + static_cast<long long>(chunk) * D; };

  // ---- double-stream blocks (model.rs:523-565) ----
  for (int i = 0; i < c.num_layers; ++i) {
    DoubleBlock& b = m->dbl[i];
    const int ji = 2 * i, jt = 2 * i + 1;  // modulation jobs: chunks = shift1, scale1, gate1, shift2, scale2, gate2
    bf16* xm_img = w.xm;
    bf16* xm_txt = w.xm + static_cast<size_t>(Mi) * D;
    bf16* qkv_img = w.qkv;
    bf16* qkv_txt = w.qkv + static_cast<size_t>(Mi) * 3 * D;
    TRY(launch_ln_modulate(w.img, l_img, 0, l_img, B, modp(ji, 0), modp(ji, 1), mstride, xm_img, D, eps, st));
    TRY(launch_ln_modulate(w.txt, l_txt, 0, l_txt, B, modp(jt, 0), modp(jt, 1), mstride, xm_txt, D, eps, st));
    {
      const bf16 *wi = nullptr, *wt = nullptr;
      GemmDesc g[2];
      if (!b.img_qkv.quant) {
        g[0] = gemm_for(b.img_qkv, b.img_qkv.w, xm_img, Mi, qkv_img, 3 * D);
        g[1] = gemm_for(b.txt_qkv, b.txt_qkv.w, xm_txt, Mt, qkv_txt, 3 * D);
        if (fuse_qk) {
          attach_qkrope(g[0], w, b.img_nq, b.img_nk, H, L, l_txt, l_img, eps);
          attach_qkrope(g[1], w, b.txt_nq, b.txt_nk, H, L, 0, l_txt, eps);
        }
        TRY(launch_gemm(g, 2, st));
      } else {  // one staging buffer: expand and run the two streams back to back
        TRY(weight_operand(m, b.img_qkv, &wi, st));
        g[0] = gemm_for(b.img_qkv, wi, xm_img, Mi, qkv_img, 3 * D);
        if (fuse_qk) attach_qkrope(g[0], w, b.img_nq, b.img_nk, H, L, l_txt, l_img, eps);
        TRY(launch_gemm(g, 1, st));
        TRY(weight_operand(m, b.txt_qkv, &wt, st));
        g[1] = gemm_for(b.txt_qkv, wt, xm_txt, Mt, qkv_txt, 3 * D);
        if (fuse_qk) attach_qkrope(g[1], w, b.txt_nq, b.txt_nk, H, L, 0, l_txt, eps);
        TRY(launch_gemm(g + 1, 1, st));
      }
    }
    if (!fuse_qk) {
      TRY(launch_qknorm_rope(qkv_txt, 3 * D, l_txt, B, H, L, 0, b.txt_nq, b.txt_nk, w.pe_cos, w.pe_sin, PE_BS, w.Q,
                             w.K, w.V, eps, st));
      TRY(launch_qknorm_rope(qkv_img, 3 * D, l_img, B, H, L, l_txt, b.img_nq, b.img_nk, w.pe_cos, w.pe_sin, PE_BS, w.Q,
                             w.K, w.V, eps, st));
    }
    {
      AttnDesc a;
      a.q = w.Q, a.k = w.K, a.v = w.V, a.B = B, a.H = H, a.L = L, a.scale = scale;
      a.out_a = w.attn_txt, a.ld_a = D, a.out_b = w.attn_img, a.ld_b = D, a.l_split = l_txt;
      TRY(launch_attention(a, st));
    }
    auto two = [&](FusedLinear& fi, FusedLinear& ft, const bf16* ai, const bf16* at, bf16* oi, bf16* ot, int64_t ldo,
                   int act, int gate_chunk, bool residual) -> int {
      GemmDesc g[2];
      const bf16 *wi = fi.w, *wt = ft.w;
      for (int s = 0; s < 2; ++s) {
        FusedLinear& f = s == 0 ? fi : ft;
        if (f.quant) {
          const bf16* wq = nullptr;
          if (int r = weight_operand(m, f, &wq, st)) return r;
          (s == 0 ? wi : wt) = wq;
        }
        g[s] = gemm_for(f, s == 0 ? wi : wt, s == 0 ? ai : at, s == 0 ? Mi : Mt, s == 0 ? oi : ot, ldo);
        g[s].act0 = act;
        if (residual) {
          g[s].gate = modp(s == 0 ? ji : jt, gate_chunk);
          g[s].gate_bstride = mstride;
          g[s].rows_per_batch = s == 0 ? l_img : l_txt;
          g[s].res = s == 0 ? oi : ot;
        }
        if (f.quant)
          if (int r = launch_gemm(&g[s], 1, st)) return r;
      }
      if (!fi.quant) return launch_gemm(g, 2, st);
      return 0;
    };
    // img/txt += gate1 * proj(attn)
    TRY(two(b.img_proj, b.txt_proj, w.attn_img, w.attn_txt, w.img, w.txt, D, ACT_NONE, 2, true));
    // MLP: x += gate2 * lin2(gelu(lin1(modulate2(LN(x)))))
    TRY(launch_ln_modulate(w.img, l_img, 0, l_img, B, modp(ji, 3), modp(ji, 4), mstride, xm_img, D, eps, st));
    TRY(launch_ln_modulate(w.txt, l_txt, 0, l_txt, B, modp(jt, 3), modp(jt, 4), mstride, xm_txt, D, eps, st));
    bf16* h_img = w.big;
    bf16* h_txt = w.big + static_cast<size_t>(Mi) * MLP_D;
    TRY(two(b.img_mlp1, b.txt_mlp1, xm_img, xm_txt, h_img, h_txt, MLP_D, ACT_GELU, 0, false));
    TRY(two(b.img_mlp2, b.txt_mlp2, h_img, h_txt, w.img, w.txt, D, ACT_NONE, 5, true));
  }

  // ---- cat(txt, img) (model.rs:827) ----
  for (int bi = 0; bi < B; ++bi) {
    FB_CHECK_CUDA(cudaMemcpyAsync(w.x + (static_cast<size_t>(bi) * L) * D, w.txt + static_cast<size_t>(bi) * l_txt * D,
                                  static_cast<size_t>(l_txt) * D * 2, cudaMemcpyDeviceToDevice, st));
    FB_CHECK_CUDA(cudaMemcpyAsync(w.x + (static_cast<size_t>(bi) * L + l_txt) * D,
                                  w.img + static_cast<size_t>(bi) * l_img * D, static_cast<size_t>(l_img) * D * 2,
                                  cudaMemcpyDeviceToDevice, st));
  }
  // ---- single-stream blocks (model.rs:638-662) ----
  const int CAT = D + MLP_D;
  for (int i = 0; i < c.num_single_layers; ++i) {
    SingleBlock& b = m->sgl[i];
    const int j = 2 * c.num_layers + i;  // chunks = shift, scale, gate
    TRY(launch_ln_modulate(w.x, L, 0, L, B, modp(j, 0), modp(j, 1), mstride, w.xm, D, eps, st));
    {
      const bf16* w1 = nullptr;
      TRY(weight_operand(m, b.lin1, &w1, st));
      GemmDesc g = gemm_for(b.lin1, w1, w.xm, Mx, w.qkv, 3 * D);
      g.n_split = 3 * D;  // q|k|v -> qkv buffer; proj_mlp -> gelu -> [attn | mlp] buffer at column D
      g.out1 = w.big, g.ld1 = CAT, g.col_off1 = D, g.act1 = ACT_GELU;
      if (fuse_qk) attach_qkrope(g, w, b.nq, b.nk, H, L, 0, L, eps);
      TRY(launch_gemm(&g, 1, st));
    }
    if (!fuse_qk)
      TRY(launch_qknorm_rope(w.qkv, 3 * D, L, B, H, L, 0, b.nq, b.nk, w.pe_cos, w.pe_sin, PE_BS, w.Q, w.K, w.V, eps, st));
    {
      AttnDesc a;
      a.q = w.Q, a.k = w.K, a.v = w.V, a.B = B, a.H = H, a.L = L, a.scale = scale;
      a.out_b = w.big, a.ld_b = CAT, a.l_split = 0;
      TRY(launch_attention(a, st));
    }
    {
      const bf16* w2 = nullptr;
      TRY(weight_operand(m, b.lin2, &w2, st));
      GemmDesc g = gemm_for(b.lin2, w2, w.big, Mx, w.x, D);
      g.gate = modp(j, 2), g.gate_bstride = mstride, g.rows_per_batch = L, g.res = w.x;
      TRY(launch_gemm(&g, 1, st));
    }
  }
  // ---- final layer on the img rows (model.rs:831-832, 694-705): chunks = scale, shift ----
  {
    const int jf = 2 * c.num_layers + c.num_single_layers;
    TRY(launch_ln_modulate(w.x, L, l_txt, l_img, B, modp(jf, 1), modp(jf, 0), mstride, w.xm, D, eps, st));
    const bf16* wf = nullptr;
    TRY(weight_operand(m, m->final_proj, &wf, st));
    GemmDesc g = gemm_for(m->final_proj, wf, w.xm, Mi, io.out, c.in_channels);
    TRY(launch_gemm(&g, 1, st));
  }
  m->last_B = B, m->last_limg = l_img, m->last_ltxt = l_txt, m->last_ws = w;
  return 0;
#undef TRY
}

static int check_ws(fluxb200_model* m, int B, int l_img, int l_txt, void* ws, uint64_t ws_bytes, Workspace* out) {
  FB_REQUIRE(m && m->finalized, "model not finalized");
  FB_REQUIRE(B >= 1 && B <= 8, "batch must be in 1..8 per call");
  FB_REQUIRE(l_img > 0 && l_txt > 0, "empty sequence");
  FB_REQUIRE(ws != nullptr, "null workspace");
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  const size_t slack = base - static_cast<uint8_t*>(ws);
  Workspace w = carve(m, base, B, l_img, l_txt);
  FB_REQUIRE(w.total + slack <= ws_bytes, "workspace too small: need " + std::to_string(w.total + 1024) + " bytes");
  *out = w;
  return 0;
}

}  // namespace fb

extern "C" {

int fluxb200_model_forward(fluxb200_model* m, const void* img, const void* img_ids, const void* txt,
                           const void* txt_ids, const void* timesteps, const void* y, const void* guidance,
                           void* out, int32_t batch, int32_t l_img, int32_t l_txt, void* workspace,
                           uint64_t workspace_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(img && img_ids && txt && txt_ids && timesteps && y && out, "forward: null tensor");
  Workspace w;
  if (int rc = check_ws(m, batch, l_img, l_txt, workspace, workspace_bytes, &w)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepIO io{static_cast<const bf16*>(img), static_cast<const bf16*>(img_ids), static_cast<const bf16*>(txt),
            static_cast<const bf16*>(txt_ids), static_cast<const bf16*>(y), static_cast<bf16*>(out)};
  if (int rc = prepare_invariants(m, w, io, batch, l_img, l_txt, st)) return rc;
  return forward_core(m, w, io, static_cast<const float*>(timesteps), static_cast<const float*>(guidance), batch,
                      l_img, l_txt, nullptr, st);
}

int fluxb200_model_denoise(fluxb200_model* m, void* img, const void* img_ids, const void* txt, const void* txt_ids,
                           const void* y, float guidance_scale, const double* timesteps, int32_t n_timesteps,
                           int32_t batch, int32_t l_img, int32_t l_txt, void* workspace, uint64_t workspace_bytes,
                           fluxb200_stream_t stream) {
  FB_REQUIRE(img && img_ids && txt && txt_ids && y && timesteps, "denoise: null tensor");
  FB_REQUIRE(n_timesteps >= 2, "denoise: need at least two timesteps");
  Workspace w;
  if (int rc = check_ws(m, batch, l_img, l_txt, workspace, workspace_bytes, &w)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepIO io{static_cast<const bf16*>(img), static_cast<const bf16*>(img_ids), static_cast<const bf16*>(txt),
            static_cast<const bf16*>(txt_ids), static_cast<const bf16*>(y), w.pred};
  if (int rc = prepare_invariants(m, w, io, batch, l_img, l_txt, st)) return rc;
  // txt_in(txt) does not depend on t: project once, keep it in the workspace and copy it into the stream buffer
  // at the start of every step.
  FB_REQUIRE(n_timesteps <= MAX_STEPS, "denoise: at most 1024 timesteps");
  const size_t txt_bytes = static_cast<size_t>(batch) * l_txt * D * 2;
  if (int rc = project_txt(m, w, io, batch, l_txt, st)) return rc;
  FB_CHECK_CUDA(cudaMemcpyAsync(w.txt_cache, w.txt, txt_bytes, cudaMemcpyDeviceToDevice, st));
  // t_vec = full(1) * t_curr (f32); guidance = full(guidance_scale) (pipelines/flux/mod.rs:300-304, sampling.rs:42).
  // All per-step scalars are uploaded once, before the loop.
  std::vector<float> hv(static_cast<size_t>(n_timesteps) * 16, 0.f);
  for (int s = 0; s + 1 < n_timesteps; ++s)
    for (int b = 0; b < batch; ++b) {
      hv[s * 16 + b] = static_cast<float>(1.0f * timesteps[s]);
      hv[s * 16 + 8 + b] = guidance_scale;
    }
  FB_CHECK_CUDA(cudaMemcpyAsync(w.tvals, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice, st));
  FB_CHECK_CUDA(cudaStreamSynchronize(st));  // `hv` is pageable host memory: make sure the copy has consumed it
  for (int s = 0; s + 1 < n_timesteps; ++s) {
    const double t_curr = timesteps[s], t_prev = timesteps[s + 1];
    const float* tv = w.tvals + s * 16;
    if (int rc = forward_core(m, w, io, tv, m->cfg.guidance_embeds ? tv + 8 : nullptr, batch, l_img, l_txt,
                              w.txt_cache, st))
      return rc;
    if (int rc = launch_euler(static_cast<bf16*>(img), w.pred, static_cast<float>(t_prev - t_curr),
                              static_cast<long long>(batch) * l_img * m->cfg.in_channels, st))
      return rc;
  }
  return 0;
}

int fluxb200_model_tap(fluxb200_model* m, int32_t which, void* out, uint64_t out_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(m && out && m->last_B > 0, "tap: no forward has run");
  const Workspace& w = m->last_ws;
  const size_t B = m->last_B, li = m->last_limg, lt = m->last_ltxt, L = li + lt;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (which) {
    case 0: src = w.vec, bytes = B * D * 2; break;
    case 1: src = w.img, bytes = B * li * D * 2; break;
    case 2: src = w.txt, bytes = B * lt * D * 2; break;
    case 3: src = w.x, bytes = B * L * D * 2; break;
    case 4: src = w.pe_cos, bytes = B * L * 64 * 2; break;
    case 5: src = w.pe_sin, bytes = B * L * 64 * 2; break;
    default: return fail("tap: unknown id");
  }
  FB_REQUIRE(out_bytes >= bytes, "tap: output buffer too small");
  FB_CHECK_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
