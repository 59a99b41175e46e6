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
  size_t cache_off = 0;  // quantised: element offset of this linear's bf16 expansion inside the per-image weight cache
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
  int* step;             // device-side step counter of the denoising loop (indexes every per-step table)
  float *t_all, *g_all;  // [steps*B] timestep / guidance value of each (step, batch element) row
  bf16* dt_tab;          // [steps][in_channels] bf16(t_prev - t_curr), the "gate" of the fused Euler epilogue
  uint2* pe2;
  bf16 *pe_cos, *pe_sin, *temb, *gemb, *e1, *e2, *e3, *e4, *vec, *svec, *mod_all;
  bf16 *lat, *img, *txt, *x, *xm, *qkv, *Q, *K, *V, *attn_img, *attn_txt, *big, *txt_cache;
  bf16* wcache;  // denoising loop, quantised models: bf16 expansion of every quantised step weight, valid for one image
  size_t total = 0;
};

// One captured denoising step (CUDA graph) for a given workspace / geometry / kernel-flag combination.
struct StepGraph {
  void* ws_base = nullptr;
  int B = 0, l_img = 0, l_txt = 0;
  int steps = 0;  // the workspace layout (and with it every pointer baked into the graph) depends on the step count
  unsigned flags_sig = 0;
  cudaGraphExec_t exec = nullptr;
  unsigned long long launches[KK_COUNT] = {0};  // kernels per replay, by kind (launch accounting)
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
  // modulation projections: all 2*19 + 38 + 1 Linears of silu(vec_); their outputs for EVERY step of an image are
  // computed before the loop as M = steps*B row GEMMs into mod_all [steps*B][mod_row_elems]
  std::vector<const FusedLinear*> mods;
  std::vector<long long> mod_off;  // offset (elements, per row) of each projection's output inside a mod_all row
  long long mod_row_elems = 0;
  bf16* wscratch[2] = {nullptr, nullptr};  // dequantised-weight staging, double-buffered (quantised models only)
  int wscratch_cur = 0;
  // Expansion pipeline of the staged quantised path: the quantised Linears of one step in the order the step uses
  // them; the expansion of weight j+1 runs on a side stream while the GEMM that consumes weight j occupies the tensor
  // cores (the expansion is HBM-bound, the GEMM is not: they co-reside on the SMs)
  std::vector<const FusedLinear*> worder;
  size_t wnext = 0, wissued = 0;
  cudaStream_t side_stream = nullptr, cap_side_stream = nullptr;
  cudaEvent_t ev_fork[4] = {nullptr, nullptr, nullptr, nullptr}, ev_ready[2] = {nullptr, nullptr};
  size_t wscratch_elems = 0;
  size_t wcache_elems = 0;  // sum over worder of N*K
  bool any_quant = false;
  // step graph (denoise): captured on a private stream, replayed on the caller's
  cudaStream_t cap_stream = nullptr;
  std::vector<StepGraph> graphs;
  int last_used_graph = 0;
  std::string graph_note;
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

// Can this quantised linear run on the fused-dequant GEMM producer?  Mirrors launch_gemm's own checks so that a layer
// that does not qualify (x_embedder with K = 64, a blocksize that does not divide K, ...) falls back to the staging
// expansion instead of failing at step time.
static bool fused_dequant_ok(const FusedLinear& fl) {
  if (fl.N % 128 != 0 || fl.K % 64 != 0 || fl.members.empty() || fl.members.size() > 4) return false;
  for (const Member& mb : fl.members) {
    if (mb.row_off % 128 != 0) return false;
    if (mb.q == Q_NF4 || mb.q == Q_FP4) {
      if (mb.blocksize % 64 != 0 || fl.K % mb.blocksize != 0) return false;
      const int nabs = fl.K / mb.blocksize;
      if (nabs < 4 || nabs % 4 != 0) return false;
      if ((reinterpret_cast<uintptr_t>(mb.absmax) & 15) != 0) return false;
    } else if (mb.q == Q_Q4K) {
      if (fl.K % 256 != 0) return false;
    } else if (mb.q == Q_INT8) {
      if (mb.scb == nullptr) return false;
    }
    if ((reinterpret_cast<uintptr_t>(mb.packed) & 15) != 0) return false;
  }
  return true;
}

static int expand_into(const FusedLinear& fl, bf16* stage, cudaStream_t st);

// Weight operand for the GEMM: dense pointer, or expand the quantised members into a staging buffer (ONE launch for
// all members of the fused linear).  The two staging buffers alternate so that the expansion of the next layer never
// has to wait for the GEMM that is still reading the previous one.
static int weight_operand(fluxb200_model* m, const FusedLinear& fl, const bf16** w, cudaStream_t st) {
  if (!fl.quant) {
    *w = fl.w;
    return 0;
  }
  if (get_flag("dequant_mode") == 2 && fused_dequant_ok(fl)) {
    *w = nullptr;  // gemm_for() switches to the fused-dequant producer
    return 0;
  }
  bf16* stage = m->wscratch[m->wscratch_cur];
  m->wscratch_cur ^= 1;
  if (int rc = expand_into(fl, stage, st)) return rc;
  *w = stage;
  return 0;
}

static int expand_into(const FusedLinear& fl, bf16* stage, cudaStream_t st) {
  DequantBatch batch;
  batch.count = 0;
  for (auto& mb : fl.members) {
    FB_REQUIRE(batch.count < DequantBatch::MAX, "expand_into: too many fused members");
    DequantJob& j = batch.job[batch.count++];
    j.packed = mb.packed, j.absmax = mb.absmax, j.scb = mb.scb;
    j.out = stage + static_cast<size_t>(mb.row_off) * fl.K;
    j.n = static_cast<long long>(mb.N) * fl.K;
    j.col = fl.K, j.blocksize = mb.blocksize;
    j.kind = mb.q == Q_NF4 ? QB_NF4 : (mb.q == Q_FP4 ? QB_FP4 : (mb.q == Q_Q4K ? QB_Q4K : QB_INT8));
  }
  return launch_dequant_batch(batch, st);
}

// ---- the three ways a quantised weight reaches the tensor cores ("dequant_mode" flag) ----
//   0  per-image cache (default): fluxb200_model_denoise expands every quantised step weight ONCE per call into a bf16
//      cache carved from the caller's workspace (+ 2 B per quantised weight, 23.8 GB for a fully quantised FLUX.1-dev:
//      nothing on a 180 GB part) and all steps run the dense GEMMs on it; the resident model stays packed (6 GB)
//   1  staged per layer: every step expands each weight into an L2-sized staging buffer right before its GEMM, one
//      weight ahead on a side stream ("dequant_overlap")
//   2  fused: the GEMM's producer warps expand the packed tile in shared memory (no bf16 copy in HBM at all)
// Same bits in all three (tests).  A single fluxb200_model_forward has no image to amortise over: it uses 1 (or 2).
static bool weight_cache_enabled(const fluxb200_model* m) {
  return m->any_quant && !m->worder.empty() && get_flag("dequant_mode") == 0;
}

// ---- expansion pipeline (see fluxb200_model::worder) ----
static bool pipe_enabled(const fluxb200_model* m) {
  return m->any_quant && !m->worder.empty() && get_flag("dequant_overlap") != 0 && get_flag("dequant_mode") != 2 &&
         !profiling_enabled();
}
static void pipe_begin(fluxb200_model* m) { m->wnext = m->wissued = 0; }
// enqueue the expansion of worder[j] on `side`, ordered after everything launched on `main` so far (in particular
// after the GEMM that last read this staging buffer)
static int pipe_issue(fluxb200_model* m, size_t j, cudaStream_t main, cudaStream_t side) {
  cudaEvent_t fork = m->ev_fork[j & 3], ready = m->ev_ready[j & 1];
  FB_CHECK_CUDA(cudaEventRecord(fork, main));
  FB_CHECK_CUDA(cudaStreamWaitEvent(side, fork, 0));
  if (int rc = expand_into(*m->worder[j], m->wscratch[j & 1], side)) return rc;
  FB_CHECK_CUDA(cudaEventRecord(ready, side));
  m->wissued = j + 1;
  return 0;
}
// weight operand of `fl` for the next GEMM on `main`; also starts the expansion of the step's following weight
static int pipe_acquire(fluxb200_model* m, const FusedLinear& fl, const bf16** w, cudaStream_t main, cudaStream_t side) {
  const size_t j = m->wnext;
  FB_REQUIRE(j < m->worder.size() && m->worder[j] == &fl, "internal: quantised weight requested out of order");
  if (m->wissued <= j)
    if (int rc = pipe_issue(m, j, main, side)) return rc;
  FB_CHECK_CUDA(cudaStreamWaitEvent(main, m->ev_ready[j & 1], 0));
  *w = m->wscratch[j & 1];
  if (j + 1 < m->worder.size())
    if (int rc = pipe_issue(m, j + 1, main, side)) return rc;
  m->wnext = j + 1;
  return 0;
}

// Dense: W read by TMA.  Quantised: `w` is the staging buffer, or nullptr when the GEMM's producer warps expand the
// packed weights on the fly (fb::QuantB, "fused_dequant" = 1).
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

// `steps` = number of denoising steps whose per-step tables (t, dt, vec_, every modulation vector) live in the
// workspace: 1 for a single Flux::forward, n_timesteps - 1 for the denoising loop.
static Workspace carve(const fluxb200_model* m, void* base, int B, int l_img, int l_txt, int steps, bool cache = false) {
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
  const size_t rows = static_cast<size_t>(steps) * B;
  const size_t C = static_cast<size_t>(m->cfg.in_channels);
  w.step = static_cast<int*>(take(64));
  w.t_all = static_cast<float*>(take(rows * 4));
  w.g_all = static_cast<float*>(take(rows * 4));
  w.dt_tab = static_cast<bf16*>(take(static_cast<size_t>(steps) * C * 2));
  w.pe_cos = static_cast<bf16*>(take(Mx * 64 * 2));
  w.pe_sin = static_cast<bf16*>(take(Mx * 64 * 2));
  w.pe2 = static_cast<uint2*>(take(Mx * 64 * 8));
  w.temb = static_cast<bf16*>(take(rows * 256 * 2));
  w.gemb = static_cast<bf16*>(take(static_cast<size_t>(B) * 256 * 2));
  w.e1 = static_cast<bf16*>(take(rows * D * 2));
  w.e2 = static_cast<bf16*>(take(rows * D * 2));
  w.e3 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.e4 = static_cast<bf16*>(take(static_cast<size_t>(B) * D * 2));
  w.vec = static_cast<bf16*>(take(rows * D * 2));
  w.svec = static_cast<bf16*>(take(rows * D * 2));
  w.mod_all = static_cast<bf16*>(take(rows * m->mod_row_elems * 2));
  w.lat = static_cast<bf16*>(take(Mi * C * 2));
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
  w.txt_cache = static_cast<bf16*>(take(Mt * D * 2));
  w.wcache = cache ? static_cast<bf16*>(take(m->wcache_elems * 2)) : nullptr;
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

// rank-2 Linear on [rows, K] (MlpEmbedder / Modulation): matmul -> bf16, then a separate bf16 broadcast_add
// (unquantized/mod.rs:67; bnb the same, bitsandbytes/mod.rs:301-312); GGUF adds the bias to the f32 result.
static int rank2_bias_mode(const FusedLinear& fl) {
  return (fl.quant && !fl.members.empty() && fl.members[0].q == Q_Q4K) ? BIAS_FUSED : BIAS_AFTER_ROUND;
}
static int linear_rank2(fluxb200_model* m, const FusedLinear& fl, const bf16* x, int rows, bf16* out, int64_t ldo,
                        cudaStream_t st) {
  const bf16* w = nullptr;
  if (int rc = weight_operand(m, fl, &w, st)) return rc;
  GemmDesc d = gemm_for(fl, w, x, rows, out, ldo);
  d.bias_mode = rank2_bias_mode(fl);
  return launch_gemm(&d, 1, st);
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

void fluxb200_model_destroy(fluxb200_model* m);

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
  // per-device kernel attributes (opt-in shared memory) are set here, not lazily inside a launch: the first launch of
  // a denoising step may happen under stream capture
  if (int rc = gemm_init_device()) return rc;
  if (int rc = attention_init_device()) return rc;
  auto* m = new fluxb200_model();
  m->cfg = *cfg;
  cudaError_t e = cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->cap_side_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&m->ev_fork[i], cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&m->ev_ready[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    fluxb200_model_destroy(m);
    return fail(std::string("model_create: stream / event creation failed: ") + cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

void fluxb200_model_destroy(fluxb200_model* m) {
  if (!m) return;
  for (auto& kv : m->raw) cudaFree(kv.second.dev);
  for (void* p : m->owned) cudaFree(p);
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  if (m->cap_side_stream) cudaStreamDestroy(m->cap_side_stream);
  if (m->side_stream) cudaStreamDestroy(m->side_stream);
  for (auto ev : m->ev_fork)
    if (ev) cudaEventDestroy(ev);
  for (auto ev : m->ev_ready)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i)
    if (m->wscratch[i]) cudaFree(m->wscratch[i]);
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
  // build_linear reads quant-state tensors back with blocking copies: load_weight's async copies (possibly on a
  // non-blocking stream) must have landed first
  FB_CHECK_CUDA(cudaStreamSynchronize(st));
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

  // modulation outputs: one row of `mod_row_elems` bf16 per (step, batch element); order = double (img, txt) ..., single, final
  m->mods.clear();
  for (auto& b : m->dbl) {
    m->mods.push_back(&b.img_mod);
    m->mods.push_back(&b.txt_mod);
  }
  for (auto& b : m->sgl) m->mods.push_back(&b.mod);
  m->mods.push_back(&m->final_mod);
  m->mod_off.clear();
  long long off = 0;
  for (auto* fl : m->mods) {
    m->mod_off.push_back(off);
    off += fl->N;
  }
  m->mod_row_elems = off;
  if (m->any_quant) {
    for (int i = 0; i < 2; ++i) FB_CHECK_CUDA(cudaMalloc(&m->wscratch[i], m->wscratch_elems * 2));
  }
  // the quantised Linears of one step, in the order step_core consumes them (pipe_acquire checks it)
  m->worder.clear();
  auto use = [&](const FusedLinear& fl) {
    if (fl.quant) m->worder.push_back(&fl);
  };
  use(m->img_in);
  for (auto& b : m->dbl) {
    use(b.img_qkv), use(b.txt_qkv), use(b.img_proj), use(b.txt_proj);
    use(b.img_mlp1), use(b.txt_mlp1), use(b.img_mlp2), use(b.txt_mlp2);
  }
  for (auto& b : m->sgl) use(b.lin1), use(b.lin2);
  use(m->final_proj);
  m->wcache_elems = 0;
  for (const FusedLinear* fl : m->worder) {
    const_cast<FusedLinear*>(fl)->cache_off = m->wcache_elems;
    m->wcache_elems += (static_cast<size_t>(fl->N) * fl->K + 511) / 512 * 512;  // keep every weight 1 KB aligned
  }
  m->finalized = true;
  return 0;
#undef TRY
}

int fluxb200_model_workspace_size(const fluxb200_model* m, int32_t batch, int32_t l_img, int32_t l_txt,
                                  uint64_t* bytes) {
  FB_REQUIRE(m && bytes && m->finalized, "workspace_size: model not finalized");
  FB_REQUIRE(batch >= 1 && batch <= 8 && l_img > 0 && l_txt > 0, "workspace_size: bad geometry (batch 1..8)");
  *bytes = carve(m, nullptr, batch, l_img, l_txt, 1).total + 1024;
  return 0;
}

int fluxb200_model_denoise_workspace_size(const fluxb200_model* m, int32_t batch, int32_t l_img, int32_t l_txt,
                                          int32_t n_timesteps, uint64_t* bytes) {
  FB_REQUIRE(m && bytes && m->finalized, "denoise_workspace_size: model not finalized");
  FB_REQUIRE(batch >= 1 && batch <= 8 && l_img > 0 && l_txt > 0, "denoise_workspace_size: bad geometry (batch 1..8)");
  FB_REQUIRE(n_timesteps >= 2 && n_timesteps <= MAX_STEPS, "denoise_workspace_size: 2..1024 timesteps");
  *bytes = carve(m, nullptr, batch, l_img, l_txt, n_timesteps - 1, weight_cache_enabled(m)).total + 1024;
  return 0;
}

}  // extern "C"

namespace fb {

struct StepIO {
  const bf16* img_ids;  // [B, l_img, 3]
  const bf16* txt_in;   // [B, l_txt, joint]
  const bf16* txt_ids;  // [B, l_txt, 3]
  const bf16* y;        // [B, pooled]
};

#define TRY(x)          \
  do {                  \
    int _rc = (x);      \
    if (_rc) return _rc; \
  } while (0)

// Step-invariant part 1: RoPE table (model.rs:807-810)
static int prepare_invariants(fluxb200_model* m, const Workspace& w, const StepIO& io, int B, int l_img, int l_txt,
                              cudaStream_t st) {
  const int L = l_img + l_txt;
  const int n1 = B * l_txt * 64, n2 = B * l_img * 64;
  count_launch(KK_MISC, 2);
  pe_table_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(io.txt_ids, l_txt, B, L, 0, w.pe_cos, w.pe_sin, w.pe2);
  pe_table_kernel<<<(n2 + 255) / 256, 256, 0, st>>>(io.img_ids, l_img, B, L, l_txt, w.pe_cos, w.pe_sin, w.pe2);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Step-invariant part 2: txt = txt_in(txt) (model.rs:811) -> w.txt_cache; the first double block reads it from there.
static int project_txt(fluxb200_model* m, const Workspace& w, const StepIO& io, int B, int l_txt, cudaStream_t st) {
  const bf16* wt = nullptr;
  TRY(weight_operand(m, m->txt_in, &wt, st));
  GemmDesc d = gemm_for(m->txt_in, wt, io.txt_in, B * l_txt, w.txt_cache, D);
  return launch_gemm(&d, 1, st);
}

// vec_ (model.rs:813-820) and EVERY AdaLN modulation projection (model.rs:244-299, 694-698) for `rows` = steps*B
// (step, batch element) pairs at once.  vec_ depends only on (t, guidance, y) and all timesteps of an image are known
// before the loop, so the 6.5 GB of modulation weights are streamed ONCE per image (M = steps*B rows on the tcgen05
// GEMM) instead of once per step; every row keeps the reference's rounding points (the Linears are row-independent).
//   t_all [rows] f32, g_all [>= B] f32 or nullptr (guidance does not depend on the step), y [B, pooled]
static int compute_modulations(fluxb200_model* m, const Workspace& w, const float* t_all, const float* g_all,
                               const bf16* y, int rows, int B, cudaStream_t st) {
  const auto& c = m->cfg;
  TRY(launch_timestep_embedding(t_all, w.temb, rows, 256, st));
  TRY(linear_rank2(m, m->time1, w.temb, rows, w.e1, D, st));
  TRY(launch_silu(w.e1, w.e1, static_cast<long long>(rows) * D, st));
  TRY(linear_rank2(m, m->time2, w.e1, rows, w.e2, D, st));
  const bf16* gvec = nullptr;
  if (c.guidance_embeds && g_all) {
    TRY(launch_timestep_embedding(g_all, w.gemb, B, 256, st));
    TRY(linear_rank2(m, m->guid1, w.gemb, B, w.e1, D, st));
    TRY(launch_silu(w.e1, w.e1, static_cast<long long>(B) * D, st));
    TRY(linear_rank2(m, m->guid2, w.e1, B, w.e3, D, st));
    gvec = w.e3;
  }
  TRY(linear_rank2(m, m->vecin1, y, B, w.e1, D, st));
  TRY(launch_silu(w.e1, w.e1, static_cast<long long>(B) * D, st));
  TRY(linear_rank2(m, m->vecin2, w.e1, B, w.e4, D, st));
  TRY(launch_vec_combine(w.e2, gvec, w.e4, w.vec, rows, B, D, st));
  TRY(launch_silu(w.vec, w.svec, static_cast<long long>(rows) * D, st));
  // dense projections share launches (grouped GEMM, 4 problems each); quantised ones go one at a time through the staging buffer
  const size_t n = m->mods.size();
  for (size_t i = 0; i < n;) {
    GemmDesc g[4];
    int cnt = 0;
    while (i < n && cnt < 4) {
      const FusedLinear& fl = *m->mods[i];
      if (fl.quant && cnt > 0) break;
      const bf16* wp = nullptr;
      TRY(weight_operand(m, fl, &wp, st));
      g[cnt] = gemm_for(fl, wp, w.svec, rows, w.mod_all + m->mod_off[i], m->mod_row_elems);
      g[cnt].bias_mode = rank2_bias_mode(fl);
      ++cnt, ++i;
      if (fl.quant) break;
    }
    TRY(launch_gemm(g, cnt, st));
  }
  return 0;
}

// One Flux::forward (model.rs:790-833) given the hoisted pieces: RoPE table, txt_in(txt) in w.txt_cache and the
// modulation vectors in w.mod_all.
//   step_ptr == nullptr : single forward; modulation row block 0; prediction written to `pred_out`
//   step_ptr != nullptr : denoising step; modulation row block *step_ptr; the Euler update
//                         img += bf16(pred * dt) (pipelines/sampling.rs:43) is the final projection's epilogue, in place
//                         on `lat` (out = res + gate * val with gate = dt_tab[step], two roundings)
static int step_core(fluxb200_model* m, const Workspace& w, const bf16* img_in, bf16* pred_out, bf16* lat,
                     const int* step_ptr, int B, int l_img, int l_txt, cudaStream_t st, cudaStream_t side) {
  const auto& c = m->cfg;
  // quantised weights: the per-image bf16 cache when the caller's workspace carries one (denoising loop); otherwise
  // staged expansion, software-pipelined one weight ahead on the side stream (or in-order on `st`), or fused
  const bf16* const cache = w.wcache;
  const bool piped = cache == nullptr && pipe_enabled(m);
  pipe_begin(m);
  auto W = [&](const FusedLinear& fl, const bf16** wp) -> int {
    if (cache && fl.quant) {
      *wp = cache + fl.cache_off;
      return 0;
    }
    if (piped && fl.quant) return pipe_acquire(m, fl, wp, st, side);
    return weight_operand(m, fl, wp, st);
  };
  const int L = l_img + l_txt;
  const int Mi = B * l_img, Mt = B * l_txt, Mx = B * L;
  const int H = c.num_attention_heads;
  const float eps = 1e-6f;
  const float scale = 1.0f / sqrtf(static_cast<float>(HEAD_DIM));
  const long long PE_BS = static_cast<long long>(L) * 64;
  const bool fuse_qk = get_flag("qkrope_fusion") != 0;
  const long long mstride = m->mod_row_elems;                   // batch stride inside one step's block of rows
  const long long sstride = static_cast<long long>(B) * mstride;  // step stride
  auto modp = [&](int job, int chunk) { return w.mod_all + m->mod_off[job] + static_cast<long long>(chunk) * D; };
  auto ln_mod = [&](const bf16* x, int in_bstride_rows, int in_row_off, int rows_per_batch, int job, int shift_chunk,
                    int scale_chunk, bf16* out) {
    return launch_ln_modulate(x, in_bstride_rows, in_row_off, rows_per_batch, B, modp(job, shift_chunk),
                              modp(job, scale_chunk), mstride, out, D, eps, st, step_ptr, sstride);
  };
  // img and txt stream of a double block in one launch: output rows [img | txt] = w.xm
  auto ln_mod2 = [&](const bf16* ximg, const bf16* xtxt, int ji, int jt, int shift_chunk, int scale_chunk) {
    LnInput in[2] = {{ximg, modp(ji, shift_chunk), modp(ji, scale_chunk), l_img, 0, l_img, B},
                     {xtxt, modp(jt, shift_chunk), modp(jt, scale_chunk), l_txt, 0, l_txt, B}};
    return launch_ln_modulate2(in, 2, mstride, w.xm, D, eps, st, step_ptr, sstride);
  };
  auto gated = [&](GemmDesc& g, int job, int gate_chunk, int rows_per_batch, const bf16* res) {
    g.gate = modp(job, gate_chunk), g.gate_bstride = mstride, g.rows_per_batch = rows_per_batch, g.res = res;
    g.step_ptr = step_ptr, g.gate_step_stride = sstride;
  };

  // ---- img_in (model.rs:812); txt_in(txt) was hoisted ----
  {
    const bf16* wi = nullptr;
    TRY(W(m->img_in, &wi));
    GemmDesc d = gemm_for(m->img_in, wi, img_in, Mi, w.img, D);
    TRY(launch_gemm(&d, 1, st));
  }
  const bf16* txt_cur = w.txt_cache;  // the txt stream is read from the hoisted projection until its first update

  // ---- double-stream blocks (model.rs:523-565) ----
  for (int i = 0; i < c.num_layers; ++i) {
    DoubleBlock& b = m->dbl[i];
    const int ji = 2 * i, jt = 2 * i + 1;  // modulation jobs: chunks = shift1, scale1, gate1, shift2, scale2, gate2
    bf16* xm_img = w.xm;
    bf16* xm_txt = w.xm + static_cast<size_t>(Mi) * D;
    bf16* qkv_img = w.qkv;
    bf16* qkv_txt = w.qkv + static_cast<size_t>(Mi) * 3 * D;
    TRY(ln_mod2(w.img, txt_cur, ji, jt, 0, 1));
    // img and txt problems share one launch when both weights are dense; a quantised weight goes through the (reused)
    // staging buffer, so those problems are launched one by one right after their expansion
    auto two = [&](FusedLinear& fi, FusedLinear& ft, GemmDesc& gi, GemmDesc& gt) -> int {
      if ((!fi.quant && !ft.quant) || cache) {  // both operands are stable bf16 tensors: one grouped launch
        GemmDesc g[2] = {gi, gt};
        if (int r = W(fi, &g[0].w)) return r;
        if (int r = W(ft, &g[1].w)) return r;
        g[0].qb = g[1].qb = nullptr;
        return launch_gemm(g, 2, st);
      }
      for (int s = 0; s < 2; ++s) {
        FusedLinear& f = s == 0 ? fi : ft;
        GemmDesc& g = s == 0 ? gi : gt;
        const bf16* wq = nullptr;
        if (int r = W(f, &wq)) return r;
        g.w = wq;
        g.qb = (f.quant && wq == nullptr) ? &f.qb : nullptr;
        if (int r = launch_gemm(&g, 1, st)) return r;
      }
      return 0;
    };
    {
      GemmDesc gi = gemm_for(b.img_qkv, nullptr, xm_img, Mi, qkv_img, 3 * D);
      GemmDesc gt = gemm_for(b.txt_qkv, nullptr, xm_txt, Mt, qkv_txt, 3 * D);
      if (fuse_qk) {
        attach_qkrope(gi, w, b.img_nq, b.img_nk, H, L, l_txt, l_img, eps);
        attach_qkrope(gt, w, b.txt_nq, b.txt_nk, H, L, 0, l_txt, eps);
      }
      TRY(two(b.img_qkv, b.txt_qkv, gi, gt));
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
    {  // img/txt += gate1 * proj(attn)
      GemmDesc gi = gemm_for(b.img_proj, nullptr, w.attn_img, Mi, w.img, D);
      GemmDesc gt = gemm_for(b.txt_proj, nullptr, w.attn_txt, Mt, w.txt, D);
      gated(gi, ji, 2, l_img, w.img);
      gated(gt, jt, 2, l_txt, txt_cur);
      TRY(two(b.img_proj, b.txt_proj, gi, gt));
      txt_cur = w.txt;
    }
    // MLP: x += gate2 * lin2(gelu(lin1(modulate2(LN(x)))))
    TRY(ln_mod2(w.img, w.txt, ji, jt, 3, 4));
    bf16* h_img = w.big;
    bf16* h_txt = w.big + static_cast<size_t>(Mi) * MLP_D;
    {
      GemmDesc gi = gemm_for(b.img_mlp1, nullptr, xm_img, Mi, h_img, MLP_D);
      GemmDesc gt = gemm_for(b.txt_mlp1, nullptr, xm_txt, Mt, h_txt, MLP_D);
      gi.act0 = gt.act0 = ACT_GELU;
      TRY(two(b.img_mlp1, b.txt_mlp1, gi, gt));
    }
    {
      GemmDesc gi = gemm_for(b.img_mlp2, nullptr, h_img, Mi, w.img, D);
      GemmDesc gt = gemm_for(b.txt_mlp2, nullptr, h_txt, Mt, w.txt, D);
      gated(gi, ji, 5, l_img, w.img);
      gated(gt, jt, 5, l_txt, w.txt);
      TRY(two(b.img_mlp2, b.txt_mlp2, gi, gt));
    }
  }

  // ---- cat(txt, img) (model.rs:827) ----
  TRY(launch_copy_rows(w.x, static_cast<long long>(L) * D * 2, txt_cur, static_cast<long long>(l_txt) * D * 2,
                       static_cast<long long>(l_txt) * D * 2, B, st));
  TRY(launch_copy_rows(w.x + static_cast<size_t>(l_txt) * D, static_cast<long long>(L) * D * 2, w.img,
                       static_cast<long long>(l_img) * D * 2, static_cast<long long>(l_img) * D * 2, B, st));
  // ---- single-stream blocks (model.rs:638-662) ----
  const int CAT = D + MLP_D;
  for (int i = 0; i < c.num_single_layers; ++i) {
    SingleBlock& b = m->sgl[i];
    const int j = 2 * c.num_layers + i;  // chunks = shift, scale, gate
    TRY(ln_mod(w.x, L, 0, L, j, 0, 1, w.xm));
    {
      const bf16* w1 = nullptr;
      TRY(W(b.lin1, &w1));
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
      TRY(W(b.lin2, &w2));
      GemmDesc g = gemm_for(b.lin2, w2, w.big, Mx, w.x, D);
      gated(g, j, 2, L, w.x);
      TRY(launch_gemm(&g, 1, st));
    }
  }
  // ---- final layer on the img rows (model.rs:831-832, 694-705): chunks = scale, shift ----
  {
    const int jf = 2 * c.num_layers + c.num_single_layers;
    TRY(ln_mod(w.x, L, l_txt, l_img, jf, 1, 0, w.xm));
    const bf16* wf = nullptr;
    TRY(W(m->final_proj, &wf));
    GemmDesc g = gemm_for(m->final_proj, wf, w.xm, Mi, step_ptr ? lat : pred_out, c.in_channels);
    if (step_ptr) {  // fused Euler update
      g.gate = w.dt_tab, g.gate_bstride = 0, g.rows_per_batch = l_img, g.res = lat;
      g.step_ptr = step_ptr, g.gate_step_stride = c.in_channels;
    }
    TRY(launch_gemm(&g, 1, st));
  }
  if (piped) FB_REQUIRE(m->wnext == m->worder.size() && m->wissued == m->wnext, "internal: expansion pipeline out of step");
  return 0;
}

static int check_ws(fluxb200_model* m, int B, int l_img, int l_txt, int steps, bool cache, void* ws, uint64_t ws_bytes,
                    Workspace* out) {
  FB_REQUIRE(m && m->finalized, "model not finalized");
  FB_REQUIRE(B >= 1 && B <= 8, "batch must be in 1..8 per call");
  FB_REQUIRE(l_img > 0 && l_txt > 0, "empty sequence");
  FB_REQUIRE(ws != nullptr, "null workspace");
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  const size_t slack = base - static_cast<uint8_t*>(ws);
  Workspace w = carve(m, base, B, l_img, l_txt, steps, cache);
  FB_REQUIRE(w.total + slack <= ws_bytes, "workspace too small: need " + std::to_string(w.total + 1024) + " bytes");
  *out = w;
  return 0;
}

static unsigned flags_signature() {
  unsigned s = 0;
  s = s * 2 + (get_flag("qkrope_fusion") & 1);
  s = s * 2 + (get_flag("pdl") & 1);
  s = s * 4 + (get_flag("dequant_mode") & 3);
  s = s * 2 + (get_flag("gemm_pair") & 1);
  s = s * 2 + (get_flag("gemm_cl4") & 1);
  s = s * 2 + (get_flag("dequant_overlap") & 1);
  s = s * 16 + (get_flag("gemm_big") & 15);
  s = s * 64 + (get_flag("attn_variant") & 63);
  return s;
}

// The denoising step as a CUDA graph: captured once per (workspace, geometry, kernel flags) on the model's private
// stream — every tensor map, launch attribute (clusters, programmatic dependent launch edges) and kernel parameter
// block is encoded at capture time — and replayed once per step; what changes from step to step is read by the
// kernels through the device-side step counter.  Returns nullptr (with m->graph_note set) when capture is not possible.
static StepGraph* step_graph_for(fluxb200_model* m, const Workspace& w, void* ws_base, int B, int l_img, int l_txt,
                                 int steps) {
  const unsigned sig = flags_signature();
  for (auto& g : m->graphs)
    if (g.ws_base == ws_base && g.B == B && g.l_img == l_img && g.l_txt == l_txt && g.steps == steps &&
        g.flags_sig == sig)
      return &g;
  if (m->graphs.size() >= 8) {  // bounded cache: drop the oldest capture
    if (m->graphs.front().exec) cudaGraphExecDestroy(m->graphs.front().exec);
    m->graphs.erase(m->graphs.begin());
  }
  StepGraph sg;
  sg.ws_base = ws_base, sg.B = B, sg.l_img = l_img, sg.l_txt = l_txt, sg.steps = steps, sg.flags_sig = sig;
  unsigned long long before[KK_COUNT], after[KK_COUNT];
  snapshot_launches(before);
  const int scratch_cur = m->wscratch_cur;
  cudaError_t e = cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    m->graph_note = std::string("cudaStreamBeginCapture failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return nullptr;
  }
  int rc = step_core(m, w, w.lat, nullptr, w.lat, w.step, B, l_img, l_txt, m->cap_stream, m->cap_side_stream);
  if (rc == 0) rc = launch_step_advance(w.step, m->cap_stream);
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(m->cap_stream, &graph);
  snapshot_launches(after);
  for (int k = 0; k < KK_COUNT; ++k) sg.launches[k] = after[k] - before[k];
  count_launch_bulk(sg.launches, -1);  // nothing was launched while capturing
  m->wscratch_cur = scratch_cur;       // every replay starts from the staging buffer the capture started from
  if (rc != 0 || e != cudaSuccess || graph == nullptr) {
    m->graph_note = rc != 0 ? std::string("capture failed: ") + last_error()
                            : std::string("cudaStreamEndCapture failed: ") + cudaGetErrorString(e);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return nullptr;
  }
  e = cudaGraphInstantiate(&sg.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    m->graph_note = std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return nullptr;
  }
  m->graphs.push_back(sg);
  return &m->graphs.back();
}

#undef TRY

}  // namespace fb

extern "C" {

int fluxb200_model_forward(fluxb200_model* m, const void* img, const void* img_ids, const void* txt,
                           const void* txt_ids, const void* timesteps, const void* y, const void* guidance,
                           void* out, int32_t batch, int32_t l_img, int32_t l_txt, void* workspace,
                           uint64_t workspace_bytes, fluxb200_stream_t stream) {
  FB_REQUIRE(img && img_ids && txt && txt_ids && timesteps && y && out, "forward: null tensor");
  Workspace w;
  if (int rc = check_ws(m, batch, l_img, l_txt, 1, false, workspace, workspace_bytes, &w)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepIO io{static_cast<const bf16*>(img_ids), static_cast<const bf16*>(txt), static_cast<const bf16*>(txt_ids),
            static_cast<const bf16*>(y)};
  if (int rc = prepare_invariants(m, w, io, batch, l_img, l_txt, st)) return rc;
  if (int rc = project_txt(m, w, io, batch, l_txt, st)) return rc;
  if (int rc = compute_modulations(m, w, static_cast<const float*>(timesteps), static_cast<const float*>(guidance),
                                   io.y, batch, batch, st))
    return rc;
  if (int rc = step_core(m, w, static_cast<const bf16*>(img), static_cast<bf16*>(out), nullptr, nullptr, batch, l_img,
                         l_txt, st, m->side_stream))
    return rc;
  m->last_B = batch, m->last_limg = l_img, m->last_ltxt = l_txt, m->last_ws = w;
  return 0;
}

int fluxb200_model_denoise(fluxb200_model* m, void* img, const void* img_ids, const void* txt, const void* txt_ids,
                           const void* y, float guidance_scale, const double* timesteps, int32_t n_timesteps,
                           int32_t batch, int32_t l_img, int32_t l_txt, void* workspace, uint64_t workspace_bytes,
                           fluxb200_stream_t stream) {
  FB_REQUIRE(img && img_ids && txt && txt_ids && y && timesteps, "denoise: null tensor");
  FB_REQUIRE(n_timesteps >= 2, "denoise: need at least two timesteps");
  FB_REQUIRE(n_timesteps <= MAX_STEPS, "denoise: at most 1024 timesteps");
  const int steps = n_timesteps - 1;
  Workspace w;
  if (int rc = check_ws(m, batch, l_img, l_txt, steps, weight_cache_enabled(m), workspace, workspace_bytes, &w)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepIO io{static_cast<const bf16*>(img_ids), static_cast<const bf16*>(txt), static_cast<const bf16*>(txt_ids),
            static_cast<const bf16*>(y)};
  const int C = m->cfg.in_channels;
  const long long lat_bytes = static_cast<long long>(batch) * l_img * C * 2;
  FB_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, "denoise: img must be 16-byte aligned");
  // ---- per-image prologue: everything that does not depend on the evolving latent ----
  // the latent lives in the workspace during the loop (the step graph is bound to workspace addresses only)
  if (int rc = launch_copy_rows(w.lat, 0, img, 0, lat_bytes, 1, st)) return rc;
  if (int rc = prepare_invariants(m, w, io, batch, l_img, l_txt, st)) return rc;
  if (int rc = project_txt(m, w, io, batch, l_txt, st)) return rc;
  // t_vec = full(1) * t_curr (f32); guidance = full(guidance_scale) (pipelines/flux/mod.rs:300-304, sampling.rs:42);
  // dt = t_prev - t_curr.  The scalars travel as kernel parameters: no staging buffer, no synchronisation.
  for (int s0 = 0; s0 < steps; s0 += StepScalars::N) {
    StepScalars sv;
    const int n = std::min(StepScalars::N, steps - s0);
    for (int i = 0; i < n; ++i) {
      sv.t[i] = static_cast<float>(1.0f * timesteps[s0 + i]);
      sv.dt[i] = static_cast<float>(timesteps[s0 + i + 1] - timesteps[s0 + i]);
    }
    if (int rc = launch_step_scalars(sv, s0, n, batch, C, guidance_scale, w.t_all, w.g_all, w.dt_tab, st)) return rc;
  }
  FB_CHECK_CUDA(cudaMemsetAsync(w.step, 0, 64, st));
  if (int rc = compute_modulations(m, w, w.t_all, m->cfg.guidance_embeds ? w.g_all : nullptr, io.y, steps * batch,
                                   batch, st))
    return rc;
  // quantised model: expand every step weight once for the whole image (see weight_cache_enabled)
  if (w.wcache)
    for (const FusedLinear* fl : m->worder)
      if (int rc = expand_into(*fl, w.wcache + fl->cache_off, st)) return rc;
  // ---- the loop: one graph replay per step (or the same kernels launched one by one) ----
  StepGraph* sg = nullptr;
  m->last_used_graph = 0;
  if (get_flag("step_graph") && !profiling_enabled()) {
    uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
    sg = step_graph_for(m, w, base, batch, l_img, l_txt, steps);
  }
  if (sg) {
    for (int s = 0; s < steps; ++s) {
      FB_CHECK_CUDA(cudaGraphLaunch(sg->exec, st));
      count_launch_bulk(sg->launches, +1);
    }
    m->last_used_graph = 1;
  } else {
    for (int s = 0; s < steps; ++s) {
      if (int rc = step_core(m, w, w.lat, nullptr, w.lat, w.step, batch, l_img, l_txt, st, m->side_stream)) return rc;
      if (int rc = launch_step_advance(w.step, st)) return rc;
    }
  }
  if (int rc = launch_copy_rows(img, 0, w.lat, 0, lat_bytes, 1, st)) return rc;
  m->last_B = batch, m->last_limg = l_img, m->last_ltxt = l_txt, m->last_ws = w;
  return 0;
}

int fluxb200_model_denoise_info(const fluxb200_model* m, int32_t* used_graph, const char** note) {
  FB_REQUIRE(m, "denoise_info: null model");
  if (used_graph) *used_graph = m->last_used_graph;
  if (note) *note = m->graph_note.c_str();
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
    case 2: src = m->cfg.num_layers > 0 ? w.txt : w.txt_cache, bytes = B * lt * D * 2; break;
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
