// fluxb200 internal interfaces between translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

typedef __nv_bfloat16 bf16;

namespace fb {

// ---- error plumbing (thread-local last error, returned through the C ABI) ----
void set_error(const std::string& msg);
const char* last_error();
int fail(const std::string& msg);  // sets error, returns -1

#define FB_CHECK_CUDA(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      return ::fb::fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " __FILE__ + ":" + \
                        std::to_string(__LINE__));                                                   \
    }                                                                                                \
  } while (0)

#define FB_REQUIRE(cond, msg)                        \
  do {                                               \
    if (!(cond)) return ::fb::fail(std::string(msg)); \
  } while (0)

// ---- tensor-map encoding (driver entry point resolved at run time; no link-time libcuda dependency) ----
int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                   uint32_t box_inner, uint32_t box_outer);
int encode_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);
int encode_tmap_4d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint64_t stride3_bytes, uint32_t b0, uint32_t b1,
                   uint32_t b2, uint32_t b3);

// un-swizzled 2-D map over raw bytes / f32 (packed quantised weights and their scales)
int encode_tmap_2d_raw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                       uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer);

int num_sms();

// Per-DEVICE one-time initialisation guard (cudaFuncSetAttribute and friends are per-device state: a process that
// drives several GPUs must run them once on each).  Usage: static DeviceOnce once; if (once.need()) { ...; once.done(); }
struct DeviceOnce {
  unsigned long long mask = 0;  // bit d: done on device d (accessed atomically)
  bool need() const;
  void done();
};
// bytes of launch accounting, adjusted by the step-graph replay (graph nodes are not launched through count_launch)
void count_launch_bulk(const unsigned long long* per_kind, int sign);
void snapshot_launches(unsigned long long* per_kind);
bool profiling_enabled();

// ---- launch accounting + optional per-kernel-class CUDA-event timing (bench.py roofline evidence) ----
enum KernelKind { KK_GEMM = 0, KK_ATTN, KK_LN_MOD, KK_QKNORM_ROPE, KK_DEQUANT, KK_GROUPNORM, KK_MISC, KK_COUNT };
void count_launch(int kind, int n = 1);
struct ProfScope {  // records an event pair around the launches issued in its lifetime when profiling is enabled
  ProfScope(int kind, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
  int kind;
  cudaStream_t stream;
  int slot;
};

// ---- tcgen05 GEMM  out[M,N] = epilogue(A[M,K] . W[N,K]^T) ----
enum { ACT_NONE = 0, ACT_GELU = 1 };
enum { BIAS_NONE = 0, BIAS_FUSED = 1, BIAS_AFTER_ROUND = 2 };

struct QuantB;

struct GemmDesc {
  // operands (bf16, row-major, K contiguous). lda/ldb in elements.
  const bf16* a = nullptr;
  int64_t lda = 0;
  const bf16* w = nullptr;
  int64_t ldb = 0;
  const QuantB* qb = nullptr;  // non-null: W is quantised and expanded inside the GEMM (w is ignored)
  int M = 0, N = 0, K = 0;
  // implicit-GEMM 3x3/1x1 convolution: A is an NHWC image [cN, cH, cW, cC]; M = cN*cH*cW output pixels,
  // K = taps*cC with W laid out [N, taps, cC]; ksize in {1,3}, padding = ksize/2, stride 1.
  int conv = 0, cN = 0, cH = 0, cW = 0, cC = 0, ksize = 0;
  // outputs: columns [0, n_split) -> out0 (ld0); columns [n_split, N) -> out1 at column (n - n_split + col_off1)
  bf16* out0 = nullptr;
  int64_t ld0 = 0;
  bf16* out1 = nullptr;
  int64_t ld1 = 0;
  int n_split = 0;  // 0 => everything to out0
  int col_off1 = 0;
  int act0 = ACT_NONE, act1 = ACT_NONE;
  const bf16* bias = nullptr;
  int bias_mode = BIAS_NONE;
  float alpha = 1.f;  // if != 1: out = bf16(bf16(acc) * alpha) (alpha must be bf16-representable)
  // out = res + gate * val (per-batch gate vector), applied when gate != nullptr
  const bf16* gate = nullptr;
  int64_t gate_bstride = 0;
  int rows_per_batch = 0;
  const bf16* res = nullptr;  // same ld as out0; may alias out0; may be used without gate (plain residual add)
  // denoising loop: gate points into a [step][...] table; the kernel adds (*step_ptr) * gate_step_stride elements
  const int* step_ptr = nullptr;
  int64_t gate_step_stride = 0;
  // Fused QK RMS-norm + RoPE + head-major relayout for columns [0, 3*qk_H*128) (the q|k|v projection): instead of
  // out0 the epilogue writes Q, K, V [B, H, L, 128] directly (SelfAttention::qkv + apply_rope, model.rs:86-95, 399-427).
  // Row r belongs to batch r / rows_per_batch and token qk_loff + r % rows_per_batch.
  int qkrope = 0;
  const bf16 *qk_wq = nullptr, *qk_wk = nullptr;  // RMS-norm weights [128]
  const uint2* qk_pe2 = nullptr;                  // [batch][pair 0..63][token]: {(cos, sin), (-sin, cos)} as 2 x bf16x2
  int64_t qk_pe_bstride = 0;                      // in uint2 elements
  bf16 *qk_Q = nullptr, *qk_K = nullptr, *qk_V = nullptr;
  int qk_H = 0, qk_L = 0, qk_loff = 0;
  float qk_eps = 1e-6f;
};
// Quantised B operand: the GEMM's producer warps expand packed weights straight into the swizzled smem tile
// (no bf16 copy of the weight ever exists in HBM).  Up to 4 members = the reference Linears fused along N.
enum { QB_NF4 = 1, QB_FP4 = 2, QB_Q4K = 3, QB_INT8 = 4 };
struct QuantMember {
  const uint8_t* packed = nullptr;  // nibbles / Q4_K blocks / int8 weights of this member, row-major [N_m, K]
  const float* absmax = nullptr;    // bnb 4-bit: f32 absmax per `blocksize` weights (nested absmax already expanded)
  const float* scb = nullptr;       // int8: per-row scale
  int row_begin = 0;                // first output column (row of W) of this member; multiple of 128
  int kind = 0;
  int blocksize = 64;
};
struct QuantB {
  QuantMember m[4];
  int count = 0;
};

// Launch helper for the hot kernels: optional 2-CTA cluster + programmatic dependent launch (the "pdl" flag).  Only
// kernels that execute griddepcontrol.wait before touching global memory may be launched through it.
template <class... KArgs, class... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             int cluster_x, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// runtime switches (A/B testing): "qkrope_fusion" (default 1), "gemm_pair" (default 1)
int get_flag(const char* name);
// One persistent launch over up to 4 problems (grouped): img + txt streams share the machine.
int launch_gemm(const GemmDesc* descs, int count, cudaStream_t stream);
int gemm_init_device();       // per-device kernel attributes; also called lazily by launch_gemm
int attention_init_device();  // same for launch_attention
int attention_num_variants();
int set_gemm_trace(long long* buf);  // debug: device buffer of 64*4 int64 written by scheduling unit 0, or nullptr

// ---- tcgen05 flash attention: q,k,v [B,H,L,128] bf16 -> out rows [B, L, H*128] split at L_split ----
struct AttnDesc {
  const bf16 *q = nullptr, *k = nullptr, *v = nullptr;
  int B = 0, H = 0, L = 0;
  float scale = 0.f;
  // token l < l_split goes to out_a[(b*l_split + l)*ld_a + h*128 ...], else out_b[(b*(L-l_split) + l-l_split)*ld_b ...]
  bf16* out_a = nullptr;
  int64_t ld_a = 0;
  bf16* out_b = nullptr;
  int64_t ld_b = 0;
  int l_split = 0;
  long long* trace = nullptr;  // debug: device buffer of 64*2*8 clock stamps written by CTA 0
};
int launch_attention(const AttnDesc& d, cudaStream_t stream);

}  // namespace fb
