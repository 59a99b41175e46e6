/* fluxb200 — C ABI of the B200-native FLUX denoising hot path (drop-in for diffusion-rs).
 *
 * Every entry point is plain C: raw device/host pointers, sizes and an explicit CUDA stream.
 * Conventions (mirroring the reference's own FFI, diffusion_rs_backend/src/bitsandbytes/ffi.rs:5-114,
 * and the ownership rules of its CustomOp::cuda_fwd call sites, bitsandbytes/op.rs:204-228):
 *   - the caller owns every activation / output / workspace buffer; the library never allocates them;
 *   - the library owns only immutable weight copies created by fluxb200_load_weight, freed by _destroy;
 *   - all work is enqueued on the caller's stream, no hidden device synchronisation;
 *   - functions return 0 on success, non-zero on failure; fluxb200_last_error() gives the reason
 *     (thread-local string). Nothing throws or aborts across the ABI.
 *   - activations are bf16, row-major; Linear weights are [out, in] row-major (torch layout).
 * There is NO CPU fallback: without an sm_100a device every compute entry point fails with an error.
 */
#ifndef FLUXB200_H
#define FLUXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fluxb200_stream_t; /* cudaStream_t / CUstream */

const char* fluxb200_last_error(void);
int fluxb200_version(void);

/* ------------------------------------------------------------------------------------------------
 * Operator-level entry points (each slots behind one QuantMethod / CustomOp of the reference)
 * ---------------------------------------------------------------------------------------------- */

/* bias_mode */
#define FLUXB200_BIAS_NONE 0
#define FLUXB200_BIAS_FUSED 1       /* bias added to the fp32 accumulator (cuBLASLt path, unquantized/mod.rs:52-66) */
#define FLUXB200_BIAS_AFTER_ROUND 2 /* matmul rounded to bf16, then bf16 broadcast_add (unquantized/mod.rs:67, bnb) */
/* act */
#define FLUXB200_ACT_NONE 0
#define FLUXB200_ACT_GELU_TANH 1 /* candle Gelu, bf16 step-wise rounding (core/op.rs:539-578) */

/* out[M,N] = epilogue(a[M,K] . w[N,K]^T).  Replaces QuantMethod::forward of UnquantLinear
 * (diffusion_rs_backend/src/unquantized/mod.rs:34-77).  Optional fused epilogue:
 *   v = bf16(acc (+bias));  v = act(v);  if gate: v = bf16(gate[b,:] * v);  if res: v = bf16(res + v)
 * gate is indexed [row / rows_per_batch][col] with batch stride gate_bstride (elements). */
int fluxb200_linear(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out, int64_t ldo,
                    int32_t M, int32_t N, int32_t K, int32_t bias_mode, int32_t act, const void* gate,
                    int64_t gate_bstride, int32_t rows_per_batch, const void* res, float alpha,
                    fluxb200_stream_t stream);

/* Same as fluxb200_linear with a QUANTISED weight that is expanded inside the GEMM's operand producer (never as a
 * bf16 tensor in HBM).  kind: 1 = bnb NF4, 2 = bnb FP4 (packed nibbles [N*K/2], aux = f32 absmax per `blocksize`
 * weights), 3 = GGUF Q4_K (144-byte super-blocks, aux = NULL), 4 = LLM.int8 (int8 [N,K], aux = f32 SCB[N]).
 * Replaces BnbLinear::forward / GgufMatMul::forward_via_half (bitsandbytes/mod.rs:301-312, gguf/mod.rs:42-49):
 * W = dequant(packed) rounded to bf16, then the bf16 GEMM.  Needs N % 128 == 0 and K % 64 == 0 (Q4_K: K % 256 == 0). */
int fluxb200_linear_quant(const void* a, int64_t lda, const void* packed, const void* aux, int32_t kind,
                          int32_t blocksize, const void* bias, void* out, int64_t ldo, int32_t M, int32_t N, int32_t K,
                          int32_t bias_mode, int32_t act, fluxb200_stream_t stream);

/* Joint attention.  q,k,v: bf16 [B,H,L,head_dim] contiguous (the (bs, qhead, seq, hidden) operands of
 * diffusion_rs_backend::ops::sdpa, ops.rs:247-262, with k/v heads == q heads); out: bf16 [B,L,H*head_dim], i.e. the
 * reference's (bs, qhead, seq, v_hidden) result ALREADY passed through `.transpose(1,2).flatten_from(2)`, which is what
 * its only FLUX call site does next (model.rs:97-102) - a shim that needs the untransposed tensor transposes back.
 * Preconditions, checked: head_dim == 128, softcapping == 1.0 (off).  Anything else returns non-zero with a message
 * and enqueues nothing, so that the caller can fall through to the stock ops::sdpa. */
int fluxb200_sdpa(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t L,
                  int32_t head_dim, float scale, float softcapping, fluxb200_stream_t stream);
/* Number of run-time selectable builds of the attention kernel ("attn_variant" flag values 0 .. n-1). */
int fluxb200_attn_variants(void);

/* Debug/profiling twin of fluxb200_sdpa: additionally records clock64 stamps of CTA 0's pipeline stages into `trace`
 * (device buffer of 64*2*8 int64: [kv block][query tile][stage]); used by scripts/attn_trace.py only. */
int fluxb200_debug_sdpa_trace(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t L,
                              float scale, void* trace, fluxb200_stream_t stream);

/* Debug/profiling: while `trace` (device buffer of 64*4 int64) is non-NULL, every fluxb200 GEMM launch records, for the
 * first 64 tiles of scheduling unit 0, the MMA thread's clock64 waits {tile start, wait for a free accumulator, wait for
 * TMA data, tile total}; used by scripts/gemm_trace.py only.  Needs a library built with FLUXB200_GEMM_TRACE=1 (the
 * instrumentation is compiled out of production builds); otherwise a non-NULL `trace` returns an error. */
int fluxb200_debug_gemm_trace(void* trace);

/* out = modulate(LayerNorm(x)) = (LN(x) * (1 + scale[b])) + shift[b]; x,out bf16 [B*rows, 3072].
 * Replaces nn::LayerNorm::forward + ModulationOut::scale_shift (model.rs:33-38, 217-221). */
int fluxb200_layernorm_modulate(const void* x, const void* shift, const void* scale, int64_t mod_bstride, void* out,
                                int32_t batch, int32_t rows_per_batch, int32_t dim, float eps,
                                fluxb200_stream_t stream);

/* QK RMS-norm + RoPE + head-major relayout. qkv: bf16 [B*rows, ld] (q|k|v at cols 0, H*128, 2*H*128);
 * pe_cos/pe_sin: bf16 [L,64]; Q,K,V: bf16 [B,H,L,128], this stream's tokens start at l_off.
 * Replaces QkNorm (model.rs:186-209) + apply_rope (model.rs:86-95) + the transposes of SelfAttention::qkv. */
int fluxb200_qknorm_rope(const void* qkv, int64_t ld, int32_t batch, int32_t rows_per_batch, int32_t H, int32_t L,
                         int32_t l_off, const void* wq, const void* wk, const void* pe_cos, const void* pe_sin,
                         void* Q, void* K, void* V, float eps, fluxb200_stream_t stream);


/* ------------------------------------------------------------------------------------------------
 * The 12 symbols of the reference's existing FFI, same names and argument lists
 * (diffusion_rs_backend/src/bitsandbytes/ffi.rs:5-114; C side kernels/bitsandbytes/dequant.cu:172-232).
 * `stream` is a CUstream; the dequantize_8bit_kernel_* family runs on the legacy default stream as in the reference.
 * ---------------------------------------------------------------------------------------------- */
void dequantize_blockwise_f32_int8(float* code, unsigned char* a, float* absmax, float* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_f32_fp4(float* code, unsigned char* a, float* absmax, float* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_f32_nf4(float* code, unsigned char* a, float* absmax, float* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_f16_int8(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_f16_fp4(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_f16_nf4(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_bf16_int8(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_bf16_fp4(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_blockwise_bf16_nf4(float* code, unsigned char* a, float* absmax, void* out, int blocksize, int n, fluxb200_stream_t stream);
void dequantize_8bit_kernel_f32(const int8_t* weight, const float* scb, float* out, int row, int col, int n);
void dequantize_8bit_kernel_f16(const int8_t* weight, const float* scb, void* out, int row, int col, int n);
void dequantize_8bit_kernel_bf16(const int8_t* weight, const float* scb, void* out, int row, int col, int n);

/* GGUF Q4_K super-blocks (144 B / 256 weights) -> bf16, the `dequantize_w` semantics of GgufMatMul
 * (diffusion_rs_backend/src/gguf/mod.rs:29-31; k_quants.rs:1568-1599): f32 -> f16 -> bf16. n = element count. */
int fluxb200_dequantize_q4k_bf16(const void* blocks, void* out, int64_t n, fluxb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Model-level entry points: the FLUX transformer behind Flux::new / Flux::forward (models/flux/model.rs:722-833)
 * ---------------------------------------------------------------------------------------------- */
typedef struct fluxb200_model fluxb200_model;

typedef struct {
  int32_t in_channels;           /* 64 */
  int32_t pooled_projection_dim; /* 768 */
  int32_t joint_attention_dim;   /* 4096 */
  int32_t num_attention_heads;   /* 24 */
  int32_t num_layers;            /* 19 double-stream blocks */
  int32_t num_single_layers;     /* 38 single-stream blocks */
  int32_t guidance_embeds;       /* 1 for FLUX.1-dev, 0 for schnell */
} fluxb200_flux_config; /* == models/flux/model.rs:21-31 Config (quantization is sniffed from tensor names) */

/* tensor dtypes accepted by fluxb200_load_weight */
#define FLUXB200_DT_BF16 0
#define FLUXB200_DT_F32 1
#define FLUXB200_DT_U8 2
#define FLUXB200_DT_I8 3
#define FLUXB200_DT_F16 4
#define FLUXB200_DT_Q4K 10 /* GGUF Q4_K blocks; shape = logical [out, in] */

int fluxb200_model_create(const fluxb200_flux_config* cfg, fluxb200_model** out);
void fluxb200_model_destroy(fluxb200_model* m);

/* Hand one checkpoint tensor to the model under its diffusers name (what VarBuilder::get resolves in the reference,
 * e.g. "transformer_blocks.0.attn.to_q.weight", "...weight.absmax", "...weight.quant_state.bitsandbytes__nf4",
 * "...SCB").  `data` may be a host or a device pointer (is_device != 0).  The bytes are copied; the caller keeps
 * ownership of `data`. */
int fluxb200_model_load_weight(fluxb200_model* m, const char* name, const void* data, int32_t dtype,
                               const int64_t* shape, int32_t rank, int32_t is_device, fluxb200_stream_t stream);

/* After the last tensor: checks that every tensor Flux::new would `vb.get` is present, picks dense / bnb / gguf per
 * layer exactly like diffusion_rs_backend::linear* (lib.rs:197-266) and repacks into the kernel layouts. */
int fluxb200_model_finalize(fluxb200_model* m, fluxb200_stream_t stream);

/* Bytes of caller-owned workspace one forward needs for this problem size. */
int fluxb200_model_workspace_size(const fluxb200_model* m, int32_t batch, int32_t l_img, int32_t l_txt,
                                  uint64_t* bytes);

/* Bytes of caller-owned workspace fluxb200_model_denoise needs: the forward workspace plus the per-step tables of the
 * whole loop (vec_ and every AdaLN modulation vector of all n_timesteps - 1 steps are projected before the loop). */
int fluxb200_model_denoise_workspace_size(const fluxb200_model* m, int32_t batch, int32_t l_img, int32_t l_txt,
                                          int32_t n_timesteps, uint64_t* bytes);

/* Flux::forward (model.rs:790-833).  All tensors are device pointers:
 *   img bf16 [B,l_img,64]; img_ids bf16 [B,l_img,3]; txt bf16 [B,l_txt,4096]; txt_ids bf16 [B,l_txt,3];
 *   timesteps f32 [B]; y bf16 [B,768]; guidance f32 [B] or NULL; out bf16 [B,l_img,64]. */
int fluxb200_model_forward(fluxb200_model* m, const void* img, const void* img_ids, const void* txt,
                           const void* txt_ids, const void* timesteps, const void* y, const void* guidance,
                           void* out, int32_t batch, int32_t l_img, int32_t l_txt, void* workspace,
                           uint64_t workspace_bytes, fluxb200_stream_t stream);

/* Sampler::sample (pipelines/sampling.rs:25-48): the Euler flow-matching loop over `timesteps` (host f64[n_t]),
 * updating `img` in place.  Everything that does not depend on the evolving latent is hoisted out of the loop (RoPE
 * table, txt projection, vec_ and all modulation projections of every step); one step is captured as a CUDA graph
 * (cached per workspace address / geometry in the handle) and replayed n_t - 1 times, the Euler update being the
 * epilogue of the step's last GEMM.  Nothing here synchronises: `timesteps` is consumed before the call returns (the
 * scalars travel as kernel parameters) and all work is enqueued on `stream`.
 * workspace: fluxb200_model_denoise_workspace_size bytes. */
int fluxb200_model_denoise(fluxb200_model* m, void* img, const void* img_ids, const void* txt, const void* txt_ids,
                           const void* y, float guidance_scale, const double* timesteps, int32_t n_timesteps,
                           int32_t batch, int32_t l_img, int32_t l_txt, void* workspace, uint64_t workspace_bytes,
                           fluxb200_stream_t stream);

/* After fluxb200_model_denoise: *used_graph = 1 if the steps were CUDA-graph replays, 0 if the kernels were launched
 * one by one ("step_graph" flag off, profiling on, or capture unavailable - *note then says why).  Either pointer may
 * be NULL. */
int fluxb200_model_denoise_info(const fluxb200_model* m, int32_t* used_graph, const char** note);

/* Debug / parity taps: copy an internal activation of the last forward into `out` (device, bf16).
 * which: 0 = vec_ [B,3072], 1 = img stream after the double blocks [B,l_img,3072], 2 = txt stream [B,l_txt,3072],
 *        3 = joint stream after the single blocks [B,l,3072], 4 = pe_cos [B,l,64], 5 = pe_sin [B,l,64]. */
int fluxb200_model_tap(fluxb200_model* m, int32_t which, void* out, uint64_t out_bytes, fluxb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * VAE decode: AutoEncoderKl::decode -> Decoder::forward (models/vaes/autoencoder_kl.rs:112-119, vae.rs:437-455)
 * ---------------------------------------------------------------------------------------------- */
typedef struct fluxb200_vae fluxb200_vae;

typedef struct {
  int32_t latent_channels;       /* 16 */
  int32_t out_channels;          /* 3 */
  int32_t block_out_channels[4]; /* 128, 256, 512, 512 */
  int32_t layers_per_block;      /* 2 */
  int32_t norm_num_groups;       /* 32 */
  int32_t mid_block_add_attention;
  float scaling_factor;          /* 0.3611 */
  float shift_factor;            /* 0.1159 */
} fluxb200_vae_config; /* == autoencoder_kl.rs:15-32 (decoder-relevant fields) */

int fluxb200_vae_create(const fluxb200_vae_config* cfg, fluxb200_vae** out);
void fluxb200_vae_destroy(fluxb200_vae* v);
/* Tensor names as the reference resolves them under the `decoder` prefix (vae.rs:371-433), e.g.
 * "decoder.up_blocks.0.resnets.1.conv1.weight" [Cout,Cin,3,3], "decoder.mid_block.attentions.0.to_q.weight" [512,512]. */
int fluxb200_vae_load_weight(fluxb200_vae* v, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                             int32_t rank, int32_t is_device, fluxb200_stream_t stream);
int fluxb200_vae_finalize(fluxb200_vae* v, fluxb200_stream_t stream);
int fluxb200_vae_workspace_size(const fluxb200_vae* v, int32_t batch, int32_t h, int32_t w, uint64_t* bytes);

/* VAEModel::decode: z bf16 NCHW [B, 16, h, w] -> out bf16 NCHW [B, 3, 8h, 8w]. */
int fluxb200_vae_decode(fluxb200_vae* v, const void* z, void* out, int32_t batch, int32_t h, int32_t w,
                        void* workspace, uint64_t workspace_bytes, fluxb200_stream_t stream);

/* Tail of FluxPipeline::forward (pipelines/flux/mod.rs:327-332): unpack the packed latents [B, h2*w2, 64], apply
 * z/scaling_factor + shift_factor, decode, clamp(-1,1) -> (x+1)*127.5 -> u8.  out is [B, 16*h2, 16*w2, 3] (image
 * layout) when nchw == 0, or the reference tensor layout [B, 3, 16*h2, 16*w2] when nchw != 0. */
int fluxb200_vae_decode_packed_u8(fluxb200_vae* v, const void* packed, void* out_u8, int32_t batch, int32_t h2,
                                  int32_t w2, int32_t nchw, void* workspace, uint64_t workspace_bytes,
                                  fluxb200_stream_t stream);

/* Operator-level: 2-D convolution on NHWC bf16 activations, stride 1, padding k/2, k in {1,3}; weight is the
 * repacked [Cout, k, k, Cin] tensor.  out = bf16(bf16(conv) + bias) (+ res).  Replaces Conv2d::forward
 * (nn/conv.rs:212-230 -> cuda_backend/mod.rs:1545-1599) without materialising an im2col buffer. */
int fluxb200_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* res, void* out,
                         int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize,
                         fluxb200_stream_t stream);
/* Repack a conv weight [Cout, Cin, k, k] -> [Cout, k, k, Cin]. */
int fluxb200_repack_conv_weight(const void* w, void* out, int32_t Cout, int32_t Cin, int32_t ksize,
                                fluxb200_stream_t stream);
/* GroupNorm (32 groups) + optional SiLU on NHWC bf16; stats_scratch: >= N*32*2 doubles of device memory.
 * Replaces nn::GroupNorm::forward (nn/group_norm.rs:39-74) (+ Activation::Silu). */
int fluxb200_groupnorm_nhwc(const void* x, const void* weight, const void* bias, void* out, int32_t N, int32_t HW,
                            int32_t C, int32_t groups, float eps, int32_t silu, void* stats_scratch,
                            fluxb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Text encoders: what FluxPipeline::forward runs once per prompt before the denoising loop
 * (pipelines/flux/mod.rs:236-262).  Token ids come from the caller (the tokenizers are not part of this library);
 * weights are bf16 tensors under the checkpoint names the reference's VarBuilder resolves.  Same ownership rules as
 * the model handles above: caller-owned ids / outputs / workspace, library-owned weight copies.
 * ---------------------------------------------------------------------------------------------- */
typedef struct fluxb200_t5 fluxb200_t5;
typedef struct {
  int32_t vocab_size, d_model, d_kv /* 64 */, d_ff, num_layers, num_heads;
  int32_t relative_attention_num_buckets, relative_attention_max_distance;
  float layer_norm_epsilon;
} fluxb200_t5_config; /* == models/t5/mod.rs:75-93 T5Config (encoder, feed_forward_proj = "gated-gelu") */

int fluxb200_t5_create(const fluxb200_t5_config* cfg, fluxb200_t5** out);
void fluxb200_t5_destroy(fluxb200_t5* m);
/* names: "shared.weight", "encoder.block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight",
 * "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", "encoder.block.{i}.layer.{0,1}.layer_norm.weight",
 * "encoder.block.{i}.layer.1.DenseReluDense.{wi_0,wi_1,wo}.weight", "encoder.final_layer_norm.weight" */
int fluxb200_t5_load_weight(fluxb200_t5* m, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                            int32_t rank, int32_t is_device, fluxb200_stream_t stream);
int fluxb200_t5_finalize(fluxb200_t5* m, fluxb200_stream_t stream);
int fluxb200_t5_workspace_size(const fluxb200_t5* m, int32_t batch, int32_t seq_len, uint64_t* bytes);
/* T5EncoderModel::forward (t5/mod.rs:659): ids int32 device [B, L] (L <= 512) -> out bf16 device [B, L, d_model]. */
int fluxb200_t5_forward(fluxb200_t5* m, const int32_t* ids, void* out, int32_t batch, int32_t seq_len, void* workspace,
                        uint64_t workspace_bytes, fluxb200_stream_t stream);

typedef struct fluxb200_clip fluxb200_clip;
typedef struct {
  int32_t vocab_size, projection_dim /* hidden width, = heads * 64 */, intermediate_size, max_position_embeddings;
  int32_t num_hidden_layers, num_attention_heads;
} fluxb200_clip_config; /* == models/clip/text.rs:22-31 ClipTextConfig (hidden_act = quick_gelu) */

int fluxb200_clip_create(const fluxb200_clip_config* cfg, fluxb200_clip** out);
void fluxb200_clip_destroy(fluxb200_clip* m);
/* names are relative to "text_model." : "embeddings.{token,position}_embedding.weight",
 * "encoder.layers.{i}.{layer_norm1,layer_norm2}.{weight,bias}", "encoder.layers.{i}.self_attn.{q,k,v,out}_proj.{weight,bias}",
 * "encoder.layers.{i}.mlp.{fc1,fc2}.{weight,bias}", "final_layer_norm.{weight,bias}" */
int fluxb200_clip_load_weight(fluxb200_clip* m, const char* name, const void* data, int32_t dtype, const int64_t* shape,
                              int32_t rank, int32_t is_device, fluxb200_stream_t stream);
int fluxb200_clip_finalize(fluxb200_clip* m, fluxb200_stream_t stream);
int fluxb200_clip_workspace_size(const fluxb200_clip* m, int32_t batch, int32_t seq_len, uint64_t* bytes);
/* ClipTextTransformer (clip/text.rs:291-316): ids int32 device [B, L] -> hidden bf16 [B, L, D] (forward_with_mask, may be
 * NULL) and pooled bf16 [B, D] = hidden at argmax(ids) (Module::forward, may be NULL). */
int fluxb200_clip_forward(fluxb200_clip* m, const int32_t* ids, void* hidden_out, void* pooled_out, int32_t batch,
                          int32_t seq_len, void* workspace, uint64_t workspace_bytes, fluxb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Launch accounting and optional per-kernel-class CUDA-event timing (evidence for bench.py; the reference has
 * only tracing spans, models/flux/model.rs:240-453).  Timing is off by default and costs two event records per
 * launch when enabled.
 * ---------------------------------------------------------------------------------------------- */
/* Runtime switches for A/B testing (each also has an environment variable read at first use):
 *   "qkrope_fusion"  QK-norm + RoPE fused into the q|k|v GEMM epilogue (default 1)
 *   "gemm_pair"      cta_group::2 GEMM (default 1; FLUXB200_GEMM_SINGLE_CTA=1 turns it off)
 *   "gemm_cl4"       cluster-of-4 GEMM with W multicast (default 0: measured slower; FLUXB200_GEMM_CL4=1)
 *   "dequant_mode"   how a quantised model weight reaches the tensor cores (FLUXB200_DEQUANT_MODE=n); same bits in all:
 *                      0 (default) fluxb200_model_denoise expands every quantised weight ONCE per call into a bf16 cache
 *                        carved from the caller's workspace (+2 B per quantised weight: 23.8 GB for a fully quantised
 *                        FLUX.1-dev, nothing on a 180 GB part) and every step runs dense GEMMs; the resident model stays
 *                        packed.  A single fluxb200_model_forward has no image to amortise over and behaves like 1.
 *                      1 staged per layer: each step expands each weight into an L2-sized staging buffer before its
 *                        GEMM, one weight ahead on a side stream ("dequant_overlap")
 *                      2 fused: the GEMM's producer warps expand the packed tile in shared memory (no bf16 copy in HBM)
 *   "fused_dequant"  older name: 1 = dequant_mode 2, 0 = dequant_mode 0
 *   "attn_variant"   build of the attention kernel, see attention.cu (default 0; FLUXB200_ATTN_VARIANT=n)
 *   "pdl"            programmatic dependent launch for GEMM / attention / LN-modulate (default 1; FLUXB200_PDL=0)
 *   "step_graph"     fluxb200_model_denoise replays one captured CUDA graph per step (default 1; FLUXB200_STEP_GRAPH=0)
 *   "dequant_overlap" staged quantised path: the expansion of the next weight runs on a side stream under the current
 *                    GEMM (default 1; FLUXB200_DEQUANT_OVERLAP=0)
 *   "ln_reread"      ln_modulate re-reads the row from L1 in its second pass instead of keeping it in registers: half the
 *                    registers, twice the warps per SM (default 1; FLUXB200_LN_REREAD=0); same bits
 *   "gemm_big"       512x256-per-CTA-pair tiles for the long-K GEMMs (EXPERIMENT, default 0: measured 13 % slower than
 *                    the 256x256 tiles; FLUXB200_GEMM_BIG=1, =2 for every eligible GEMM; 3 / 4: the
 *                    same with the two sub-tiles side by side along N)
 * FLUXB200_GEMM_MAX_UNITS=n limits the pair GEMM to n CTA pairs (experiments only). */
int fluxb200_set_flag(const char* name, int value);
void fluxb200_profile_enable(int on);
int fluxb200_profile_kinds(void);
const char* fluxb200_profile_kind_name(int kind);
/* Synchronises the device; ms/flops/bytes/count must each hold fluxb200_profile_kinds() entries. */
int fluxb200_profile_collect(double* ms, double* flops, double* bytes, unsigned long long* count);
/* Kernel launches issued by the library since load; kind < 0 => all kinds. */
unsigned long long fluxb200_launch_count(int kind);

#ifdef __cplusplus
}
#endif
#endif /* FLUXB200_H */
