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

/* Joint attention. q,k,v: bf16 [B,H,L,128]; out: bf16 [B,L,H*128] (== transpose(1,2).flatten_from(2)).
 * Replaces diffusion_rs_backend::ops::sdpa (ops.rs:247-262) + the casts in model.rs:40-51, softcapping = 1. */
int fluxb200_sdpa(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t L,
                  float scale, fluxb200_stream_t stream);

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

#ifdef __cplusplus
}
#endif
#endif /* FLUXB200_H */
