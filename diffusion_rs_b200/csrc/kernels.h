// fluxb200 — launchers of the non-GEMM kernels (internal).
#pragma once
#include "internal.h"

namespace fb {

struct GemvJob {
  const bf16* w;     // [N, K] row-major
  const bf16* bias;  // [N] or null
  long long out_off;  // out_base[out_off + b * out_ld + n]
  long long out_ld;
  int N;
  int row_begin;  // prefix sum of N over the jobs of one launch
  int fused_bias;  // 1: bf16(acc + bias) (gguf f32 path); 0: bf16(bf16(acc) + bias) (rank-2 dense / bnb path)
  int pad_;
};

int launch_ln_modulate(const bf16* x, long long in_bstride_rows, int in_row_off, int rows_per_batch, int batch,
                       const bf16* shift, const bf16* scale, long long mod_bstride, bf16* out, int D, float eps,
                       cudaStream_t stream);
int launch_qknorm_rope(const bf16* qkv, long long ld, int rows_per_batch, int batch, int H, int L, int l_off,
                       const bf16* wq, const bf16* wk, const bf16* pe_cos, const bf16* pe_sin, long long pe_bstride,
                       bf16* Q, bf16* K, bf16* V, float eps, cudaStream_t stream);
int launch_gemv_jobs(const GemvJob* jobs_dev, int njobs, int row_base, int total_rows, const bf16* x, long long x_ld,
                     int B, int K, bf16* out_base, cudaStream_t stream);
int launch_silu(const bf16* x, bf16* y, long long n, cudaStream_t stream);
int launch_timestep_embedding(const float* t, bf16* out, int B, int dim, cudaStream_t stream);
int launch_vec_combine(const bf16* a, const bf16* g, const bf16* y, bf16* out, int n, cudaStream_t stream);
int launch_euler(bf16* img, const bf16* pred, float dt, long long n, cudaStream_t stream);
int launch_affine(const bf16* x, bf16* y, float mul, float add, long long n, cudaStream_t stream);

// ---- quantised-weight expansion (quant.cu) ----
int launch_dequant_bnb4(const uint8_t* packed, const float* absmax, bf16* out, int blocksize, long long n, int is_nf4,
                        cudaStream_t stream);
int launch_dequant_int8(const int8_t* w, const float* scb, bf16* out, int col, long long n, cudaStream_t stream);
int launch_dequant_q4k(const uint8_t* blocks, bf16* out, long long n, cudaStream_t stream);

// ---- VAE decode kernels (vae.cu) ----
int launch_groupnorm_silu(const bf16* x, const bf16* w, const bf16* b, bf16* y, int N, int HW, int C, int groups,
                          float eps, int apply_silu, double* stats, cudaStream_t stream);
int launch_repack_conv_weight(const bf16* w, bf16* out, int Cout, int Cin, int taps, cudaStream_t stream);
int launch_upsample2x_nhwc(const bf16* x, bf16* y, int N, int H, int W, int C, cudaStream_t stream);
int launch_softmax_rows_bf16(bf16* x, long long rows, int cols, cudaStream_t stream);
int launch_nchw_to_nhwc(const bf16* x, bf16* y, int N, int C, int H, int W, cudaStream_t stream);
int launch_nhwc_to_nchw(const bf16* x, bf16* y, int N, int C, int H, int W, cudaStream_t stream);
int launch_transpose_2d(const bf16* x, bf16* y, int rows, int cols, cudaStream_t stream);
int launch_postprocess_u8(const bf16* nhwc, uint8_t* out, int N, int C, int H, int W, int to_nchw,
                          cudaStream_t stream);
int launch_unpack_latents(const bf16* packed, bf16* nhwc, int N, int h2, int w2, float inv_scale, float shift,
                          cudaStream_t stream);

}  // namespace fb
