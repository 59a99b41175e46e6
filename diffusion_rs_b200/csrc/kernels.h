// fluxb200 — launchers of the non-GEMM kernels (internal).
#pragma once
#include "internal.h"

namespace fb {

// per-step scalars of one chunk of the denoising loop, passed by value as a kernel parameter
struct StepScalars {
  static constexpr int N = 256;
  float t[N];   // t_curr of each step
  float dt[N];  // t_prev - t_curr
};

int launch_ln_modulate(const bf16* x, long long in_bstride_rows, int in_row_off, int rows_per_batch, int batch,
                       const bf16* shift, const bf16* scale, long long mod_bstride, bf16* out, int D, float eps,
                       cudaStream_t stream, const int* step_ptr = nullptr, long long step_stride = 0);
// two row segments (the img and txt streams of a double block) in one launch; output rows are contiguous
struct LnInput {
  const bf16* x;
  const bf16* shift;
  const bf16* scale;
  long long in_bstride_rows;
  int in_row_off, rows_per_batch, batch;
};
int launch_ln_modulate2(const LnInput* in, int nseg, long long mod_bstride, bf16* out, int D, float eps,
                        cudaStream_t stream, const int* step_ptr = nullptr, long long step_stride = 0);
int launch_qknorm_rope(const bf16* qkv, long long ld, int rows_per_batch, int batch, int H, int L, int l_off,
                       const bf16* wq, const bf16* wk, const bf16* pe_cos, const bf16* pe_sin, long long pe_bstride,
                       bf16* Q, bf16* K, bf16* V, float eps, cudaStream_t stream);
int launch_silu(const bf16* x, bf16* y, long long n, cudaStream_t stream);
int launch_timestep_embedding(const float* t, bf16* out, int B, int dim, cudaStream_t stream);
int launch_vec_combine(const bf16* a, const bf16* g, const bf16* y, bf16* out, int rows, int B, int Dm,
                       cudaStream_t stream);
int launch_step_scalars(const StepScalars& v, int s0, int n, int B, int C, float guidance, float* t_all, float* g_all,
                        bf16* dt_tab, cudaStream_t stream);
int launch_step_advance(int* step, cudaStream_t stream);
int launch_copy_rows(void* dst, long long dst_bstride_bytes, const void* src, long long src_bstride_bytes,
                     long long bytes_per_batch, int batch, cudaStream_t stream);
int launch_euler(bf16* img, const bf16* pred, float dt, long long n, cudaStream_t stream);
int launch_affine(const bf16* x, bf16* y, float mul, float add, long long n, cudaStream_t stream);

// ---- quantised-weight expansion (quant.cu) ----
int launch_dequant_bnb4(const uint8_t* packed, const float* absmax, bf16* out, int blocksize, long long n, int is_nf4,
                        cudaStream_t stream);
int launch_dequant_int8(const int8_t* w, const float* scb, bf16* out, int col, long long n, cudaStream_t stream);
int launch_dequant_q4k(const uint8_t* blocks, bf16* out, long long n, cudaStream_t stream);
// all members of one fused Linear in one launch (kind = QB_* of internal.h)
struct DequantJob {
  const uint8_t* packed;
  const float* absmax;
  const float* scb;
  bf16* out;
  long long n;  // weights
  int kind, blocksize, col, pad_;
};
struct DequantBatch {
  static constexpr int MAX = 4;
  DequantJob job[MAX];
  int count;
};
int launch_dequant_batch(const DequantBatch& batch, cudaStream_t stream);

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
