// fluxb200 — quantised-weight expansion: bitsandbytes NF4 / FP4 / blockwise-int8 / LLM.int8 and GGUF Q4_K.
//
// Exports the 12 `extern "C"` symbols of the reference's own FFI with identical signatures
// (diffusion_rs_backend/src/bitsandbytes/ffi.rs:5-114; C side kernels/bitsandbytes/dequant.cu:172-232) so the Rust
// `BnbLinear` links against this library unchanged, plus bf16 launchers used by the fused linear path.
// Semantics are the reference CUDA kernel's (ground truth per SURVEY N4):
//   4-bit : out[2j] = T(LUT[q>>4] * absmax[j / (blocksize/2)]),  out[2j+1] = T(LUT[q&15] * absmax[...])
//   8-bit : out[i]  = T(code[q[i]] * absmax[i / blocksize])
//   int8  : out[i]  = T(float(w[i]) * scb[i / col] / 127)
// HBM-bound: one thread expands 16 packed bytes (32 weights) with 128-bit loads/stores.
#include <cuda_fp16.h>

#include <algorithm>

#include "internal.h"
#include "kernels.h"
#include "ptx.cuh"

namespace fb {

__constant__ float kNF4[16] = {-1.0f,
                               -0.6961928009986877f,
                               -0.5250730514526367f,
                               -0.39491748809814453f,
                               -0.28444138169288635f,
                               -0.18477343022823334f,
                               -0.09105003625154495f,
                               0.0f,
                               0.07958029955625534f,
                               0.16093020141124725f,
                               0.24611230194568634f,
                               0.33791524171829224f,
                               0.44070982933044434f,
                               0.5626170039176941f,
                               0.7229568362236023f,
                               1.0f};
// sign-magnitude tree of dDequantizeFP4Tree (dequant.cu:12-37); bit 3 is the sign
__constant__ float kFP4[16] = {0.0f,  5.208333333e-03f,  0.66666667f,  1.0f,  0.33333333f,  0.5f,  0.16666667f,  0.25f,
                               -0.0f, -5.208333333e-03f, -0.66666667f, -1.0f, -0.33333333f, -0.5f, -0.16666667f, -0.25f};

template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __half cvt_out<__half>(float v) {
  return __float2half_rn(v);
}
template <>
__device__ __forceinline__ bf16 cvt_out<bf16>(float v) {
  return __float2bfloat16_rn(v);
}

// DATA_TYPE: 1 = FP4, 2 = NF4.  `half_block` = blocksize/2 = packed bytes per absmax entry.
template <typename T, int DATA_TYPE>
__global__ void __launch_bounds__(256) dequant_4bit_kernel(const uint8_t* __restrict__ A,
                                                           const float* __restrict__ absmax, T* __restrict__ out,
                                                           int half_block, long long n) {
  // the 16-entry code book goes to shared memory: lanes index it with different nibbles, which a __constant__ bank
  // would serialise (one address per cycle) — that alone held the first version of this kernel at 0.7 TB/s
  __shared__ float lut[16];
  if (threadIdx.x < 16) lut[threadIdx.x] = (DATA_TYPE == 2) ? kNF4[threadIdx.x] : kFP4[threadIdx.x];
  __syncthreads();
  const long long nbytes = (n + 1) / 2;
  const long long byte0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 16;
  if (byte0 >= nbytes) return;
  const bool fast = (byte0 + 16 <= nbytes) && (2 * (byte0 + 16) <= n) && (half_block % 16 == 0) &&
                    ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (fast) {
    const uint4 pk = *reinterpret_cast<const uint4*>(A + byte0);
    const float am = __ldg(&absmax[byte0 / half_block]);
    const uint32_t words[4] = {pk.x, pk.y, pk.z, pk.w};
    T vals[32];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t q = (words[w] >> (8 * b)) & 0xffu;
        vals[(w * 4 + b) * 2] = cvt_out<T>(lut[q >> 4] * am);
        vals[(w * 4 + b) * 2 + 1] = cvt_out<T>(lut[q & 15] * am);
      }
    constexpr int VEC = 16 / sizeof(T);  // elements per 16-byte store
    T* o = out + byte0 * 2;
#pragma unroll
    for (int i = 0; i < 32 / VEC; ++i) reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<const uint4*>(vals)[i];
  } else {
    for (int j = 0; j < 16; ++j) {
      const long long bi = byte0 + j;
      if (bi >= nbytes) break;
      const uint32_t q = A[bi];
      const float am = absmax[bi / half_block];
      if (2 * bi < n) out[2 * bi] = cvt_out<T>(lut[q >> 4] * am);
      if (2 * bi + 1 < n) out[2 * bi + 1] = cvt_out<T>(lut[q & 15] * am);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dequant_8bit_blockwise_kernel(const float* __restrict__ code,
                                                                     const uint8_t* __restrict__ A,
                                                                     const float* __restrict__ absmax,
                                                                     T* __restrict__ out, int blocksize, long long n) {
  __shared__ float scode[256];
  scode[threadIdx.x] = code[threadIdx.x];
  __syncthreads();
  const long long i0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 16;
  for (int j = 0; j < 16; ++j) {
    const long long i = i0 + j;
    if (i >= n) return;
    out[i] = cvt_out<T>(scode[A[i]] * absmax[i / blocksize]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dequant_int8_rowwise_kernel(const int8_t* __restrict__ w,
                                                                   const float* __restrict__ scb, T* __restrict__ out,
                                                                   int col, long long n) {
  const long long i0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 16;
  for (int j = 0; j < 16; ++j) {
    const long long i = i0 + j;
    if (i >= n) return;
    out[i] = cvt_out<T>((static_cast<float>(w[i]) * scb[i / col]) / 127.f);
  }
}

template <typename T, int DATA_TYPE>
static void dequant_blockwise(const float* code, const uint8_t* A, const float* absmax, T* out, int blocksize, int n,
                              cudaStream_t stream) {
  if (n <= 0) return;
  if (DATA_TYPE > 0) {
    const long long nbytes = (static_cast<long long>(n) + 1) / 2;
    const unsigned grid = static_cast<unsigned>((nbytes + 16 * 256 - 1) / (16 * 256));
    dequant_4bit_kernel<T, DATA_TYPE><<<grid, 256, 0, stream>>>(A, absmax, out, blocksize / 2, n);
  } else {
    const unsigned grid = static_cast<unsigned>((static_cast<long long>(n) + 16 * 256 - 1) / (16 * 256));
    dequant_8bit_blockwise_kernel<T><<<grid, 256, 0, stream>>>(code, A, absmax, out, blocksize, n);
  }
}

int launch_dequant_bnb4(const uint8_t* packed, const float* absmax, bf16* out, int blocksize, long long n, int is_nf4,
                        cudaStream_t stream) {
  FB_REQUIRE(blocksize >= 2 && blocksize % 2 == 0, "dequant_bnb4: bad blocksize");
  const long long nbytes = (n + 1) / 2;
  const unsigned grid = static_cast<unsigned>((nbytes + 16 * 256 - 1) / (16 * 256));
  ProfScope _ps(KK_DEQUANT, 0, 2.5 * n, stream);
  count_launch(KK_DEQUANT);
  if (is_nf4)
    dequant_4bit_kernel<bf16, 2><<<grid, 256, 0, stream>>>(packed, absmax, out, blocksize / 2, n);
  else
    dequant_4bit_kernel<bf16, 1><<<grid, 256, 0, stream>>>(packed, absmax, out, blocksize / 2, n);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_dequant_int8(const int8_t* w, const float* scb, bf16* out, int col, long long n, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>((n + 16 * 256 - 1) / (16 * 256));
  count_launch(KK_DEQUANT);
  dequant_int8_rowwise_kernel<bf16><<<grid, 256, 0, stream>>>(w, scb, out, col, n);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Model path: every member of one fused Linear (q|k|v|proj_mlp ...) is expanded into the bf16 staging buffer by ONE
// launch (blockIdx.y = member).  Fully coalesced: a thread reads 4 packed bytes and writes one 16-byte vector, so a warp
// reads 128 contiguous bytes and writes 512 contiguous bytes per iteration; 4 independent iterations are in flight
// per thread.  The 4-bit code books are expanded to a byte -> (first, second) value-pair table in shared memory
// (the high nibble is the first weight, dequant.cu:155-156).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void q4k_scale_min(const uint8_t* sc, int is, float d, float dmin, float& s, float& m) {
  uint32_t dd, mm;
  if (is < 4) {
    dd = sc[is] & 63;
    mm = sc[is + 4] & 63;
  } else {
    dd = (sc[is + 4] & 0xF) | ((sc[is - 4] >> 6) << 4);
    mm = (sc[is + 4] >> 4) | ((sc[is] >> 6) << 4);
  }
  s = __fmul_rn(d, static_cast<float>(dd));
  m = __fmul_rn(dmin, static_cast<float>(mm));
}

// Grid-stride over (units, members); <= 40 registers per thread.  (A one-block-per-SM persistent variant, meant to
// co-reside with the GEMM's CTAs while the expansion runs on a side stream, measured 1.6 TB/s instead of 2.7 and made
// the overlapped step slower: the expansion became the critical path behind the short txt-stream GEMMs.)
__global__ void __launch_bounds__(256, 6) dequant_batch_kernel(const DequantBatch batch) {
  __shared__ float2 lut2[256];
  constexpr int IT = 8;
  const long long stride = static_cast<long long>(gridDim.x) * (256 * IT);
  int lut_kind = 0;
  for (int mi = 0; mi < batch.count; ++mi) {
    const DequantJob& j = batch.job[mi];
    if ((j.kind == QB_NF4 || j.kind == QB_FP4) && j.kind != lut_kind) {
      __syncthreads();
      const float* cb = j.kind == QB_NF4 ? kNF4 : kFP4;
      lut2[threadIdx.x] = make_float2(cb[threadIdx.x >> 4], cb[threadIdx.x & 15]);
      __syncthreads();
      lut_kind = j.kind;
    }
    if (j.kind == QB_NF4 || j.kind == QB_FP4) {
      // unit = 4 packed bytes -> 8 weights (16 bytes of bf16)
      const long long units = j.n / 8;
      const int bs_units = j.blocksize / 8;  // units per absmax entry (blocksize is a power of two >= 64)
      for (long long base = static_cast<long long>(blockIdx.x) * (256 * IT) + threadIdx.x; base < units; base += stride) {
        uint32_t pk[IT];
        float am[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const long long u = base + it * 256;
          pk[it] = u < units ? __ldg(reinterpret_cast<const uint32_t*>(j.packed) + u) : 0u;
          am[it] = u < units ? __ldg(j.absmax + u / bs_units) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const long long u = base + it * 256;
          if (u >= units) break;
          uint4 o;
          float2 c;
          c = lut2[pk[it] & 0xffu];         o.x = pack_bf16(c.x * am[it], c.y * am[it]);
          c = lut2[(pk[it] >> 8) & 0xffu];  o.y = pack_bf16(c.x * am[it], c.y * am[it]);
          c = lut2[(pk[it] >> 16) & 0xffu]; o.z = pack_bf16(c.x * am[it], c.y * am[it]);
          c = lut2[pk[it] >> 24];           o.w = pack_bf16(c.x * am[it], c.y * am[it]);
          reinterpret_cast<uint4*>(j.out)[u] = o;
        }
      }
    } else if (j.kind == QB_INT8) {
      // unit = 8 int8 weights -> 16 bytes of bf16
      const long long units = j.n / 8;
      const int col_units = j.col / 8;
      for (long long base = static_cast<long long>(blockIdx.x) * (256 * IT) + threadIdx.x; base < units; base += stride) {
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const long long u = base + it * 256;
          if (u >= units) break;
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(j.packed) + u);
          const float sc = __ldg(j.scb + u / col_units);
          const uint32_t w2[2] = {q.x, q.y};
          uint32_t o[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int v0 = static_cast<int8_t>((w2[h >> 1] >> (16 * (h & 1))) & 0xffu);
            const int v1 = static_cast<int8_t>((w2[h >> 1] >> (16 * (h & 1) + 8)) & 0xffu);
            o[h] = pack_bf16((static_cast<float>(v0) * sc) / 127.f, (static_cast<float>(v1) * sc) / 127.f);
          }
          reinterpret_cast<uint4*>(j.out)[u] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    } else {
      // Q4_K: lane-unit = 4 q bytes of one 64-weight group -> 4 low-nibble + 4 high-nibble weights; 32 lane-units per
      // 144-byte super-block
      const long long nblocks = j.n / 256;
      const long long units = nblocks * 32;
      for (long long base = static_cast<long long>(blockIdx.x) * (256 * IT) + threadIdx.x; base < units; base += stride) {
#pragma unroll 2
        for (int it = 0; it < IT; ++it) {
          const long long u = base + it * 256;
          if (u >= units) break;
          const long long blk = u >> 5;
          const int lane = static_cast<int>(u & 31);
          const uint8_t* p = j.packed + blk * 144;
          const float d = __half2float(*reinterpret_cast<const __half*>(p));
          const float dmin = __half2float(*reinterpret_cast<const __half*>(p + 2));
          const int g = lane >> 3, off = (lane & 7) * 4;
          float d1, m1, d2, m2;
          q4k_scale_min(p + 4, 2 * g, d, dmin, d1, m1);
          q4k_scale_min(p + 4, 2 * g + 1, d, dmin, d2, m2);
          const uint32_t q4 = *reinterpret_cast<const uint32_t*>(p + 16 + g * 32 + off);
          bf16* o = j.out + blk * 256 + g * 64 + off;
          float lo[4], hi[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t q = (q4 >> (8 * b)) & 0xff;
            lo[b] = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(d1, static_cast<float>(q & 0xF)), m1)));
            hi[b] = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(d2, static_cast<float>(q >> 4)), m2)));
          }
          uint2 v;
          v.x = pack_bf16(lo[0], lo[1]), v.y = pack_bf16(lo[2], lo[3]);
          *reinterpret_cast<uint2*>(o) = v;
          v.x = pack_bf16(hi[0], hi[1]), v.y = pack_bf16(hi[2], hi[3]);
          *reinterpret_cast<uint2*>(o + 32) = v;
        }
      }
    }
  }
}

int launch_dequant_batch(const DequantBatch& batch, cudaStream_t stream) {
  FB_REQUIRE(batch.count >= 1 && batch.count <= DequantBatch::MAX, "dequant_batch: 1..4 members");
  long long max_units = 0;
  double bytes = 0;
  for (int i = 0; i < batch.count; ++i) {
    const DequantJob& j = batch.job[i];
    FB_REQUIRE(j.packed && j.out && j.n > 0, "dequant_batch: null member");
    FB_REQUIRE((reinterpret_cast<uintptr_t>(j.packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.out) & 15) == 0,
               "dequant_batch: 16-byte alignment");
    long long units = 0;
    if (j.kind == QB_NF4 || j.kind == QB_FP4) {
      FB_REQUIRE(j.absmax && j.blocksize >= 64 && (j.blocksize & (j.blocksize - 1)) == 0 && j.n % j.blocksize == 0,
                 "dequant_batch: 4-bit members need absmax and a power-of-two blocksize >= 64 dividing the weight");
      units = j.n / 8, bytes += 2.5 * j.n + 4.0 * j.n / j.blocksize;
    } else if (j.kind == QB_INT8) {
      FB_REQUIRE(j.scb && j.col % 8 == 0, "dequant_batch: int8 members need SCB and K % 8 == 0");
      units = j.n / 8, bytes += 3.0 * j.n;
    } else if (j.kind == QB_Q4K) {
      FB_REQUIRE(j.n % 256 == 0, "dequant_batch: Q4_K element count must be a multiple of 256");
      units = j.n / 8, bytes += 2.5625 * j.n;  // 32 lane-units per 256 weights
    } else {
      return fail("dequant_batch: bad kind");
    }
    max_units = std::max(max_units, units);
  }
  ProfScope _ps(KK_DEQUANT, 0, bytes, stream);
  count_launch(KK_DEQUANT);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((max_units + 2047) / 2048, 16LL * num_sms()));
  dequant_batch_kernel<<<grid, 256, 0, stream>>>(batch);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GGUF Q4_K: 256 weights per 144-byte super-block {f16 d, f16 dmin, 12 B packed 6-bit scales/mins, 128 B nibbles}
// (diffusion_rs_common/src/core/quantized/k_quants.rs:130-136, to_float :1568-1599, get_scale_min_k4 utils.rs:49-59).
// GgufMatMul::dequantize_w = dequantize (f32) -> f16 -> out dtype (gguf/mod.rs:29-31): both roundings are kept.
// One warp per super-block; lane l expands bytes 4l..4l+3 of the 128 nibble bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dequant_q4k_kernel(const uint8_t* __restrict__ blocks, bf16* __restrict__ out,
                                                          long long nblocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long blk = blockIdx.x * 8LL + warp;
  if (blk >= nblocks) return;
  const uint8_t* p = blocks + blk * 144;
  const float d = __half2float(*reinterpret_cast<const __half*>(p));
  const float dmin = __half2float(*reinterpret_cast<const __half*>(p + 2));
  const uint8_t* sc = p + 4;
  const uint8_t* qs = p + 16;
  // lane -> 64-weight group j = lane/8, 4 consecutive bytes inside the group's 32 bytes
  const int j = lane >> 3;
  const int off = (lane & 7) * 4;
  auto scale_min = [&](int is, float& s, float& m) {
    uint32_t dd, mm;
    if (is < 4) {
      dd = sc[is] & 63;
      mm = sc[is + 4] & 63;
    } else {
      dd = (sc[is + 4] & 0xF) | ((sc[is - 4] >> 6) << 4);
      mm = (sc[is + 4] >> 4) | ((sc[is] >> 6) << 4);
    }
    s = __fmul_rn(d, static_cast<float>(dd));
    m = __fmul_rn(dmin, static_cast<float>(mm));
  };
  float d1, m1, d2, m2;
  scale_min(2 * j, d1, m1);
  scale_min(2 * j + 1, d2, m2);
  const uint32_t q4 = *reinterpret_cast<const uint32_t*>(qs + j * 32 + off);
  bf16* o = out + blk * 256 + j * 64 + off;
  float lo[4], hi[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t q = (q4 >> (8 * b)) & 0xff;
    lo[b] = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(d1, static_cast<float>(q & 0xF)), m1)));
    hi[b] = __half2float(__float2half_rn(__fsub_rn(__fmul_rn(d2, static_cast<float>(q >> 4)), m2)));
  }
  uint2 u;
  u.x = pack_bf16(lo[0], lo[1]), u.y = pack_bf16(lo[2], lo[3]);
  *reinterpret_cast<uint2*>(o) = u;
  u.x = pack_bf16(hi[0], hi[1]), u.y = pack_bf16(hi[2], hi[3]);
  *reinterpret_cast<uint2*>(o + 32) = u;
}

int launch_dequant_q4k(const uint8_t* blocks, bf16* out, long long n, cudaStream_t stream) {
  FB_REQUIRE(n % 256 == 0, "dequant_q4k: element count must be a multiple of 256");
  const long long nblocks = n / 256;
  ProfScope _ps(KK_DEQUANT, 0, 2.5625 * n, stream);
  count_launch(KK_DEQUANT);
  dequant_q4k_kernel<<<static_cast<unsigned>((nblocks + 7) / 8), 256, 0, stream>>>(blocks, out, nblocks);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb

// ------------------------------------------------------------------------------------------------
// The reference's FFI surface (same names, same argument lists, void return; bitsandbytes/ffi.rs:5-114)
// ------------------------------------------------------------------------------------------------
using fb::dequant_blockwise;
extern "C" {
#define FB_BNB(NAME, T, DT)                                                                                 \
  void NAME(float* code, unsigned char* A, float* absmax, T* out, int blocksize, const int n, cudaStream_t stream) { \
    dequant_blockwise<T, DT>(code, A, absmax, out, blocksize, n, stream);                                   \
  }
FB_BNB(dequantize_blockwise_f32_int8, float, 0)
FB_BNB(dequantize_blockwise_f32_fp4, float, 1)
FB_BNB(dequantize_blockwise_f32_nf4, float, 2)
FB_BNB(dequantize_blockwise_f16_int8, __half, 0)
FB_BNB(dequantize_blockwise_f16_fp4, __half, 1)
FB_BNB(dequantize_blockwise_f16_nf4, __half, 2)
FB_BNB(dequantize_blockwise_bf16_int8, __nv_bfloat16, 0)
FB_BNB(dequantize_blockwise_bf16_fp4, __nv_bfloat16, 1)
FB_BNB(dequantize_blockwise_bf16_nf4, __nv_bfloat16, 2)
#undef FB_BNB

#define FB_INT8(NAME, T)                                                                               \
  void NAME(const int8_t* weight, const float* scb, T* out, const int row, const int col, const int n) { \
    (void)row;                                                                                         \
    if (n <= 0) return;                                                                                \
    const unsigned grid = static_cast<unsigned>((static_cast<long long>(n) + 16 * 256 - 1) / (16 * 256)); \
    fb::dequant_int8_rowwise_kernel<T><<<grid, 256, 0, 0>>>(weight, scb, out, col, n); /* legacy default stream, as the reference */ \
  }
FB_INT8(dequantize_8bit_kernel_f32, float)
FB_INT8(dequantize_8bit_kernel_f16, __half)
FB_INT8(dequantize_8bit_kernel_bf16, __nv_bfloat16)
#undef FB_INT8
}
