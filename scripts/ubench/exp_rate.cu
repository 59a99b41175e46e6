// Microbenchmark of the softmax exp2 phase: clocks per 128x128 score tile for different instruction mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffusion_rs_b200/csrc exp_rate.cu -o exp_rate
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace fb;

FB_DEVICE void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
FB_DEVICE void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
FB_DEVICE void ex2_poly2(float& y0, float& y1, float x0, float x1) {
  x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
  float t0, t1, u0, u1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, 12582912.0f, 12582912.0f);
  fadd2(u0, u1, t0, t1, -12582912.0f, -12582912.0f);
  ffma2(f0, f1, u0, u1, -1.0f, -1.0f, x0, x1);
  ffma2(p0, p1, f0, f1, 0.055922036f, 0.055922036f, 0.242640083f, 0.242640083f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.693121034f, 0.693121034f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.999924481f, 0.999924481f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// COLS per thread (128: 4 warps, 64: 8 warps); MODE bits: 1 = MUFU ex2, 2 = scale FFMA2, 4 = row-sum FADD2, 8 = bf16 pack,
// 16 = tcgen05.st of P, 32 = tcgen05.ld of S each iteration; POLY_MOD as in the kernel
template <int COLS, int MODE, int POLY_MOD>
__global__ void __launch_bounds__(COLS == 128 ? 128 : 256, 1) k(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const int q = warp & 3, h = warp >> 2;
  const uint32_t tS = tb + ((q * 32u) << 16) + h * 64;
  const uint32_t tP = tb + 256 + ((q * 32u) << 16) + h * 32;
  uint32_t s[COLS];
#pragma unroll
  for (int i = 0; i < COLS; ++i) s[i] = __float_as_uint(-0.01f * i - lane);
  if (MODE & 32) {
#pragma unroll
    for (int c = 0; c < COLS / 32; ++c) tmem_st32(tS + c * 32, &s[c * 32]);
    tc_wait_st();
  }
  float ls0 = 0.f, ls1 = 0.f;
  const float sl2 = 0.1275f, nmb = 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE & 32) {
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) tmem_ld32(tS + c * 32, &s[c * 32]);
      tc_wait_ld();
    }
#pragma unroll
    for (int c = 0; c < COLS / 16; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x0 = __uint_as_float(s[c * 16 + 2 * i]), x1 = __uint_as_float(s[c * 16 + 2 * i + 1]), p0, p1;
        if (MODE & 2) ffma2(x0, x1, x0, x1, sl2, sl2, nmb, nmb);
        if (POLY_MOD > 0 && (i % POLY_MOD == POLY_MOD - 1)) ex2_poly2(p0, p1, x0, x1);
        else if (MODE & 1) { p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
        else { p0 = x0; p1 = x1; }
        if (MODE & 4) fadd2(ls0, ls1, ls0, ls1, p0, p1);
        if (MODE & 8) pk[i] = pack_bf16(p0, p1); else pk[i] = __float_as_uint(p0) ^ __float_as_uint(p1);
        if (!(MODE & 32)) { s[c * 16 + 2 * i] = __float_as_uint(p0); s[c * 16 + 2 * i + 1] = __float_as_uint(p1); }
      }
      if (MODE & 16) tmem_st8(tP + c * 8, pk);
      else { ls0 += __uint_as_float(pk[0] ^ pk[1] ^ pk[2] ^ pk[3] ^ pk[4] ^ pk[5] ^ pk[6] ^ pk[7]); }
    }
    if (MODE & 16) tc_wait_st();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float acc = ls0 + ls1;
#pragma unroll
  for (int i = 0; i < COLS; ++i) acc += __uint_as_float(s[i]);
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int COLS, int MODE, int POLY_MOD>
void run(const char* name, long long* d, float* sink) {
  const int iters = 200;
  k<COLS, MODE, POLY_MOD><<<148, COLS == 128 ? 128 : 256>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  printf("cols/thread %3d  %-52s poly_mod %d : %7.1f clk per 128x128 tile\n", COLS, name, POLY_MOD, avg / iters);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  float* sink; cudaMalloc(&sink, 1024 * sizeof(float));
  run<128, 1, 0>("mufu only", d, sink);
  run<128, 3, 0>("mufu + ffma2", d, sink);
  run<128, 7, 0>("mufu + ffma2 + fadd2", d, sink);
  run<128, 15, 0>("mufu + ffma2 + fadd2 + pack", d, sink);
  run<128, 31, 0>("mufu + ffma2 + fadd2 + pack + sttm", d, sink);
  run<128, 63, 0>("full (ldtm + ... + sttm)", d, sink);
  run<128, 63, 4>("full", d, sink);
  run<128, 63, 2>("full", d, sink);
  run<128, 14, 0>("no mufu: ffma2 + fadd2 + pack", d, sink);
  run<128, 8, 0>("pack only", d, sink);
  run<128, 6, 0>("ffma2 + fadd2", d, sink);
  run<64, 1, 0>("mufu only", d, sink);
  run<64, 15, 0>("mufu + ffma2 + fadd2 + pack", d, sink);
  run<64, 63, 0>("full", d, sink);
  run<64, 63, 4>("full", d, sink);
  run<64, 63, 3>("full", d, sink);
  run<64, 63, 2>("full", d, sink);
  run<64, 62, 1>("all poly", d, sink);
  return 0;
}
