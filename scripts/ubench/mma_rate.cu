// Microbenchmark: issue rate of tcgen05.mma shapes used by the attention kernel (clock64 around batches, one CTA/SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffusion_rs_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace fb;

// mode: 0 SS N=128 K-major B | 1 TS N=128 MN-major B | 2 SS N=128 MN-major B | 3 TS N=128 K-major B
//       4 alternate (8x TS MN) + (8x SS K) | 5 SS N=256 | 6 SS N=64 | 7 SS N=128 with concurrent st.shared traffic from 8 warps
//       8 alternate, with one commit+wait per batch pair (dependency through mbarrier like the real kernel)
__global__ void __launch_bounds__(320, 1) k(int mode, int batches, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  volatile int* stopflag = reinterpret_cast<volatile int*>(slot + 1);
  if (threadIdx.x == 0) *stopflag = 0;
  __syncthreads();
  if (warp == 1 && lane == 0) {
    const uint32_t id_kk128 = umma_idesc_bf16(128, 128, 0, 0), id_mn128 = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t id_kk256 = umma_idesc_bf16(128, 256, 0, 0), id_kk64 = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t a0 = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bK = umma_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
    const uint64_t bMN = umma_smem_desc_sw128(smem_u32(smem) + 65536, 16384, 1024);
    auto ss_batch = [&](uint32_t d, uint64_t bd, uint32_t idesc) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
        umma_ss(d, a0 + off, bd + off, idesc, kk != 0);
      }
    };
    auto ss_batch_mn = [&](uint32_t d) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
        umma_ss(d, a0 + off, bMN + static_cast<uint64_t>((kk * 2048) >> 4), id_mn128, kk != 0);
      }
    };
    auto ts_batch = [&](uint32_t d, uint32_t a, bool mn) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (mn) umma_ts(d, a + kk * 8, bMN + static_cast<uint64_t>((kk * 2048) >> 4), id_mn128, 1);
        else {
          const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
          umma_ts(d, a + kk * 8, bK + off, id_kk128, 1);
        }
      }
    };
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      switch (mode) {
        case 0: case 7: ss_batch(tb, bK, id_kk128); break;
        case 1: ts_batch(tb + 256, tb, true); break;
        case 2: ss_batch_mn(tb); break;
        case 3: ts_batch(tb + 256, tb, false); break;
        case 4: ts_batch(tb + 256, tb, true); ss_batch(tb + 128, bK, id_kk128); break;
        case 5: ss_batch(tb, bK, id_kk256); break;
        case 6: ss_batch(tb, bK, id_kk64); break;
        case 9:  // SS N=128, two accumulators alternating every MMA
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
            umma_ss(tb + (kk & 1) * 128, a0 + off, bK + off, id_kk128, kk > 1);
          }
          break;
        case 10: ss_batch(tb, bK, umma_idesc_bf16(128, 192, 0, 0)); break;
        case 11:  // PV k-step and S k-step interleaved (different accumulators)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
            umma_ts(tb + 256, tb + kk * 8, bMN + static_cast<uint64_t>((kk * 2048) >> 4), id_mn128, 1);
            umma_ss(tb + 128, a0 + off, bK + off, id_kk128, kk != 0);
          }
          break;
        case 12:  // SS N=128, four accumulators round-robin
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
            umma_ss(tb + (kk & 3) * 128, a0 + off, bK + off, id_kk128, kk > 3);
          }
          break;
        case 13: ss_batch(tb, bK, umma_idesc_bf16(128, 160, 0, 0)); break;
        case 8:
          ts_batch(tb + 256, tb, true); ss_batch(tb + 128, bK, id_kk128);
          tc_commit(bar); mbar_wait(bar, ph); ph ^= 1; tc_fence_after();
          break;
      }
    }
    if (mode != 8) { tc_commit(bar); mbar_wait(bar, 0); }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    *stopflag = 1;
  } else if (warp >= 2 && mode == 7) {
    // background shared-memory store traffic (like P/TMA writes): 8 warps x 16 B per thread
    uint4* dst = reinterpret_cast<uint4*>(smem + 98304) + (threadIdx.x - 64);
    uint4 v = make_uint4(1, 2, 3, 4);
    while (!*stopflag) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[i * 256] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// CTA pair: M=256 (128 rows per CTA), B split across the pair (N/2 rows per CTA)
// mode: 0 SS N=128 | 1 SS N=256 | 2 SS N=128 two accumulators alternating | 3 SS N=64
__global__ void __launch_bounds__(128, 1) k2(int mode, int batches, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_2sm(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (rank == 0 && warp == 1 && lane == 0) {
    const uint64_t a0 = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bK = umma_smem_desc_sw128(smem_u32(smem) + 32768, 16, 1024);
    const int N = mode == 1 ? 256 : (mode == 3 ? 64 : 128);
    const uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const uint32_t idesc_mn = umma_idesc_bf16(256, 128, 0, 1);
    const uint64_t bMN = umma_smem_desc_sw128(smem_u32(smem) + 65536, 16384, 1024);
    long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (mode >= 4) {
        if (mode != 6) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // PV-like: A from TMEM, B = [128 kv x 64 d] MN-major per CTA
            umma_ts_2sm(tb + 256, tb + kk * 8, bMN + static_cast<uint64_t>((kk * 2048) >> 4), idesc_mn, 1);
        }
        if (mode == 5) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // S-like: B = [64 kv x 128 d] K-major per CTA
            const uint64_t offq = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
            const uint64_t offk = static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4);
            umma_ss_2sm(tb + 128, a0 + offq, bK + offk, idesc, kk != 0);
          }
        }
        if (mode == 6) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // SS with MN-major B
            umma_ss_2sm(tb, a0 + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                        bMN + static_cast<uint64_t>((kk * 2048) >> 4), idesc_mn, kk != 0);
        }
        continue;
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t off = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
        const uint32_t d = tb + ((mode == 2) ? (kk & 1) * 128 : 0);
        umma_ss_2sm(d, a0 + off, bK + off, idesc, mode == 2 ? kk > 1 : kk != 0);
      }
    }
    tc_commit_2sm(bar, 1);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tb, 512); }
}

int main() {
  const int smem = 192 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  const char* names[] = {"SS N=128 B K-major (S=QK^T)", "TS N=128 B MN-major (O+=PV)", "SS N=128 B MN-major", "TS N=128 B K-major",
                         "alternate TS-MN / SS-K", "SS N=256", "SS N=64", "SS N=128 + st.shared background", "alternate + commit/wait per pair",
                         "SS N=128, 2 accumulators alternating", "SS N=192", "TS-MN / SS-K interleaved per MMA", "SS N=128, 4 accumulators",
                         "SS N=160"};
  const int batches = 256;
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 14; ++mode) {
      k<<<grid, 320, smem>>>(mode, batches, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= grid;
      const int per = (mode == 4 || mode == 8 || mode == 11) ? 16 : 8;
      printf("grid %3d mode %d %-36s clk/MMA = %7.1f\n", grid, mode, names[mode], avg / (batches * per));
    }
  }
  {
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* n2[] = {"pair SS M=256 N=128", "pair SS M=256 N=256", "pair SS M=256 N=128, 2 accumulators", "pair SS M=256 N=64", "pair TS N=128 B MN-major (PV)",
                        "pair alternate TS-MN / SS-K", "pair SS N=128 B MN-major"};
    for (int mode = 0; mode < 7; ++mode) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr, cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, k2, mode, batches, d);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pair mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(74);
      cudaMemcpy(h.data(), d, 74 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= 74;
      printf("pairs 74 mode %d %-36s clk/MMA = %7.1f\n", mode, n2[mode], avg / (batches * (mode == 5 ? 16 : 8)));
    }
  }
  return 0;
}
