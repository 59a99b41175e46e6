// fluxb200 — tcgen05 flash attention (non-causal, head_dim 128) for the joint txt||img sequence.
//
// Replaces the reference's naive f32 path  softmax(q.k^T * scale).v  that materialises the LxL scores
// (diffusion_rs_backend/src/ops.rs:247-262, called from diffusion_rs_core/src/models/flux/model.rs:40-51, 97-102).
//
// One CTA = one (batch, head) x 256 query rows (two 128-row tiles that ping-pong on the tensor core):
//   warp 0       TMA producer: Q0,Q1 once; K_j, V_j through 2-stage rings (128B swizzle)
//   warp 1       MMA issuer  : S_g = Q_g.K_j^T (SS) and O_g += P_g.V_j (A = P from TMEM, B = V MN-major from smem)
//   warps 2..5   softmax for tile 0, warps 6..9 softmax for tile 1 (ping-pong on the MUFU unit through a pair of
//                named barriers): one thread per query row, online softmax in
//                fp32 with lazy rescaling of the TMEM-resident O accumulator; P is written back to TMEM as bf16
//                over the S columns it came from.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512) fp32 columns.
#include "internal.h"
#include "ptx.cuh"

namespace fb {

static constexpr int HD = 128;            // head dim
static constexpr int TQ = 128;            // query rows per tile
static constexpr int TKV = 128;           // kv rows per block
static constexpr int TILE_BYTES = TQ * HD * 2;   // 32 KB (two 16 KB swizzle-atom columns)
static constexpr int HALF_BYTES = TILE_BYTES / 2;
static constexpr int ATT_THREADS = 320;
static constexpr bool PINGPONG = true;
static constexpr size_t ATT_SMEM = 6 * TILE_BYTES + 1024 + 256;

struct AttnParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;  // 3D {128, L, B*H}, box {64, 128, 1}
  int B, H, L, l_split;
  int q_pairs;  // ceil(L / 256)
  int nkv;      // ceil(L / 128)
  float sl2;    // scale * log2(e)
  bf16* out_a;
  bf16* out_b;
  long long ld_a, ld_b;
};

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tcgen05_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // 2 tiles
  uint8_t* sK = smem + 2 * TILE_BYTES;   // 2 stages
  uint8_t* sV = smem + 4 * TILE_BYTES;   // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
  uint64_t* q_full = bars;           // [1]
  uint64_t* k_full = bars + 1;       // [2]
  uint64_t* k_empty = bars + 3;      // [2]
  uint64_t* v_full = bars + 5;       // [2]
  uint64_t* v_empty = bars + 7;      // [2]
  uint64_t* s_full = bars + 9;       // [2] per q tile
  uint64_t* p_ready = bars + 11;     // [2]
  uint64_t* o_done = bars + 13;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.x / P.q_pairs;
  const int qp = blockIdx.x - bh * P.q_pairs;
  const int q_row0 = qp * 2 * TQ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_q);
    tma_prefetch_desc(&P.tmap_k);
    tma_prefetch_desc(&P.tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
      for (int g = 0; g < 2; ++g) {
        tma_load_3d(sQ + g * TILE_BYTES, &P.tmap_q, q_full, 0, q_row0 + g * TQ, bh);
        tma_load_3d(sQ + g * TILE_BYTES + HALF_BYTES, &P.tmap_q, q_full, 64, q_row0 + g * TQ, bh);
      }
      for (int j = 0; j < P.nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_3d(sK + st * TILE_BYTES, &P.tmap_k, &k_full[st], 0, j * TKV, bh);
        tma_load_3d(sK + st * TILE_BYTES + HALF_BYTES, &P.tmap_k, &k_full[st], 64, j * TKV, bh);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_3d(sV + st * TILE_BYTES, &P.tmap_v, &v_full[st], 0, j * TKV, bh);
        tma_load_3d(sV + st * TILE_BYTES + HALF_BYTES, &P.tmap_v, &v_full[st], 64, j * TKV, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TKV, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, HD, 0, 1);   // P (TMEM)    x V (MN-major)
      // The issuing thread is a single lane: keep its instruction stream short.  All shared-memory descriptors are
      // built once; inside the loops a descriptor is `base + compile-time constant` (the address field is the low
      // 14 bits in 16-byte units and smem < 256 KB, so the add never carries into the next field).
      const uint64_t qd0 = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t qd1 = umma_smem_desc_sw128(smem_u32(sQ) + TILE_BYTES, 16, 1024);
      const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t kd1 = umma_smem_desc_sw128(smem_u32(sK) + TILE_BYTES, 16, 1024);
      const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(sV), HALF_BYTES, 1024);
      const uint64_t vd1 = umma_smem_desc_sw128(smem_u32(sV) + TILE_BYTES, HALF_BYTES, 1024);
      auto issue_s = [&](int g, int st) {
        const uint64_t qd = g ? qd1 : qd0;
        const uint64_t kd = st ? kd1 : kd0;
        const uint32_t ts = tmem_base + g * 128;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          // d 0..63 sit in the first swizzle-atom column, 64..127 in the second; 32 B per k-step inside an atom
          const uint64_t off = static_cast<uint64_t>(((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4);
          umma_ss(ts, qd + off, kd + off, idesc_s, k != 0);
        }
      };
      auto issue_pv = [&](int g, int st, bool accumulate) {
        const uint64_t vd = st ? vd1 : vd0;
        const uint32_t to = tmem_base + 256 + g * 128;
        const uint32_t tp = tmem_base + g * 128;
#pragma unroll
        for (int k = 0; k < TKV / 16; ++k) {
          // A: 16 bf16 of P per row = 8 TMEM columns per k-step. B: 16 kv rows = 2048 B per k-step;
          // the two 64-wide d chunks are HALF_BYTES apart (LBO), 8-row groups 1024 B apart (SBO).
          umma_ts(to, tp + k * 8, vd + static_cast<uint64_t>((k * 2048) >> 4), idesc_o, accumulate || (k != 0));
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      tc_commit(&s_full[0]);
      issue_s(1, 0);
      tc_commit(&s_full[1]);
      tc_commit(&k_empty[0]);  // K_0 consumed by S0_0 / S1_0
      for (int j = 0; j < P.nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const bool has_next = (j + 1 < P.nkv);
        const int nst = (j + 1) & 1;
        mbar_wait(&v_full[st], ph);
        if (has_next) mbar_wait(&k_full[nst], ((j + 1) >> 1) & 1);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_ready[g], j & 1);
          tc_fence_after();
          issue_pv(g, st, j > 0);
          if (has_next) {
            issue_s(g, nst);
            tc_commit(&s_full[g]);
          } else {
            tc_commit(&o_done[g]);
          }
        }
        tc_commit(&v_empty[st]);
        if (has_next) tc_commit(&k_empty[nst]);  // K_{j+1} fully consumed by S0/S1_{j+1}
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue ----------------
    const int g = (warp - 2) >> 2;  // query tile
    const int q = warp & 3;         // TMEM lane quarter
    const int r = q * 32 + lane;    // row in tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + g * 128 + lane_off;
    const float sl2 = P.sl2;
    float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
    float l_run = 0.f;
    // The two tiles' softmax warps share the SM's MUFU unit.  A ping-pong pair of named barriers lets only one tile
    // be in its exp2 phase at a time, which keeps the tiles in anti-phase: while tile g exponentiates, the tensor
    // core runs the other tile's P.V and next Q.K^T.
    if (PINGPONG && g == 1) named_bar_arrive(1, 256);  // tile 0 goes first

    for (int j = 0; j < P.nkv; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      uint32_t s[128];
      tmem_ld32(tS + 0, &s[0]);
      tmem_ld32(tS + 32, &s[32]);
      tmem_ld32(tS + 64, &s[64]);
      tmem_ld32(tS + 96, &s[96]);
      tc_wait_ld();
      const int kv_valid = P.L - j * TKV;  // >= 128 except on the last block
      float bmax = -INFINITY;
      if (kv_valid >= TKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i) bmax = fmaxf(bmax, __uint_as_float(s[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          if (i >= kv_valid) s[i] = __float_as_uint(-INFINITY);
          bmax = fmaxf(bmax, __uint_as_float(s[i]));
        }
      }
      const float m_new = fmaxf(m_run, bmax);
      // lazy rescale: keep a stale max while exp2 stays below 2^8
      const bool need = (m_new - m_run) * sl2 > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        float factor = 1.0f;
        if (need) {
          factor = ex2_approx((m_run - m_new) * sl2);  // 0 on the first block
          m_run = m_new;
          l_run *= factor;
        }
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st32(tO + c * 32, o);
          }
        }
      }
      const float mb = m_run * sl2;
      float lsum = 0.f;
      if (PINGPONG) named_bar_sync(1 + g, 256);  // wait for this tile's turn on the MUFU
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + 2 * i]), sl2, -mb));
          float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), sl2, -mb));
          lsum += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);
      }
      l_run += lsum;
      if (PINGPONG) named_bar_arrive(2 - g, 256);  // hand the MUFU to the other tile
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[g]);
    }

    // epilogue: O / l -> bf16 -> global
    mbar_wait(&o_done[g], 0);
    tc_fence_after();
    const int l = q_row0 + g * TQ + r;
    const int b = bh / P.H;
    const int h = bh - b * P.H;
    bf16* dst = nullptr;
    if (l < P.L) {
      if (l < P.l_split)
        dst = P.out_a + (static_cast<long long>(b) * P.l_split + l) * P.ld_a + h * HD;
      else
        dst = P.out_b + (static_cast<long long>(b) * (P.L - P.l_split) + (l - P.l_split)) * P.ld_b + h * HD;
    }
    const float inv = 1.0f / l_run;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tc_wait_ld();
      if (dst) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[i * 8 + 0]) * inv, __uint_as_float(o[i * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[i * 8 + 2]) * inv, __uint_as_float(o[i * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[i * 8 + 4]) * inv, __uint_as_float(o[i * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[i * 8 + 6]) * inv, __uint_as_float(o[i * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_attention(const AttnDesc& d, cudaStream_t stream) {
  FB_REQUIRE(d.q && d.k && d.v, "attention: null q/k/v");
  FB_REQUIRE(d.B > 0 && d.H > 0 && d.L > 0, "attention: empty problem");
  FB_REQUIRE(d.l_split >= 0 && d.l_split <= d.L, "attention: bad l_split");
  FB_REQUIRE(d.l_split == 0 || d.out_a != nullptr, "attention: out_a required when l_split > 0");
  FB_REQUIRE(d.l_split == d.L || d.out_b != nullptr, "attention: out_b required");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(ATT_SMEM)));
    attr_set = true;
  }
  AttnParams P;
  memset(&P, 0, sizeof(P));
  const uint64_t bh = static_cast<uint64_t>(d.B) * d.H;
  const void* ptrs[3] = {d.q, d.k, d.v};
  CUtensorMap* maps[3] = {&P.tmap_q, &P.tmap_k, &P.tmap_v};
  for (int i = 0; i < 3; ++i) {
    FB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "attention: q/k/v must be 16-byte aligned");
    int rc = encode_tmap_3d(maps[i], ptrs[i], HD, d.L, bh, HD * 2, static_cast<uint64_t>(d.L) * HD * 2, 64, TQ, 1);
    if (rc) return rc;
  }
  P.B = d.B, P.H = d.H, P.L = d.L, P.l_split = d.l_split;
  P.q_pairs = (d.L + 2 * TQ - 1) / (2 * TQ);
  P.nkv = (d.L + TKV - 1) / TKV;
  P.sl2 = d.scale * 1.4426950408889634f;
  P.out_a = d.out_a, P.out_b = d.out_b, P.ld_a = d.ld_a, P.ld_b = d.ld_b;
  const int grid = static_cast<int>(bh) * P.q_pairs;
  ProfScope _ps(KK_ATTN, 4.0 * bh * static_cast<double>(d.L) * d.L * HD, 4.0 * bh * d.L * HD * 2.0, stream);
  count_launch(KK_ATTN);
  attention_tcgen05_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(P);
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
