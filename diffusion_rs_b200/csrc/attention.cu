// fluxb200 — tcgen05 flash attention (non-causal, head_dim 128) for the joint txt||img sequence.
//
// Replaces the reference's naive f32 path  softmax(q.k^T * scale).v  that materialises the LxL scores
// (diffusion_rs_backend/src/ops.rs:247-262, called from diffusion_rs_core/src/models/flux/model.rs:40-51, 97-102).
//
// One CTA = one (batch, head) x 256 query rows (two 128-row tiles that ping-pong on the tensor core):
//   warps 0..3   softmax for tile 0, warps 4..7 softmax for tile 1 (ping-pong on the MUFU unit through a pair of
//                named barriers): one thread per query row, online softmax in fp32 with lazy rescaling of the
//                TMEM-resident O accumulator; P is written back to TMEM as bf16 over the S columns it came from.
//   warp 8       TMA producer: Q0,Q1 once; K_j, V_j through rings (128B swizzle)
//   warp 9       MMA issuer  : S_g = Q_g.K_j^T (SS) and O_g += P_g.V_j (A = P from TMEM, B = V MN-major from smem)
// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512) fp32 columns.
//
// What the measurements of this round say (scripts/attn_variants.py = clock64 trace of CTA 0, scripts/ubench/*.cu):
//  * per tile the work is a strict chain  S ready -> TMEM load -> row max -> exp2 -> P -> P.V -> next Q.K^T;  the two
//    tiles fill each other's gaps.  First cut: 3480 clk per 128-row kv block (tensor pipe 52 %).
//  * tcgen05.mma with M=128 (cta_group::1) issues at best every ~92 clk for any N <= 192 (77 clk when SS and TS MMAs
//    alternate), i.e. 70-83 % of the tensor peak for the N=128 shapes of attention; the same MMAs for a CTA pair
//    (M=256) take the 64-clk floor.  The pair build of this kernel (PAIR=true) is correct and its MMAs are faster,
//    but the cross-CTA P handoff adds ~300 clk of latency to the chain, so it is not (yet) faster end to end.
//  * the MMA-issuing warp must not lose issue slots to the softmax warps (highest warp id wins arbitration) and
//    its operands must be uniform (no per-MMA R2UR/ELECT waterfall): 3480 -> 3250 clk.
//  * the exp2 phase is MUFU-bound (16/clk/SM); a packed f32x2 degree-3 polynomial for every 4th pair: 3250 -> 3030.
//  * two threads per row (RPT=2: half the TMEM load / max scan / exp2 work per thread) is correct but slower
//    (3430-3640 clk): five warps per sub-partition slow the TMEM load and the max exchange costs more than it saves.
//  * inside the DiT step (power-capped clocks) all of this is worth 844 -> 1006 TFLOP/s.
#include "internal.h"
#include "ptx.cuh"

namespace fb {

static constexpr int HD = 128;            // head dim
static constexpr int TQ = 128;            // query rows per tile
static constexpr int TKV = 128;           // kv rows per block
static constexpr int TILE_BYTES = TQ * HD * 2;   // 32 KB (two 16 KB swizzle-atom columns)
static constexpr int HALF_BYTES = TILE_BYTES / 2;

struct AttnParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;  // 3D {128, L, B*H}, box {64, 128, 1}
  int B, H, L, l_split;
  int q_pairs;  // ceil(L / 256)
  int nkv;      // ceil(L / 128)
  float sl2;    // scale * log2(e)
  bf16* out_a;
  bf16* out_b;
  long long ld_a, ld_b;
  long long* trace;  // debug: clock64 stamps of CTA 0, [kv block][tile][8]; nullptr in production
};

// packed f32x2 arithmetic (FFMA2 / FADD2): one issue slot for two lanes of work
FB_DEVICE void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
FB_DEVICE void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
FB_DEVICE float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 2^x for two lanes on the FMA pipe (see ex2_poly): 10 instructions per pair, no MUFU
FB_DEVICE void ex2_poly2(float& y0, float& y1, float x0, float x1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
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

// The exp2 phase of one thread's share of a 128x128 score tile (a whole query row, or half of it when two threads
// serve a row).  `s` holds the raw scores; P is written back over the first 64 TMEM columns of the tile as bf16
// pairs, 16 scores (8 packed columns = one k-step of O += P.V) at a time, and published to the MMA thread in NHAND
// instalments so that most of P.V runs under the remaining exponentials.  The wait for an instalment's TMEM stores
// is issued one chunk late (after the next chunk's arithmetic), so the MUFU never idles behind tcgen05.wait::st.
//   POLY_MASK bit i set: the i-th PAIR of every group of 8 pairs is evaluated on the FMA pipe (ex2_poly2) instead of the
//             MUFU unit (0 = all on MUFU, 0x88 = every 4th pair, 0xAA = every other pair).  Measured
//             (scripts/ubench/exp_rate.cu): a 128x128 tile costs 1600 clk with all exponentials on the MUFU (16/clk/SM)
//             and 1080 clk with every 4th pair on the FMA pipe.
template <int POLY_MASK, int NHAND, int NCH, class Arrive>
FB_DEVICE void softmax_exp_store(uint32_t (&s)[NCH * 16], uint32_t tP, float sl2, float nmb, float& l_run, int lane,
                                 Arrive&& arrive) {
  constexpr int CH = NCH / NHAND;  // chunks per instalment
  float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x0, x1, p0, p1;
      ffma2(x0, x1, __uint_as_float(s[c * 16 + 2 * i]), __uint_as_float(s[c * 16 + 2 * i + 1]), sl2, sl2, nmb, nmb);
      if ((POLY_MASK >> i) & 1) {
        ex2_poly2(p0, p1, x0, x1);
      } else {
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      fadd2(ls0, ls1, ls0, ls1, p0, p1);
      pk[i] = pack_bf16(p0, p1);
    }
    if (c > 0 && c % CH == 0) {  // chunks [c-CH, c) were stored one chunk's worth of work ago
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(c / CH - 1);
    }
    tmem_st8(tP + c * 8, pk);
  }
  l_run += ls0 + ls1;
}

// One kernel, two launch modes:
//   PAIR = false  one CTA = one (batch, head) x 256 query rows; tcgen05.mma.cta_group::1, M = 128.
//   PAIR = true   a 2-CTA cluster = one (batch, head) x 512 query rows; each CTA keeps its own two 128-row query
//                 tiles, softmax state and O accumulators, but the K/V block is split between the two CTAs (CTA r stages
//                 kv rows [64r, 64r+64) of K and head-dim columns [64r, 64r+64) of V) and the leader CTA's MMA thread
//                 issues every MMA for both (cta_group::2, M = 256).  Why: measured on B200
//                 (scripts/ubench/mma_rate.cu) a cta_group::1 MMA with M=128 never issues faster than one per ~92 clk
//                 for N <= 192, i.e. the N=128 MMAs of attention run at 70 % of the tensor peak, while the same MMA
//                 issued for a CTA pair takes the 64-clk floor, with half the operand traffic per SM.
//                 "Data is ready" barriers live in the leader CTA (remote arrivals, 2-SM TMA), "data has been
//                 consumed" barriers are signalled in both CTAs by multicast commits.
// Other knobs (selected at run time through the "attn_variant" flag; see launch_attention):
//   POLY_MASK see softmax_exp_store.
//   NHAND     number of instalments in which P is handed to the MMA thread (1, 2 or 4).
//   PP        ping-pong the two query tiles on the MUFU through a pair of named barriers.
//   RPT       threads per query row (1 or 2).  With 2, warps w and w+4 of a tile share a TMEM lane quarter and split the
//             128 score columns: half the TMEM load, max scan and exp2 work per thread, two issuing warps per
//             sub-partition in the exp2 phase; the two threads of a row agree on the running max through a
//             double-buffered shared-memory slot and a 64-thread named barrier.
// Warp roles (RPT=1): warps 0..3 softmax of tile 0, 4..7 softmax of tile 1, 8 TMA producer, 9 MMA issuer, 10..11 idle
// (RPT=2: 0..7, 8..15, 16, 17, 18..19).  The control warps come last on purpose: the sub-partition arbiter favours the
// highest warp id, and the MMA thread's issue latency sits on the critical path of both tiles.
template <bool PAIR>
struct AttnCfg {
  static constexpr int ST = PAIR ? 4 : 2;                           // K / V ring depth
  static constexpr int KV_BYTES = PAIR ? TILE_BYTES / 2 : TILE_BYTES;  // per stage: K [64|128 kv x 128 d], V [128 kv x 64|128 d]
  static constexpr int K_HALF = KV_BYTES / 2;                       // the two 64-wide d halves of a K stage
  static constexpr int BAR_OFF = 2 * TILE_BYTES + 2 * ST * KV_BYTES;
  static constexpr int XCH_OFF = BAR_OFF + 512;  // float [parity][tile][half][row]: row-max / row-sum exchange (RPT=2)
  static constexpr size_t SMEM = XCH_OFF + 2 * 2 * 2 * 128 * sizeof(float);
};

template <bool PAIR, int POLY_MASK, int NHAND, bool PP, int RPT, bool TRACE, bool COOP = false>
__global__ void __launch_bounds__((COOP ? 12 : 8 * RPT + 4) * 32, 1) attention_tcgen05_kernel(const __grid_constant__ AttnParams P) {
  static_assert(RPT == 1 || (RPT == 2 && NHAND == 1), "two threads per row hand P over in one piece");
  static_assert(!COOP || RPT == 2, "COOP: the 8 softmax warps split every row of both tiles in two");
  constexpr int NSW = 4 * RPT;                   // softmax warps that serve one tile (arrivals per p_ready barrier)
  constexpr int W_TMA = COOP ? NSW : 2 * NSW;    // first control warp = number of softmax warps
  constexpr int W_MMA = W_TMA + 1;
  constexpr int COLS = 128 / RPT;     // score columns per softmax thread
  constexpr int SM_THREADS = W_TMA * 32;
  using Cfg = AttnCfg<PAIR>;
  constexpr int ST = Cfg::ST;
  constexpr int KVB = Cfg::KV_BYTES;
  // 128B-swizzled tiles need 1024-byte alignment.  The kernel has no static shared memory, so the dynamic window
  // starts at the CTA's shared base; declaring the alignment keeps every address a link-time constant, which lets
  // ptxas hold all MMA operands in uniform registers (checked at run time below).
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                     // 2 tiles x 32 KB
  uint8_t* sK = smem + 2 * TILE_BYTES;    // ST stages
  uint8_t* sV = sK + ST * KVB;            // ST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;                // [1]   (PAIR: leader's is used)
  uint64_t* k_full = bars + 1;            // [ST]  (leader)
  uint64_t* v_full = k_full + ST;         // [ST]  (leader)
  uint64_t* k_empty = v_full + ST;        // [ST]  both CTAs (multicast commit)
  uint64_t* v_empty = k_empty + ST;       // [ST]
  uint64_t* s_full = v_empty + ST;        // [2]   per q tile
  uint64_t* o_done = s_full + 2;          // [2]
  uint64_t* p_ready = o_done + 2;         // [4 instalments][2 tiles] (leader): one arrival per softmax warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_ready + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int bh = unit / P.q_pairs;  // q_pairs = query groups per (batch, head): 256 rows each (PAIR: 512)
  const int qp = unit - bh * P.q_pairs;
  const int q_row0 = PAIR ? (qp * 4 * TQ + static_cast<int>(rank) * 2 * TQ) : qp * 2 * TQ;
  long long* const trace = (TRACE && P.trace != nullptr && blockIdx.x == 0) ? P.trace : nullptr;

  if (warp == W_TMA && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("fluxb200: attention shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&P.tmap_q);
    tma_prefetch_desc(&P.tmap_k);
    tma_prefetch_desc(&P.tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&p_ready[i], (PAIR ? 2 : 1) * NSW);
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    if (PAIR) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // (peer) barriers initialised before any arrival
  tc_fence_after();
  // The CTA owns all 512 TMEM columns, so the allocation starts at column 0 / lane 0; the kernel relies on it so
  // that TMEM addresses are compile-time constants.
  if (*tmem_slot != 0) {
    if (threadIdx.x == 0) printf("fluxb200: unexpected TMEM base %u\n", *tmem_slot);
    __trap();
  }
  constexpr uint32_t TM_S = 0, TM_O = 256;
  // everything above overlapped the previous kernel's tail; q/k/v are valid from here on
  pdl_launch_dependents();
  pdl_wait();

  // Register file: 3 (RPT=2: 5) warps per SM sub-partition cap every thread at 168 (96) registers at launch; the
  // control warpgroup hands part of its share to the softmax warpgroups, which keep their score columns in registers
  // (120 + 2 x 192 = 3 x 168;  64 + 4 x 104 = 5 x 96).
  if (warp >= W_TMA) {
    if (RPT == 1 || COOP) asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    else                  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == W_TMA) {
      if (lane == 0) {
        // ---------------- TMA producer (PAIR: one per CTA; bytes are credited to the leader's barriers) ----------------
        if (PAIR) {
          const uint32_t qbar = mapa_u32(smem_u32(q_full), 0);
          if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * 2 * TILE_BYTES);
          for (int g = 0; g < 2; ++g) {
            tma_load_3d_2sm(sQ + g * TILE_BYTES, &P.tmap_q, qbar, 0, q_row0 + g * TQ, bh);
            tma_load_3d_2sm(sQ + g * TILE_BYTES + HALF_BYTES, &P.tmap_q, qbar, 64, q_row0 + g * TQ, bh);
          }
        } else {
          mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
          for (int g = 0; g < 2; ++g) {
            tma_load_3d(sQ + g * TILE_BYTES, &P.tmap_q, q_full, 0, q_row0 + g * TQ, bh);
            tma_load_3d(sQ + g * TILE_BYTES + HALF_BYTES, &P.tmap_q, q_full, 64, q_row0 + g * TQ, bh);
          }
        }
        int st = 0;
        uint32_t ph = 0;
        for (int j = 0; j < P.nkv; ++j) {
          uint8_t* dk = sK + st * KVB;
          uint8_t* dv = sV + st * KVB;
          mbar_wait(&k_empty[st], ph ^ 1);
          if (PAIR) {
            const uint32_t kbar = mapa_u32(smem_u32(&k_full[st]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&k_full[st], 2 * KVB);
            tma_load_3d_2sm(dk, &P.tmap_k, kbar, 0, j * TKV + static_cast<int>(rank) * 64, bh);
            tma_load_3d_2sm(dk + Cfg::K_HALF, &P.tmap_k, kbar, 64, j * TKV + static_cast<int>(rank) * 64, bh);
          } else {
            mbar_arrive_expect_tx(&k_full[st], KVB);
            tma_load_3d(dk, &P.tmap_k, &k_full[st], 0, j * TKV, bh);
            tma_load_3d(dk + Cfg::K_HALF, &P.tmap_k, &k_full[st], 64, j * TKV, bh);
          }
          mbar_wait(&v_empty[st], ph ^ 1);
          if (PAIR) {
            const uint32_t vbar = mapa_u32(smem_u32(&v_full[st]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&v_full[st], 2 * KVB);
            tma_load_3d_2sm(dv, &P.tmap_v, vbar, static_cast<int>(rank) * 64, j * TKV, bh);
          } else {
            mbar_arrive_expect_tx(&v_full[st], KVB);
            tma_load_3d(dv, &P.tmap_v, &v_full[st], 0, j * TKV, bh);
            tma_load_3d(dv + HALF_BYTES, &P.tmap_v, &v_full[st], 64, j * TKV, bh);
          }
          if (++st == ST) st = 0, ph ^= 1;
        }
      }
    } else if (warp == W_MMA) {
      if (rank == 0) {
        // ---------------- MMA issuer (PAIR: leader CTA only) ----------------
        // The whole warp walks the loop (uniform control flow); one elected lane issues the MMAs and commits.
        const bool el = elect_one() != 0;
        // Everything this thread feeds to tcgen05.mma is a compile-time constant (shared-memory symbol + offset, TMEM
        // column, stage index through full unrolling): the operands stay in uniform registers and an MMA costs a
        // handful of issue slots.
        constexpr uint32_t idesc_s = umma_idesc_bf16(PAIR ? 2 * TQ : TQ, TKV, 0, 0);  // Q (K-major) x K (K-major)
        constexpr uint32_t idesc_o = umma_idesc_bf16(PAIR ? 2 * TQ : TQ, HD, 0, 1);   // P (TMEM) x V (MN-major)
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
        auto wait_full = [&](uint64_t* bar, uint32_t parity) {
          if (PAIR) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
        };
        auto commit = [&](uint64_t* bar) {
          if (el) {
            if (PAIR) tc_commit_2sm(bar, 3); else tc_commit(bar);
          }
        };
        auto issue_s = [&](int g, int st) {
          const uint64_t qd = umma_smem_desc_sw128(aQ + g * TILE_BYTES, 16, 1024);
          const uint64_t kd = umma_smem_desc_sw128(aK + st * KVB, 16, 1024);
          if (el)
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            // d 0..63 sit in the first swizzle-atom column, 64..127 in the second; 32 B per k-step inside an atom
            const uint64_t offq = static_cast<uint64_t>(((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4);
            const uint64_t offk = static_cast<uint64_t>(((k >> 2) * Cfg::K_HALF + (k & 3) * 32) >> 4);
            if (PAIR) umma_ss_2sm(TM_S + g * 128, qd + offq, kd + offk, idesc_s, k != 0);
            else      umma_ss(TM_S + g * 128, qd + offq, kd + offk, idesc_s, k != 0);
          }
        };
        // kv rows [k0*16, k1*16) of O_g += P_g.V
        auto issue_pv = [&](int g, int st, bool accumulate, int k0, int k1) {
          // B: 16 kv rows = 2048 B per k-step; 64-wide d chunks HALF_BYTES apart (LBO; PAIR: one chunk per CTA),
          // 8-row groups 1024 B apart (SBO).  A: 16 bf16 of P per row = 8 TMEM columns per k-step.
          const uint64_t vd = umma_smem_desc_sw128(aV + st * KVB, HALF_BYTES, 1024);
          if (el)
#pragma unroll
          for (int k = k0; k < k1; ++k) {
            const uint64_t off = static_cast<uint64_t>((k * 2048) >> 4);
            if (PAIR) umma_ts_2sm(TM_O + g * 128, TM_S + g * 128 + k * 8, vd + off, idesc_o, accumulate || (k != 0));
            else      umma_ts(TM_O + g * 128, TM_S + g * 128 + k * 8, vd + off, idesc_o, accumulate || (k != 0));
          }
        };
        wait_full(q_full, 0);
        wait_full(&k_full[0], 0);
        tc_fence_after();
        issue_s(0, 0);
        commit(&s_full[0]);
        issue_s(1, 0);
        commit(&s_full[1]);
        commit(&k_empty[0]);  // K_0 consumed by S0_0 / S1_0
        int st = 0;
        uint32_t ph = 0;
#pragma unroll 1
        for (int j = 0; j < P.nkv; ++j) {
          {
            const bool has_next = (j + 1 < P.nkv);
            const int nst = (st + 1 == ST) ? 0 : st + 1;
            const uint32_t nph = (st + 1 == ST) ? (ph ^ 1) : ph;
            wait_full(&v_full[st], ph);
            if (has_next) wait_full(&k_full[nst], nph);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#pragma unroll
              for (int hnd = 0; hnd < NHAND; ++hnd) {
                wait_full(&p_ready[hnd * 2 + g], j & 1);
                tc_fence_after();
                if (TRACE && hnd == 0 && el && trace && j < 64) trace[(j * 2 + g) * 8 + 6] = clock64();
                issue_pv(g, st, j > 0, hnd * (8 / NHAND), (hnd + 1) * (8 / NHAND));
              }
              if (has_next) {
                issue_s(g, nst);
                commit(&s_full[g]);
              } else {
                commit(&o_done[g]);
              }
              if (TRACE && el && trace && j < 64) trace[(j * 2 + g) * 8 + 7] = clock64();
            }
            commit(&v_empty[st]);
            if (has_next) commit(&k_empty[nst]);  // K_{j+1} fully consumed by S0/S1_{j+1}
            st = nst, ph = nph;
          }
        }
      }
    }
  } else {
    if (RPT == 1 || COOP) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    else                  asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------- softmax / correction / epilogue: thread = (row r, column slice h) of tile g ----------------
    // COOP: the same 8 warps serve BOTH query tiles, kv block by kv block (tile 0, then tile 1): while a tile waits for
    // its P.V + next Q.K^T on the tensor core, all warps work on the other tile, so the softmax part of a tile's
    // dependency chain is half as long (64 instead of 128 columns per thread) at the same total work per warp.
    constexpr int NT = COOP ? 2 : 1;     // tiles served by this warp
    const int g_first = COOP ? 0 : warp / NSW;
    const int h = COOP ? (warp >> 2) : ((warp % NSW) >> 2);  // which COLS-wide slice of the row (0 when RPT == 1)
    const int q = warp & 3;              // TMEM lane quarter
    const int r = q * 32 + lane;         // row in tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sl2 = P.sl2;
    float m_run[NT], l_run[NT];  // running (possibly stale) row max of raw scores; row sum over my columns
#pragma unroll
    for (int t = 0; t < NT; ++t) m_run[t] = -INFINITY, l_run[t] = 0.f;
    long long* const tr = (TRACE && trace && r == 0 && h == 0) ? trace : nullptr;
    // The two tiles' softmax warps share the SM's MUFU unit.  A ping-pong pair of named barriers lets only one tile
    // be in its exp2 phase at a time, which keeps the tiles in anti-phase: while tile g exponentiates, the tensor
    // core runs the other tile's P.V and next Q.K^T.  (COOP: the tiles alternate by construction.)
    if (PP && !COOP && g_first == 1) named_bar_arrive(1, SM_THREADS);  // tile 0 goes first

    for (int j = 0; j < P.nkv; ++j) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int g = g_first + t;
        const uint32_t tS = TM_S + g * 128 + h * COLS + lane_off;        // my score columns
        const uint32_t tP = TM_S + g * 128 + h * (COLS / 2) + lane_off;  // my P values, bf16 pairs
        const uint32_t tO = TM_O + g * 128 + h * COLS + lane_off;        // my O columns
        const uint32_t pbar0 = PAIR ? mapa_u32(smem_u32(&p_ready[g]), 0) : smem_u32(&p_ready[g]);  // + 16 B per instalment
        auto arrive_p = [&](int hnd) {
          if (PAIR) mbar_arrive_cluster_relaxed(pbar0 + hnd * 16);
          else      mbar_arrive(&p_ready[hnd * 2 + g]);
        };
        // RPT == 2: exchange slots with the thread that holds the other half of this row
        const uint32_t x_mine = smem_u32(smem + Cfg::XCH_OFF) + ((g * 2 + h) * 128 + r) * 4;
        const uint32_t x_other = smem_u32(smem + Cfg::XCH_OFF) + ((g * 2 + (1 - h)) * 128 + r) * 4;
        const uint32_t xbar = 3 + g * 4 + q;  // named barrier shared by the two warps that split this row quarter

        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 0] = clock64();
        uint32_t s[COLS];
        const int kv_valid = P.L - j * TKV - h * COLS;  // valid columns among mine; >= COLS except on the last block
        // row max over my columns: four independent 3-input max chains
        float bm0 = -INFINITY, bm1 = -INFINITY, bm2 = -INFINITY, bm3 = -INFINITY;
        auto max_over = [&](int i0, int i1) {
#pragma unroll
          for (int i = i0; i < i1; i += 8) {
            bm0 = fmax3(bm0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
            bm1 = fmax3(bm1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
            bm2 = fmax3(bm2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
            bm3 = fmax3(bm3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
          }
        };
        if (COLS == 128 && kv_valid >= COLS) {
          // the TMEM read of a 128-column row is the long pole of this phase: scan the first half for its maximum while
          // the second half is still in flight
          tmem_ld32(tS, &s[0]);
          tmem_ld32(tS + 32, &s[32]);
          tc_wait_ld();
          tmem_ld32(tS + 64, &s[COLS / 2]);
          tmem_ld32(tS + 96, &s[COLS / 2 + 32 < COLS ? COLS / 2 + 32 : 0]);
          max_over(0, COLS / 2);
          tc_wait_ld();
          if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 1] = clock64();
          max_over(COLS / 2, COLS);
        } else {
#pragma unroll
          for (int c = 0; c < COLS / 32; ++c) tmem_ld32(tS + c * 32, &s[c * 32]);
          tc_wait_ld();
          if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 1] = clock64();
          if (kv_valid < COLS) {
#pragma unroll
            for (int i = 0; i < COLS; ++i)
              if (i >= kv_valid) s[i] = __float_as_uint(-INFINITY);
          }
          max_over(0, COLS);
        }
        float bmax = fmaxf(fmaxf(bm0, bm1), fmaxf(bm2, bm3));
        if (RPT == 2) {  // both threads of the row must use the same reference
          const uint32_t xo = (j & 1) * 2048;
          st_shared_f32(x_mine + xo, bmax);
          named_bar_sync(xbar, 64);
          bmax = fmaxf(bmax, ld_shared_f32(x_other + xo));
        }
        const float m_new = fmaxf(m_run[t], bmax);
        // lazy rescale: keep a stale max while exp2 stays below 2^8 (same decision in both threads of a row)
        const bool need = (m_new - m_run[t]) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          float factor = 1.0f;
          if (need) {
            factor = ex2_approx((m_run[t] - m_new) * sl2);  // 0 on the first block
            m_run[t] = m_new;
            l_run[t] *= factor;
          }
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < COLS / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st32(tO + c * 32, o);
            }
          }
        }
        const float nmb = -m_run[t] * sl2;
        if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 2] = clock64();
        if (PP && !COOP) named_bar_sync(1 + g, SM_THREADS);  // wait for this tile's turn on the MUFU
        if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 3] = clock64();
        softmax_exp_store<POLY_MASK, NHAND, COLS / 16>(s, tP, sl2, nmb, l_run[t], lane, arrive_p);
        if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 4] = clock64();
        if (PP && !COOP) named_bar_arrive(2 - g, SM_THREADS);  // hand the MUFU to the other tile
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_p(NHAND - 1);
        if (TRACE && tr && j < 64) tr[(j * 2 + g) * 8 + 5] = clock64();
      }
    }

#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int g = g_first + t;
      const uint32_t tO = TM_O + g * 128 + h * COLS + lane_off;
      float l_row = l_run[t];
      if (RPT == 2) {  // row sum = my half + the partner's half (both accumulated against the same running max)
        const uint32_t x_mine = smem_u32(smem + Cfg::XCH_OFF) + ((g * 2 + h) * 128 + r) * 4;
        const uint32_t x_other = smem_u32(smem + Cfg::XCH_OFF) + ((g * 2 + (1 - h)) * 128 + r) * 4;
        const uint32_t xo = (P.nkv & 1) * 2048;
        st_shared_f32(x_mine + xo, l_run[t]);
        named_bar_sync(3 + g * 4 + q, 64);
        l_row += ld_shared_f32(x_other + xo);
      }

      // epilogue: O / l -> bf16 -> global (my COLS of the 128 head-dim columns)
      mbar_wait(&o_done[g], 0);
      tc_fence_after();
      const int l = q_row0 + g * TQ + r;
      const int b = bh / P.H;
      const int hd = bh - b * P.H;
      bf16* dst = nullptr;
      if (l < P.L) {
        if (l < P.l_split)
          dst = P.out_a + (static_cast<long long>(b) * P.l_split + l) * P.ld_a + hd * HD + h * COLS;
        else
          dst = P.out_b + (static_cast<long long>(b) * (P.L - P.l_split) + (l - P.l_split)) * P.ld_b + hd * HD + h * COLS;
      }
      const float inv = 1.0f / l_row;
#pragma unroll 1
      for (int c = 0; c < COLS / 32; ++c) {
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
  }

  tc_fence_before();
  // PAIR: the peer's shared memory and barriers stay alive until every MMA / multicast commit has landed
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(0, 512); else tmem_dealloc(0, 512);
  }
}

typedef void (*AttnKernel)(const AttnParams);
struct AttnVariant {
  AttnKernel fn, fn_traced;
  int pair;     // launched as 2-CTA clusters, 512 query rows per pair
  int threads;  // (8 * RPT + 4) * 32
  const char* what;
};
#define FB_ATTN_VARIANT(PAIR, POLY, NHAND, PP, RPT, WHAT)                                              \
  {attention_tcgen05_kernel<PAIR, POLY, NHAND, PP, RPT, false>,                                        \
   attention_tcgen05_kernel<PAIR, POLY, NHAND, PP, RPT, true>, PAIR ? 1 : 0, (8 * RPT + 4) * 32, WHAT}
#define FB_ATTN_VARIANT_COOP(PAIR, POLY, WHAT)                                                         \
  {attention_tcgen05_kernel<PAIR, POLY, 1, false, 2, false, true>,                                     \
   attention_tcgen05_kernel<PAIR, POLY, 1, false, 2, true, true>, PAIR ? 1 : 0, 12 * 32, WHAT}
// run-time selectable builds of the kernel ("attn_variant" flag); index 0 is the production default
static const AttnVariant kAttnVariants[] = {
    FB_ATTN_VARIANT(false, 0x88, 1, true, 1, "1 CTA, every 4th exp2 pair on the FMA pipe, whole-P handoff"),
    FB_ATTN_VARIANT(false, 0, 1, true, 1, "1 CTA, all exp2 on the MUFU"),
    FB_ATTN_VARIANT(false, 0x88, 4, true, 1, "1 CTA, poly 1/4, P in 4 instalments"),
    FB_ATTN_VARIANT(true, 0x88, 1, true, 1, "CTA pair, poly 1/4"),
    FB_ATTN_VARIANT(false, 0x88, 1, true, 2, "1 CTA, 2 threads per row, poly 1/4"),
    FB_ATTN_VARIANT(false, 0xAA, 1, true, 1, "1 CTA, every other exp2 pair on the FMA pipe"),
    FB_ATTN_VARIANT(false, 0x92, 1, true, 1, "1 CTA, 3 of 8 exp2 pairs on the FMA pipe"),
    FB_ATTN_VARIANT(false, 0xDA, 1, true, 1, "1 CTA, 5 of 8 exp2 pairs on the FMA pipe"),
    FB_ATTN_VARIANT(false, 0xAA, 1, false, 1, "1 CTA, poly 1/2, no MUFU ping-pong"),
    FB_ATTN_VARIANT_COOP(false, 0x88, "1 CTA, 8 softmax warps serve both tiles (half a row per thread), poly 1/4"),
    FB_ATTN_VARIANT_COOP(false, 0, "1 CTA, cooperative softmax warps, all exp2 on the MUFU"),
    FB_ATTN_VARIANT_COOP(true, 0x88, "CTA pair, cooperative softmax warps, poly 1/4"),
    FB_ATTN_VARIANT(false, 0x88, 1, false, 1, "1 CTA, poly 1/4, no MUFU ping-pong"),
    FB_ATTN_VARIANT(false, 0x08, 1, true, 1, "1 CTA, poly 1/8, ping-pong"),
    FB_ATTN_VARIANT(false, 0x08, 1, false, 1, "1 CTA, poly 1/8, no MUFU ping-pong"),
    FB_ATTN_VARIANT(false, 0x88, 4, false, 1, "1 CTA, poly 1/4, P in 4 instalments, no MUFU ping-pong"),
    FB_ATTN_VARIANT(false, 0x88, 2, true, 1, "1 CTA, poly 1/4, P in 2 instalments"),
    FB_ATTN_VARIANT(false, 0x88, 2, false, 1, "1 CTA, poly 1/4, P in 2 instalments, no MUFU ping-pong"),
};
static constexpr int kNumAttnVariants = sizeof(kAttnVariants) / sizeof(kAttnVariants[0]);

int attention_num_variants() { return kNumAttnVariants; }

int attention_init_device() {
  static DeviceOnce attr_once;
  if (attr_once.need()) {
    for (int i = 0; i < kNumAttnVariants; ++i) {
      const int bytes = static_cast<int>(kAttnVariants[i].pair ? AttnCfg<true>::SMEM : AttnCfg<false>::SMEM);
      FB_CHECK_CUDA(cudaFuncSetAttribute(kAttnVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      FB_CHECK_CUDA(cudaFuncSetAttribute(kAttnVariants[i].fn_traced, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    attr_once.done();
  }
  return 0;
}

int launch_attention(const AttnDesc& d, cudaStream_t stream) {
  FB_REQUIRE(d.q && d.k && d.v, "attention: null q/k/v");
  FB_REQUIRE(d.B > 0 && d.H > 0 && d.L > 0, "attention: empty problem");
  FB_REQUIRE(d.l_split >= 0 && d.l_split <= d.L, "attention: bad l_split");
  FB_REQUIRE(d.l_split == 0 || d.out_a != nullptr, "attention: out_a required when l_split > 0");
  FB_REQUIRE(d.l_split == d.L || d.out_b != nullptr, "attention: out_b required");
  if (int rc = attention_init_device()) return rc;
  const int variant = get_flag("attn_variant");
  FB_REQUIRE(variant >= 0 && variant < kNumAttnVariants, "attention: unknown attn_variant");
  const bool pair = kAttnVariants[variant].pair != 0;
  AttnParams P;
  memset(&P, 0, sizeof(P));
  const uint64_t bh = static_cast<uint64_t>(d.B) * d.H;
  const void* ptrs[3] = {d.q, d.k, d.v};
  CUtensorMap* maps[3] = {&P.tmap_q, &P.tmap_k, &P.tmap_v};
  for (int i = 0; i < 3; ++i) {
    FB_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "attention: q/k/v must be 16-byte aligned");
    // pair mode: each CTA stages 64 of the 128 kv rows of a K block (V: all 128 kv rows x 64 of the 128 columns)
    const uint32_t box_rows = (pair && i == 1) ? TKV / 2 : TQ;
    int rc = encode_tmap_3d(maps[i], ptrs[i], HD, d.L, bh, HD * 2, static_cast<uint64_t>(d.L) * HD * 2, 64, box_rows, 1);
    if (rc) return rc;
  }
  P.B = d.B, P.H = d.H, P.L = d.L, P.l_split = d.l_split;
  P.q_pairs = pair ? (d.L + 4 * TQ - 1) / (4 * TQ) : (d.L + 2 * TQ - 1) / (2 * TQ);
  P.nkv = (d.L + TKV - 1) / TKV;
  P.sl2 = d.scale * 1.4426950408889634f;
  P.out_a = d.out_a, P.out_b = d.out_b, P.ld_a = d.ld_a, P.ld_b = d.ld_b;
  P.trace = d.trace;
  const int grid = static_cast<int>(bh) * P.q_pairs * (pair ? 2 : 1);
  ProfScope _ps(KK_ATTN, 4.0 * bh * static_cast<double>(d.L) * d.L * HD, 4.0 * bh * d.L * HD * 2.0, stream);
  count_launch(KK_ATTN);
  const AttnKernel kern = d.trace ? kAttnVariants[variant].fn_traced : kAttnVariants[variant].fn;
  FB_CHECK_CUDA(launch_ex(kern, dim3(grid), dim3(kAttnVariants[variant].threads), pair ? AttnCfg<true>::SMEM : AttnCfg<false>::SMEM, stream,
                          pair ? 2 : 1, get_flag("pdl") != 0, P));
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
