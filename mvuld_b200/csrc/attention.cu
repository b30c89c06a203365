// Fused attention for sm_100a (tcgen05 + TMEM + TMA), two flavours of one kernel:
//
//  MODE_SWIN  SwinV2 scaled-cosine window attention, swin_transformer_v2.py:155-176:
//             S = q^ k^T   (q^ already L2-normalised, scaled by exp(min(logit_scale, ln 100)) * log2 e, fp16)
//               + 16*sigmoid(cpb_mlp(table))[relative_position_index]   (table precomputed once, indexed analytically)
//               + shifted-window mask (0 / -100, computed from token coordinates, never materialised)
//             softmax, P V, output written token-major with window_reverse + inverse cyclic shift folded in.
//  MODE_SEQ   RoBERTa self-attention as called from unixcoder.py:36: softmax(q k^T / sqrt(hd) + key-pad mask) v.
//             Keys >= len are skipped (their additive -10000 underflows to exactly 0 in fp32 for valid queries).
//
// One CTA per (window-or-sequence, head).  K and V of that head stay resident in shared memory; the CTA walks all
// query tiles of 128 rows.  Everything between QK^T and PV stays on chip and never touches shared memory:
//   S (fp32) and O (fp32) are TMEM accumulators, and P (bf16) goes back into TMEM with tcgen05.st and is consumed as
//   the A operand of the PV MMA straight from there (V is the MN-major B operand, so no transpose is ever written).
// Warp roles: warp 0 TMA producer; warps 1 / 2 PV issuers of softmax group 0 / 1 (one thread each, blocking on that
// group's barriers so a ready P is consumed at once); warp 3 QK^T issuer of both groups; warps 4-7 / 8-11 two softmax
// warpgroups (thread == query row == TMEM lane, no cross-thread reductions) that work on alternate query tiles.
// TMEM columns of group g (256 each): S [0, KT) | P buffers [KT, KT + NPB * KT/2) | O [256 - HD, 256).
// Split remainder tile (ws = 28: 784 tokens = 6 full query tiles + 16 rows): a thread-per-row sweep would spend a whole
// tile-time on 16 rows, so the 16 queries are REPLICATED over the 8 16-lane groups of the tile and lane group u only
// exponentiates key tile u (P = 0 elsewhere): O'[16u + q] accumulates key tile u's contribution alone, and the 8 partial
// (O', l) pairs of a query are summed through shared memory.  Exact under the constant softmax reference (all partials
// share it), so only heads on that path take it; a warp runs 2 sweeps instead of 7.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace mv {

#ifndef MV_ATT_POLL_NS
#define MV_ATT_POLL_NS 40
#endif
constexpr int ATT_THREADS = 384;
constexpr int ATT_BM = 128;          // query rows per tile (= TMEM lanes)
constexpr int ATT_MAX_KT = 8;        // max kv tiles resident
constexpr int ATT_REGS_CTRL = 88;      // setmaxnreg for warps 0-3 (128 threads)
constexpr int ATT_REGS_SOFTMAX = 208;  // 8 softmax warps: 128*88 + 256*208 = 64512 = 384 threads * 168 regs at launch

enum { MODE_SWIN = 0, MODE_SEQ = 1 };

#ifdef MV_ATT_TRACE
// debug timeline: (tag, g, t, j, clock) records of block 0 -- writer threads (3 MMA issuers, row 0 of each softmax
// group), each with a private region and counter (no atomics: a store costs the writer nothing but its issue slot)
__device__ long long g_att_trace[5 * 2 * 4096];
__device__ unsigned int g_att_trace_n[5];
__device__ __forceinline__ void att_trace(unsigned int& i, int tag, int g, int t, int j) {
  const long long clk = clock64();
  if (blockIdx.x != 0) return;
  const int w = tag >= 12 ? 2 + g : (tag >= 10 ? 4 : g);
  if (i < 4096u) {
    g_att_trace[(w * 4096 + i) * 2] = ((long long)tag << 48) | ((long long)g << 32) | ((long long)t << 16) | j;
    g_att_trace[(w * 4096 + i) * 2 + 1] = clk;
  }
  ++i;
  g_att_trace_n[w] = i;
}
#define ATT_TRACE(tag, g, t, j) att_trace(trace_i, tag, g, t, j)
// three-group kernel: writer slot = group (0..2), 3 = QK^T issuer, 4 = PV issuer
__device__ __forceinline__ void att_trace3(unsigned int& i, int slot, int tag, int g, int t, int j) {
  const long long clk = clock64();
  if (blockIdx.x != 0) return;
  if (i < 4096u) {
    g_att_trace[(slot * 4096 + i) * 2] = ((long long)tag << 48) | ((long long)g << 32) | ((long long)t << 16) | j;
    g_att_trace[(slot * 4096 + i) * 2 + 1] = clk;
  }
  ++i;
  g_att_trace_n[slot] = i;
}
#define ATT_TRACE3(slot, tag, g, t, j) att_trace3(trace_i, slot, tag, g, t, j)
#else
#define ATT_TRACE(tag, g, t, j)
#define ATT_TRACE3(slot, tag, g, t, j)
#endif

struct AttnParams {
  int Nq, Nkv;            // tokens per window / sequence
  int nH;                 // heads
  // MODE_SWIN
  const float* bias_rev;  // [nH, (2ws-1)^2] fp32, = 16*sigmoid(.)*log2e, w-axis reversed (see cpb kernel)
  const float* bias_max;  // [nH] max of the head's table (softmax reference bound)
  const float* q_norm;    // [nH] |q^| of the head (= exp(min(logit_scale, ln 100)) * log2 e), or null: enables the
                          // fixed softmax reference for heads whose whole logit range fits fp32 (see the kernel)
  int H, W, shift;        // token grid and cyclic shift of this block
  int C;                  // channels (= nH * HD)
  // MODE_SEQ
  const int* kv_len;      // [B] valid keys per sequence
  // packed sequences (several short sequences per row, block-diagonal attention): token (b, i) attends to keys
  // [seg_lo[b * Nq + i], seg_hi[b * Nq + i]) of its own row; null = one sequence per row
  const int* seg_lo;
  const int* seg_hi;
  // optional, packed only: key-tile range [tile_lo, tile_hi) (units of KT keys) that the 128-row query tile (b, t)
  // needs, [B, ceil(Nq / 128)]: short lines touch one or two of the four key tiles
  const int* tile_lo;
  const int* tile_hi;
  void* out;              // bf16 [tokens, C]
  float* lse;             // optional (training): log2-domain log-sum-exp of every row, [windows * nH, Nq]
  int no_split;           // debug: disable the split remainder tile (MVULD_ATT_NOSPLIT)
};

template <int HD>
struct AttnLayout {
  static constexpr int ROW_BYTES = HD * 2;                    // 64 (SW64) or 128 (SW128)
  static constexpr int SWZ = ROW_BYTES;                       // swizzle span == row
  static constexpr int LAYOUT = (HD == 32) ? 4 : 2;           // UMMA layout_type
  static constexpr int SBO = 8 * ROW_BYTES;                   // 8-row core-matrix group
  static constexpr int Q_BYTES = ATT_BM * ROW_BYTES;
};

constexpr int ATT_SPLIT_ROWS = 16;                             // query rows of a split remainder tile
constexpr int ATT_MERGE_LD = 33;                               // padded row of the merge buffer (floats)
constexpr int ATT_MERGE_BYTES = ATT_BM * (ATT_MERGE_LD + 1) * 4;   // partial O' [128][33] + partial l [128]
__host__ __device__ constexpr int att_smem_bytes(int HD, int KT, int nkt, int table_floats, int merge_bytes = 0) {
  return 2 * nkt * KT * HD * 2      // K, V
         + 2 * ATT_BM * HD * 2      // Q x2
         + ((table_floats * 4 + 1023) / 1024) * 1024 + 512 /*barriers*/ + merge_bytes + 1024 /*align*/;
}

// Row stride of the bias table in shared memory.  Lane l of a warp owns query slot i0 + l, i.e. (hi, wi) walks a
// window row and then wraps to the next one; its table address is base(hi) - wi + wj.  With stride S = -WS (mod 32)
// the wrap continues the same descending bank sequence, so the 32 lanes always hit 32 distinct banks.
__host__ __device__ constexpr int att_tab_stride(int ws) {
  int s = 2 * ws - 1;
  while ((s + ws) % 32 != 0) ++s;
  return s;
}

template <int MODE, int HD, int WS, int KT, bool QK_FP16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQ16, AttnParams p) {
  using L = AttnLayout<HD>;
  constexpr int SIDE = 2 * WS - 1;
  constexpr int TSTRIDE = att_tab_stride(WS);
  constexpr int TBL = (MODE == MODE_SWIN) ? SIDE * TSTRIDE : 0;          // shared-memory floats (padded rows)
  constexpr int ROWS_PER_TILE = (MODE == MODE_SWIN) ? KT / WS : 1;
  constexpr int SPLIT = WS - WS / 2;                 // first column / row of the "shifted-in" band
  constexpr int NSEG = (MODE == MODE_SWIN) ? ROWS_PER_TILE * 2 : 4;     // independent max chains
  constexpr int NCH = KT / 8;                        // 8-column chunks
  constexpr int PW = KT / 2;                         // 32-bit TMEM columns of one P tile (two bf16 per cell)
  constexpr int NPB = (KT + 2 * PW + HD <= 256) ? 2 : 1;                // P buffers per group
  constexpr int COL_P = KT, COL_O = 256 - HD;
  constexpr bool CAN_SPLIT = (MODE == MODE_SWIN) && HD == 32 && WS * WS % ATT_BM == ATT_SPLIT_ROWS;
  static_assert(MODE != MODE_SWIN || KT % WS == 0, "kv tile must hold whole window rows");
  static_assert(KT % 16 == 0 && KT <= 128, "kv tile");
  static_assert(KT + NPB * PW + HD <= 256, "TMEM budget of one softmax group");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  const int nkt_all = (p.Nkv + KT - 1) / KT;
  const int nq = (p.Nq + ATT_BM - 1) / ATT_BM;
  const int bh = blockIdx.x;            // (window or sequence) * nH + head
  const int head = bh % p.nH;
  const int bwin = bh / p.nH;

  int nkt = nkt_all;
  int kv_valid = p.Nkv;
  if (MODE == MODE_SEQ) {
    kv_valid = p.kv_len[bwin];
    kv_valid = kv_valid < 1 ? 1 : (kv_valid > p.Nkv ? p.Nkv : kv_valid);
    nkt = (kv_valid + KT - 1) / KT;
  }

  // key tiles query tile t has to visit (all of them unless the packed layout narrows it)
  auto kv_range = [&](int t, int& jlo, int& jhi) {
    jlo = 0;
    jhi = nkt;
    if (MODE == MODE_SEQ && p.tile_lo != nullptr) {
      jlo = __ldg(p.tile_lo + bwin * nq + t);
      jhi = __ldg(p.tile_hi + bwin * nq + t);
      jlo = max(0, min(jlo, nkt - 1));
      jhi = max(jlo + 1, min(jhi, nkt));
    }
  };

  // split remainder tile (see the header): last query tile holds <= 16 rows, at most 8 key tiles, constant-reference head
  bool split = false;
  if (CAN_SPLIT && p.q_norm != nullptr && !p.no_split) {
    const int rem = p.Nq - (nq - 1) * ATT_BM;
    split = nq >= 2 && rem > 0 && rem <= ATT_SPLIT_ROWS && nkt <= ATT_BM / ATT_SPLIT_ROWS &&
            (2.0f * __ldg(p.q_norm + head) + __ldg(p.bias_max + head) <= 100.0f);
  }

  uint8_t* sK = smem;
  uint8_t* sV = sK + nkt_all * KT * L::ROW_BYTES;
  uint8_t* sQ = sV + nkt_all * KT * L::ROW_BYTES;
  float* sTab = reinterpret_cast<float*>(sQ + 2 * L::Q_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sTab) + ((TBL * 4 + 1023) / 1024) * 1024);
  uint64_t* k_full = bars;                    // [ATT_MAX_KT]
  uint64_t* v_full = k_full + ATT_MAX_KT;     // [ATT_MAX_KT]
  uint64_t* q_full = v_full + ATT_MAX_KT;     // [2]
  uint64_t* q_empty = q_full + 2;             // [2]
  uint64_t* s_full = q_empty + 2;             // [2]
  uint64_t* s_free = s_full + 2;              // [2]
  uint64_t* o_free = s_free + 2;              // [2]
  uint64_t* p_full = o_free + 2;              // [2][2]  (group, buffer)
  uint64_t* pv_done = p_full + 4;             // [2][2]
  uint64_t* turn = pv_done + 4;               // [2]  exponential sweeps of the two groups alternate (ping-pong)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);
  float* sMerge = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [128][33] O', then [128] l

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef MV_ATT_TRACE
  unsigned int trace_i = 0;
#endif

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    if (CAN_SPLIT) prefetch_tmap(&tmQ16);
    for (int i = 0; i < ATT_MAX_KT; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1);
      mbar_init(&q_empty[g], 1);
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&o_free[g], 128);
      mbar_init(&turn[g], 128);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&p_full[2 * g + b], 128);
        mbar_init(&pv_done[2 * g + b], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (MODE == MODE_SWIN) {
    // stage this head's bias table (12 KB for ws=28) -- plain coalesced loads
    const float* src = p.bias_rev + (size_t)head * SIDE * SIDE;
    for (int i = threadIdx.x; i < SIDE * SIDE; i += ATT_THREADS) sTab[(i / SIDE) * TSTRIDE + (i % SIDE)] = __ldg(src + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dec<ATT_REGS_CTRL>();                  // producer / MMA warps need few registers
    if (warp == 0) {
      // =========================================== TMA producer ===========================================
      if (lane == 0) {
        auto load_q = [&](int g, int t) {
          mbar_arrive_expect_tx(&q_full[g], L::Q_BYTES);
          if (CAN_SPLIT && split && t == nq - 1) {
            // the 16 remainder queries, once per 16-lane group (1 KB pieces keep the swizzle phase of a 128-row box)
            for (int u = 0; u < ATT_BM / ATT_SPLIT_ROWS; ++u)
              tma_load_3d(sQ + g * L::Q_BYTES + u * ATT_SPLIT_ROWS * L::ROW_BYTES, &tmQ16, &q_full[g], 0, t * ATT_BM, bh);
          } else {
            tma_load_3d(sQ + g * L::Q_BYTES, &tmQ, &q_full[g], 0, t * ATT_BM, bh);
          }
        };
        load_q(0, 0);
        for (int j = 0; j < nkt; ++j) {
          mbar_arrive_expect_tx(&k_full[j], KT * L::ROW_BYTES);
          tma_load_3d(sK + j * KT * L::ROW_BYTES, &tmK, &k_full[j], 0, j * KT, bh);
          if (j == 0 && nq > 1) load_q(1, 1);
          mbar_arrive_expect_tx(&v_full[j], KT * L::ROW_BYTES);
          tma_load_3d(sV + j * KT * L::ROW_BYTES, &tmV, &v_full[j], 0, j * KT, bh);
        }
        uint32_t ph[2] = {0, 0};
        for (int t = 2; t < nq; ++t) {
          const int g = t & 1;
          mbar_wait(&q_empty[g], ph[g], 10);
          ph[g] ^= 1;
          load_q(g, t);
        }
      }
    } else if (warp == 3) {
      // ======================================= QK^T issuer (both groups) =======================================
      // S(t, j+1) may be issued as soon as the group has pulled S(t, j) into registers (s_free); it is not needed
      // before that group finishes tile j, so one polling thread serves both groups.
      if (lane == 0) {
        constexpr uint32_t idesc_s =
            make_idesc_bf16(ATT_BM, KT, 0, 0) & ~((QK_FP16 ? 1u : 0u) * ((1u << 7) | (1u << 10)));
        const uint32_t q0 = smem_u32(sQ), k0 = smem_u32(sK);
        int s_t[2] = {0, 1}, s_j[2] = {0, 0}, s_count[2] = {0, 0};
        int s_lo[2] = {0, 0}, s_hi[2] = {nkt, nkt};
        if (nq > 0) kv_range(0, s_lo[0], s_hi[0]);
        if (nq > 1) kv_range(1, s_lo[1], s_hi[1]);
        s_j[0] = s_lo[0];
        s_j[1] = s_lo[1];
        uint32_t ph_q[2] = {0, 0};
        long long t_last = clock64();
        while (s_t[0] < nq || s_t[1] < nq) {
          bool progress = false;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (s_t[g] >= nq) continue;
            const int j = s_j[g];
            bool ready = (j != s_lo[g]) || mbar_test_wait(&q_full[g], ph_q[g]);
            if (ready && s_count[g] > 0) ready = mbar_test_wait(&s_free[g], (s_count[g] - 1) & 1);
            if (ready && (s_t[g] == g || MODE == MODE_SEQ)) ready = mbar_test_wait(&k_full[j], 0);
            if (!ready) continue;
            ATT_TRACE(10, g, s_t[g], j);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint64_t ad = make_smem_desc(q0 + g * L::Q_BYTES + k * 32, 16, L::SBO, L::LAYOUT);
              const uint64_t bd = make_smem_desc(k0 + j * KT * L::ROW_BYTES + k * 32, 16, L::SBO, L::LAYOUT);
              umma_ss(tmem_base + g * 256, ad, bd, idesc_s, k != 0);
            }
            umma_commit(&s_full[g]);
            ATT_TRACE(11, g, s_t[g], j);
            ++s_count[g];
            if (j == s_lo[g]) ph_q[g] ^= 1;
            if (j + 1 == s_hi[g]) {
              umma_commit(&q_empty[g]);          // last read of this Q tile is in flight
              s_t[g] += 2;
              if (s_t[g] < nq) kv_range(s_t[g], s_lo[g], s_hi[g]);
              s_j[g] = s_lo[g];
            } else {
              s_j[g] = j + 1;
            }
            progress = true;
          }
          if (progress) {
            t_last = clock64();
          } else {
            __nanosleep(MV_ATT_POLL_NS);          // a hot poll would take issue slots from the softmax warps of this scheduler
            if (clock64() - t_last > MV_WATCHDOG_CYCLES) __trap();
          }
        }
      }
    } else {
      // ========================================= PV issuer of group g =========================================
      // uniform control flow, one elected lane issues, descriptor low word advanced by adds (see attn_swin3_kernel)
      const int g = warp - 1;
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, HD, 0, 1);
      const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 16, L::SBO, L::LAYOUT) >> 32);
      const uint32_t v_lo0 = (uint32_t)make_smem_desc(smem_u32(sV), 16, L::SBO, L::LAYOUT);
      const uint32_t tO = tmem_base + g * 256 + COL_O;
      const bool leader = elect_one();
      uint32_t n = 0, ph_o = 0;
      for (int t = g; t < nq; t += 2) {
        int jlo, jhi;
        kv_range(t, jlo, jhi);
        for (int j = jlo; j < jhi; ++j, ++n) {
          const int b = n % NPB;
          mbar_wait(&p_full[2 * g + b], (n / NPB) & 1, 22);
          if (j == jlo && t != g) {
            mbar_wait(&o_free[g], ph_o, 23);
            ph_o ^= 1;
          }
          if (t == g || MODE == MODE_SEQ) mbar_wait(&v_full[j], 0, 24);
          ATT_TRACE(12, g, t, j);
          tc_fence_after();
          if (leader) {
            const uint32_t tP = tmem_base + g * 256 + COL_P + b * PW;
            const uint32_t v_lo = v_lo0 + (uint32_t)j * (KT * L::ROW_BYTES >> 4);
#pragma unroll
            for (int s = 0; s < KT / 16; ++s)
              umma_ts(tO, tP + s * 8, ((uint64_t)desc_hi << 32) | (v_lo + s * (16 * L::ROW_BYTES >> 4)), idesc_pv,
                      (j != jlo) || (s != 0));
            umma_commit(&pv_done[2 * g + b]);
          }
          __syncwarp();
          ATT_TRACE(13, g, t, j);
        }
      }
    }
  } else {
    // ========================================= softmax warpgroups =========================================
    reg_inc<ATT_REGS_SOFTMAX>();
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row within the query tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 256 + lane_off;
    const uint32_t tO = tS + COL_O;
    uint32_t ph_s = 0;
    uint32_t n = 0;                                          // kv tiles processed by this group (all query tiles)

    // window coordinates (MODE_SWIN)
    const int nWw = (MODE == MODE_SWIN) ? p.W / WS : 1;
    const int wr = (MODE == MODE_SWIN) ? (bwin % ((p.H / WS) * nWw)) / nWw : 0;
    const int wc = (MODE == MODE_SWIN) ? (bwin % ((p.H / WS) * nWw)) % nWw : 0;
    const bool rowflag = (MODE == MODE_SWIN) && p.shift > 0 && (wr == p.H / WS - 1);
    const bool colflag = (MODE == MODE_SWIN) && p.shift > 0 && (wc == nWw - 1);
    const float NEG100 = -100.0f * 1.4426950408889634f;
    const float bmax = (MODE == MODE_SWIN) ? __ldg(p.bias_max + head) : 0.f;
    // Cosine attention bounds every raw score by |q^| |k^| = q_norm, and the bias lies in [0, bmax]: when
    // 2 q_norm + bmax <= 100 (log2 units) the constant C = q_norm + bmax is a softmax reference under which the largest
    // term of any row is >= 2^-100 and none exceeds 1 -- exact (the reference cancels in O / l), with no pass over the
    // scores for a maximum and no rescaling of O, ever.  (Masked terms, -100 log2 e below the rest, flush to 0 either
    // way.)  Heads with a larger logit scale take the running-maximum path below.  Uniform per CTA.
    float q_norm = INFINITY;
    if (MODE == MODE_SWIN && p.q_norm != nullptr) q_norm = __ldg(p.q_norm + head);
    const bool fixed_ref = (MODE == MODE_SWIN) && (2.0f * q_norm + bmax <= 100.0f);

    // PV(m) of this group has retired (its P buffer is free again, O includes it)
    auto wait_pv = [&](uint32_t m) { mbar_wait(&pv_done[2 * g + (m % NPB)], (m / NPB) & 1, 31); };

    for (int t = g; t < nq; t += 2) {
      // split remainder tile: lane group u = r / 16 holds query t * 128 + r % 16 and owns key tile u
      const bool split_t = CAN_SPLIT && split && (t == nq - 1);
      const int i = t * ATT_BM + (split_t ? (r & (ATT_SPLIT_ROWS - 1)) : r);   // slot inside the window / position in the sequence
      int hi = 0, wi = 0;
      if (MODE == MODE_SWIN) {
        hi = i / WS;
        wi = i - hi * WS;
        if (hi > WS - 1) hi = WS - 1;                        // rows past the window only exist as padding
      }
      const bool ri = hi >= SPLIT, ci = wi >= SPLIT;
      // packed sequences: this query row's key range (padding rows attend to themselves so that they stay finite)
      int k_lo = 0, k_hi = 0x7fffffff;
      if (MODE == MODE_SEQ && p.seg_lo != nullptr) {
        const int ii = i < p.Nq ? i : p.Nq - 1;
        k_lo = __ldg(p.seg_lo + (size_t)bwin * p.Nq + ii);
        k_hi = __ldg(p.seg_hi + (size_t)bwin * p.Nq + ii);
      }
      float m_run = -INFINITY;
      uint64_t ls[2] = {0ull, 0ull};                         // row sum as two packed fp32x2 accumulators
      int jlo, jhi;
      kv_range(t, jlo, jhi);
      if (MODE == MODE_SEQ && p.seg_lo != nullptr && i >= kv_valid) {   // tail padding: a key this tile does visit
        k_lo = jlo * KT;
        k_hi = k_lo + 1;
      }

      for (int j = jlo; j < jhi; ++j, ++n) {
        const int ncols = min(KT, kv_valid - j * KT);        // valid kv columns in this tile
        if (CAN_SPLIT && split_t && (j >> 1) != quarter) {
          // split tile, key tile owned by another warp's lanes: keep the S / P hand-shakes going and contribute P = 0
          mbar_wait(&s_full[g], ph_s, 30);
          ph_s ^= 1;
          mbar_arrive(&s_free[g]);
          if (n >= (uint32_t)NPB) wait_pv(n - NPB);
          tc_fence_after();
          const uint32_t tPz = tS + COL_P + (n % NPB) * PW;
          uint32_t zero[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) zero[q] = 0u;
          constexpr int Z32 = (PW / 32) * 32, Z16 = Z32 + ((PW - Z32) / 16) * 16;
#pragma unroll
          for (int c0 = 0; c0 < Z32; c0 += 32) tmem_st32p(tPz + c0, zero);
          if (Z16 > Z32) tmem_st16p(tPz + Z32, zero);
          if (PW > Z16) tmem_st8p(tPz + Z16, zero);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[2 * g + (n % NPB)]);
          continue;
        }

        // per-segment additive constants (shift mask) and bias-table row bases
        float cseg[NSEG];
        const float* tb[ROWS_PER_TILE];
#pragma unroll
        for (int rr = 0; rr < ROWS_PER_TILE; ++rr) {
          if (MODE == MODE_SWIN) {
            int hj = j * ROWS_PER_TILE + rr;
            if (hj > WS - 1) hj = WS - 1;
            const bool rdiff = rowflag && ((hj >= SPLIT) != ri);
            cseg[2 * rr] = (rdiff || (colflag && ci)) ? NEG100 : 0.f;           // wj <  SPLIT
            cseg[2 * rr + 1] = (rdiff || (colflag && !ci)) ? NEG100 : 0.f;      // wj >= SPLIT
            tb[rr] = sTab + (hi - hj + WS - 1) * TSTRIDE + (WS - 1 - wi);
          } else {
            cseg[0] = 0.f;
            tb[0] = nullptr;
          }
        }

        // ---- pull the whole S row into registers, then hand the TMEM buffer back to the QK^T issuer ----
        if (r == 0) ATT_TRACE(0, g, t, j);
        mbar_wait(&s_full[g], ph_s, 30);
        ph_s ^= 1;
        tc_fence_after();
        if (r == 0) ATT_TRACE(1, g, t, j);
        uint32_t sv[KT];
#pragma unroll
        for (int c0 = 0; c0 < KT; c0 += 32) {
          if (c0 + 32 <= KT) tmem_ld32p(tS + c0, sv + c0);
          else tmem_ld16p(tS + c0, sv + c0);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[g]);
        if (r == 0) ATT_TRACE(2, g, t, j);

        if (ncols < KT) {
#pragma unroll
          for (int c = 0; c < KT; ++c) sv[c] = (c < ncols) ? sv[c] : 0xff800000u;   // -inf
        }
        if (CAN_SPLIT && split_t && (r >> 4) != j) {         // the other lane group of this warp owns key tile j
#pragma unroll
          for (int c = 0; c < KT; ++c) sv[c] = 0xff800000u;
        }
        if (MODE == MODE_SEQ && p.seg_lo != nullptr) {
          // block-diagonal mask of packed sequences: keys outside [k_lo, k_hi) of this row are -inf
          const unsigned rel0 = (unsigned)(j * KT - k_lo), span = (unsigned)(k_hi - k_lo);
#pragma unroll
          for (int c = 0; c < KT; ++c) sv[c] = (rel0 + (unsigned)c < span) ? sv[c] : 0xff800000u;
        }
        // sweep 1: reference point of the tile = max over the RAW scores per mask segment + segment constant + the
        // head's largest bias.  It bounds the true row maximum from above by at most (max - min) of the bias table
        // (<= 16 log2 e), so every 2^(.) below is <= 2^8 and the largest term of a row is >= 2^-32: an exact softmax
        // (the reference point cancels in O / l) that does not touch the bias table before the exponentials.
        if (fixed_ref) {
          m_run = q_norm + bmax;
        } else {
        float segmax[NSEG];
#pragma unroll
        for (int s = 0; s < NSEG; ++s) segmax[s] = -INFINITY;
#pragma unroll
        for (int c = 0; c < KT; ++c) {
          const int seg = (MODE == MODE_SWIN) ? ((c / WS) * 2 + ((c % WS) >= SPLIT ? 1 : 0)) : (c & (NSEG - 1));
          segmax[seg] = fmaxf(segmax[seg], __uint_as_float(sv[c]));
        }
        float m_tile = -INFINITY;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) m_tile = fmaxf(m_tile, segmax[s] + ((MODE == MODE_SWIN) ? cseg[s] : 0.f));
        m_tile += bmax;
        if (MODE == MODE_SEQ) m_tile = fmaxf(m_tile, -1.0e30f);   // a fully masked tile (packed rows): keep the reference finite

        // running reference with lazy rescale (only when it moves by more than 2^8)
        const float m_new = fmaxf(m_run, m_tile);
        const bool need = (j > jlo) && (m_new > m_run + 8.0f);
        if (j == jlo) m_run = m_new;
        if (j > jlo && __any_sync(0xffffffffu, need)) {
          wait_pv(n - 1);                                    // O holds every PV of this query tile issued so far
          tc_fence_after();
          const float f = need ? ex2_approx(m_run - m_new) : 1.0f;
          if (need) m_run = m_new;
          const uint64_t f2 = pack2f(f, f);
          asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(ls[0]) : "l"(f2));
          asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(ls[1]) : "l"(f2));
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tO + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * f);
            tmem_st32(tO + c0, o);
          }
          tmem_st_wait();
        }
        }
        if (r == 0) ATT_TRACE(3, g, t, j);

        // sweep 2 (fused, software pipelined over 8-column chunks so no instruction waits on the one before it):
        //   stage A  bias loads of chunk st          (LDS, conflict free by the table stride)
        //   stage B  s + (cseg - m_run) + bias       (FADD2) of chunk st-1
        //   stage C  2^(.)                           (MUFU)  of chunk st-2
        //   stage D  row sum (FADD2) + bf16 pack     (F2FP)  of chunk st-3
        // Masked / out-of-range columns carry -inf and come out as exactly 0.
        float csm[NSEG];
#pragma unroll
        for (int s = 0; s < NSEG; ++s) csm[s] = ((MODE == MODE_SWIN) ? cseg[s] : 0.f) - m_run;
        float bb[3][8];
        uint32_t pw[PW];
        // P goes to TMEM buffer n % NPB (A operand of the PV MMA) once PV(n - NPB) has released it; the stores are
        // issued from inside the sweep as soon as a run of columns is complete, so only the last one is exposed
        if (n >= (uint32_t)NPB) wait_pv(n - NPB);
        tc_fence_after();
        const uint32_t tP = tS + COL_P + (n % NPB) * PW;
        // one-time stagger: group 1 starts its first sweep when group 0 is half way through its own, so the two
        // groups' MUFU-heavy sweeps and their latency-bound TMEM / barrier phases interleave instead of coinciding
        if (n == 0 && g == 1) mbar_wait(&turn[0], 0, 33);
        if (r == 0) ATT_TRACE(6, g, t, j);
#pragma unroll
        for (int st = 0; st < NCH + 3; ++st) {
          // one MUFU per slot, the other pipes' work of the neighbouring stages interleaved between them: the XU pipe
          // takes a warp instruction every 8 cycles and issue is in order, so clustered MUFUs would block the rest
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (st >= 2 && st < NCH + 2) {                   // stage C
              const int c = (st - 2) * 8 + q;
              sv[c] = __float_as_uint(ex2_approx(__uint_as_float(sv[c])));
            }
            if (MODE == MODE_SWIN && st < NCH) {             // stage A
              const int c = st * 8 + q;
              bb[st % 3][q] = tb[c / WS][c % WS];
            }
            if (st >= 1 && st < NCH + 1 && (q & 1) == 0) {   // stage B
              const int k = st - 1;
              const int c = k * 8 + q;
              const int seg = (MODE == MODE_SWIN) ? ((c / WS) * 2 + ((c % WS) >= SPLIT ? 1 : 0)) : 0;
              const int seg1 = (MODE == MODE_SWIN) ? (((c + 1) / WS) * 2 + (((c + 1) % WS) >= SPLIT ? 1 : 0)) : 0;
              uint64_t x = add2(pack2u(sv[c], sv[c + 1]), pack2f(csm[seg], csm[seg1]));
              if (MODE == MODE_SWIN) x = add2(x, pack2f(bb[k % 3][q], bb[k % 3][q + 1]));
              unpack2u(x, sv[c], sv[c + 1]);
            }
            if (st >= 3 && (q & 1) == 1) {                   // stage D
              const int c = (st - 3) * 8 + q - 1;
              ls[(q >> 1) & 1] = add2(ls[(q >> 1) & 1], pack2u(sv[c], sv[c + 1]));
              pw[c >> 1] = pack_bf16x2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
            }
          }
          if (st == NCH / 2 && n == 0 && g == 0) mbar_arrive(&turn[0]);
          if (st >= 3) {
            // words [0, 4 (st - 2)) of pw are final: flush 32-, 16- and 8-column runs as they complete
            constexpr int R32 = (PW / 32) * 32, R16 = R32 + ((PW - R32) / 16) * 16;
            const int done = 4 * (st - 2);
            if (done <= R32 && done % 32 == 0) tmem_st32p(tP + done - 32, pw + done - 32);
            else if (done == R16 && R16 > R32) tmem_st16p(tP + R32, pw + R32);
            else if (done == PW && PW > R16) tmem_st8p(tP + R16, pw + R16);
          }
        }
        if (r == 0) ATT_TRACE(4, g, t, j);
        tmem_st_wait();
        tc_fence_before();             // orders our tcgen05.ld/st before the issuer's MMA
        mbar_arrive(&p_full[2 * g + (n % NPB)]);
        if (r == 0) ATT_TRACE(7, g, t, j);
      }

      // ---- epilogue: O / l -> bf16, token-major store ----
      wait_pv(n - 1);
      tc_fence_after();
      float l0, l1, l2, l3;
      unpack2f(ls[0], l0, l1);
      unpack2f(ls[1], l2, l3);
      float lsum = (l0 + l1) + (l2 + l3);
      if (CAN_SPLIT && split_t) {
        // merge the 8 key-tile partials of each query (same softmax reference: plain sums) through shared memory
        static_assert(!CAN_SPLIT || HD == 32, "split merge assumes one 32-column O load");
        uint32_t o[32];
        tmem_ld32(tO, o);
        tmem_ld_wait();
        float* mrow = sMerge + r * ATT_MERGE_LD;
#pragma unroll
        for (int q = 0; q < 32; ++q) mrow[q] = __uint_as_float(o[q]);
        sMerge[ATT_BM * ATT_MERGE_LD + r] = lsum;
        named_bar_sync(1 + g, 128);
        const int qq = r & (ATT_SPLIT_ROWS - 1), cg = (r >> 4) * 4;     // this thread: query qq, O columns cg .. cg + 3
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        lsum = 0.f;
#pragma unroll
        for (int u = 0; u < ATT_BM / ATT_SPLIT_ROWS; ++u) {
          const float* src = sMerge + (u * ATT_SPLIT_ROWS + qq) * ATT_MERGE_LD + cg;
          a0 += src[0];
          a1 += src[1];
          a2 += src[2];
          a3 += src[3];
          lsum += sMerge[ATT_BM * ATT_MERGE_LD + u * ATT_SPLIT_ROWS + qq];
        }
        const float invs = lsum > 0.f ? 1.0f / lsum : 0.f;
        if (p.lse != nullptr && cg == 0 && i < p.Nq) p.lse[(size_t)bh * p.Nq + i] = m_run + log2f(lsum);
        if (i < p.Nq) {
          const int hl = i / WS, wl = i - hl * WS;
          int hh = wr * WS + hl + p.shift;
          if (hh >= p.H) hh -= p.H;
          int ww = wc * WS + wl + p.shift;
          if (ww >= p.W) ww -= p.W;
          const int b = bwin / ((p.H / WS) * nWw);
          const size_t orow_s = (size_t)b * p.H * p.W + (size_t)hh * p.W + ww;
          uint2 w2;
          w2.x = pack_bf16x2(a0 * invs, a1 * invs);
          w2.y = pack_bf16x2(a2 * invs, a3 * invs);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + orow_s * p.C + head * HD + cg) = w2;
        }
        tc_fence_before();
        mbar_arrive(&o_free[g]);
        continue;
      }
      const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
      if (p.lse != nullptr && i < p.Nq) p.lse[(size_t)bh * p.Nq + i] = m_run + log2f(lsum);
      size_t orow;
      if (MODE == MODE_SWIN) {
        const int hl = i / WS, wl = i - hl * WS;
        int hh = wr * WS + hl + p.shift;
        if (hh >= p.H) hh -= p.H;
        int ww = wc * WS + wl + p.shift;
        if (ww >= p.W) ww -= p.W;
        const int b = bwin / ((p.H / WS) * nWw);
        orow = (size_t)b * p.H * p.W + (size_t)hh * p.W + ww;
      } else {
        orow = (size_t)bwin * p.Nq + i;
      }
      bf16* op = reinterpret_cast<bf16*>(p.out) + orow * p.C + head * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t o[32];
        tmem_ld32(tO + c0, o);
        tmem_ld_wait();
        if (i < p.Nq) {
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[q]) * inv, __uint_as_float(o[q + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[q + 2]) * inv, __uint_as_float(o[q + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[q + 4]) * inv, __uint_as_float(o[q + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[q + 6]) * inv, __uint_as_float(o[q + 7]) * inv);
            *reinterpret_cast<uint4*>(op + c0 + q) = w;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[g]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// =========================================================================================================
// attn_swin3_kernel: SwinV2 window attention for 28x28 windows (784 tokens, head dim 32) on heads whose softmax can
// use the CONSTANT reference |q^| + max bias (2 |q^| + max bias <= 100, see attn_fwd_kernel) -- every head of a model
// whose logit scales stay below ~ln 27, the initialisation (ln 10) included.  Same arithmetic per score as
// attn_fwd_kernel<MODE_SWIN>, different schedule:
//  * THREE softmax warpgroups (12 warps, 3 per SM sub-partition instead of 2) hide the TMEM / mbarrier hand-offs that
//    left the MUFU pipe idle half of the time with two (profiles/r1_ncu_attn.md: XU 52 %, issue 49 %).
//  * Work units are (query tile t, key tile j), dealt round-robin to the groups: the 49 units of a head split 17/16/16
//    instead of 4/3 query tiles.  A constant reference needs no per-row state between key tiles, so the groups that
//    share a query tile just accumulate into the SAME O accumulator (one O buffer: the tile's writer drains it while
//    the first units of the next tile are still in their sweeps).
//  * P (bf16) overwrites the S columns it was computed from (the sweep reads S in pieces ahead of the P stores), so a
//    unit needs ONE 112-column TMEM buffer; FOUR of them rotate over the units (unit u -> buffer u % 4, group u % 3),
//    so QK^T of unit u + 3 is issued into the spare buffer while the PV product of unit u is still pending and a group
//    finds its next S waiting (with one buffer per group a third of the softmax warps' time went into that wait).
//    TMEM: 4 x 112 (S / P) + 48 (O).
//  * The row sum comes out of the tensor core: the PV product runs with N = 48, the extra 16 columns of the B operand
//    are a [1, 0, ..] block in shared memory (second MN atom of the descriptor, reached through its leading byte
//    offset), so O[:, 32] = sum_k P[q, k] of exactly the bf16 P the numerator used -- no FADD chain in the sweep and
//    no cross-group merge of partial sums.
// Warp roles: 0 TMA producer, 1 QK^T issuer, 2 PV issuer (+ TMEM owner), 3 idle, 4-15 softmax groups 0-2 (thread ==
// query row == TMEM lane).  The group that sweeps unit (t, 6) also writes tile t's output.
// =========================================================================================================
constexpr int A3_THREADS = 640;
constexpr int A3_G = 4;
constexpr int A3_KT = 112;
constexpr int A3_NT = 7;                    // query tiles == key tiles of a 784-token window
constexpr int A3_NU = A3_NT * A3_NT;
constexpr int A3_NB = 4;                    // rotating S / P buffers
constexpr int A3_NO = 48;                   // PV accumulator columns: 32 head dims + the row-sum column (+ 15 unused)
constexpr int A3_REGS_CTRL = 64;
constexpr int A3_REGS_SOFTMAX = 104;        // 128 * 64 + 512 * 104 = 61440 = 640 threads * 96 registers at launch

// processing order of the query tiles: the 16-row remainder tile first, then tiles 0 .. 5
__host__ __device__ constexpr int a3_tile(int pt) { return pt == 0 ? A3_NT - 1 : pt - 1; }
__host__ __device__ constexpr int a3_smem_bytes(int table_floats) {
  return 2 * A3_NT * A3_KT * 64 + 1024 /*ones*/ + 2 * ATT_BM * 64 + ((table_floats * 4 + 1023) / 1024) * 1024 +
         512 /*barriers*/ + ATT_MERGE_BYTES + 1024 /*align*/;
}

template <int WS>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_swin3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQ16, AttnParams p) {
  constexpr int HD = 32, KT = A3_KT;
  using L = AttnLayout<HD>;
  constexpr int SIDE = 2 * WS - 1;
  constexpr int TSTRIDE = att_tab_stride(WS);
  constexpr int TBL = SIDE * TSTRIDE;
  constexpr int ROWS_PER_TILE = KT / WS;
  constexpr int SPLIT = WS - WS / 2;
  constexpr int NSEG = ROWS_PER_TILE * 2;
  constexpr int NCH = KT / 8;
  constexpr int PW = KT / 2;
  static_assert(WS * WS == A3_NT * KT && WS * WS == (A3_NT - 1) * ATT_BM + ATT_SPLIT_ROWS, "784-token windows only");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sK = smem;
  uint8_t* sV = sK + A3_NT * KT * 64;
  uint8_t* sOnes = sV + A3_NT * KT * 64;                       // 16 keys x 64 B: column 0 = 1.0 (64 B swizzle applied)
  uint8_t* sQ = sOnes + 1024;
  float* sTab = reinterpret_cast<float*>(sQ + 2 * L::Q_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sTab) + ((TBL * 4 + 1023) / 1024) * 1024);
  uint64_t* k_full = bars;                    // [7]
  uint64_t* v_full = k_full + A3_NT;          // [7]
  uint64_t* q_full = v_full + A3_NT;          // [2]
  uint64_t* q_empty = q_full + 2;             // [2]
  uint64_t* s_full = q_empty + 2;             // [4]  S of the unit using this buffer is in TMEM
  uint64_t* p_full = s_full + A3_NB;          // [4]  P of that unit is in TMEM
  uint64_t* pv_done = p_full + A3_NB;         // [4]  its PV product retired: the buffer is free
  uint64_t* tile_done = pv_done + A3_NB;      // all seven PV products of a query tile retired
  uint64_t* o_free = tile_done + 1;           // the tile's output has been read out of the O accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  float* sMerge = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [128][33]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int head = bh % p.nH;
  const int bwin = bh / p.nH;
#ifdef MV_ATT_TRACE
  unsigned int trace_i = 0;
#endif

  const float bmax = __ldg(p.bias_max + head);
  const float q_norm = __ldg(p.q_norm + head);
  if (!(2.0f * q_norm + bmax <= 100.0f)) __trap();   // the host promised constant-reference heads (entry point contract)

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmQ16);
    for (int i = 0; i < A3_NT; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
    }
    mbar_init(tile_done, 1);
    mbar_init(o_free, 4);
    for (int b = 0; b < A3_NB; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 4);
      mbar_init(&pv_done[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  {
    const float* src = p.bias_rev + (size_t)head * SIDE * SIDE;
    for (int i = threadIdx.x; i < SIDE * SIDE; i += A3_THREADS) sTab[(i / SIDE) * TSTRIDE + (i % SIDE)] = __ldg(src + i);
    // ones block: key row k, logical 16-byte chunk 0 sits at physical chunk (k >> 1) & 3 under the 64 B swizzle
    if (threadIdx.x < 64) {
      const int k = threadIdx.x >> 2, c = threadIdx.x & 3;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (c == ((k >> 1) & 3)) val.x = 0x3f80u;                // bf16 1.0 in element 0
      *reinterpret_cast<uint4*>(sOnes + k * 64 + c * 16) = val;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dec<A3_REGS_CTRL>();
    if (warp == 0) {
      // =========================================== TMA producer ===========================================
      if (lane == 0) {
        // query tiles are processed remainder tile first (a3_tile): its seven one-warp units overlap the K / V loads
        // instead of forming a tail in which two thirds of the softmax warps idle
        auto load_q = [&](int pt) {
          const int b = pt & 1, t = a3_tile(pt);
          mbar_arrive_expect_tx(&q_full[b], L::Q_BYTES);
          if (t == A3_NT - 1) {
            for (int u = 0; u < ATT_BM / ATT_SPLIT_ROWS; ++u)
              tma_load_3d(sQ + b * L::Q_BYTES + u * ATT_SPLIT_ROWS * L::ROW_BYTES, &tmQ16, &q_full[b], 0, t * ATT_BM, bh);
          } else {
            tma_load_3d(sQ + b * L::Q_BYTES, &tmQ, &q_full[b], 0, t * ATT_BM, bh);
          }
        };
        load_q(0);
        for (int j = 0; j < A3_NT; ++j) {
          mbar_arrive_expect_tx(&k_full[j], KT * 64);
          tma_load_3d(sK + j * KT * 64, &tmK, &k_full[j], 0, j * KT, bh);
          if (j == 0) load_q(1);
          mbar_arrive_expect_tx(&v_full[j], KT * 64);
          tma_load_3d(sV + j * KT * 64, &tmV, &v_full[j], 0, j * KT, bh);
        }
        for (int pt = 2; pt < A3_NT; ++pt) {
          mbar_wait(&q_empty[pt & 1], ((pt - 2) >> 1) & 1, 10);
          load_q(pt);
        }
      }
    } else if (warp == 1) {
      // ============================================ QK^T issuer ============================================
      // The whole warp runs the (uniform) control flow and one elected lane issues: descriptors are then built on the
      // uniform datapath, and incrementally -- a descriptor assembled from scratch in a one-lane branch cost ~100
      // cycles per MMA (VIADD / SHF / LOP3 chain + R2UR), 700 cycles per unit on the serial issue path.
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, KT, 0, 0) & ~((1u << 7) | (1u << 10));   // fp16 operands
      const uint32_t hi = (uint32_t)(make_smem_desc(0, 16, L::SBO, L::LAYOUT) >> 32);
      const uint32_t q_lo = (uint32_t)make_smem_desc(smem_u32(sQ), 16, L::SBO, L::LAYOUT);
      const uint32_t k_lo = (uint32_t)make_smem_desc(smem_u32(sK), 16, L::SBO, L::LAYOUT);
      const bool leader = elect_one();
      for (int u = 0; u < A3_NU; ++u) {
        const int t = u / A3_NT, j = u - t * A3_NT, b = u % A3_NB, n = u / A3_NB;      // t: processing index of the tile
        if (j == 0) mbar_wait(&q_full[t & 1], (t >> 1) & 1, 20);
        if (t == 0) mbar_wait(&k_full[j], 0, 21);
        if (n > 0) mbar_wait(&pv_done[b], (n - 1) & 1, 22);
        ATT_TRACE3(3, 10, u % A3_G, t, j);
        tc_fence_after();
        if (leader) {
          const uint32_t a0 = q_lo + (uint32_t)(t & 1) * (L::Q_BYTES >> 4);
          const uint32_t b0 = k_lo + (uint32_t)j * (KT * 64 >> 4);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_ss(tmem_base + b * KT, ((uint64_t)hi << 32) | (a0 + k * 2), ((uint64_t)hi << 32) | (b0 + k * 2), idesc_s,
                    k != 0);
          umma_commit(&s_full[b]);
          if (j == A3_NT - 1) umma_commit(&q_empty[t & 1]);
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ============================================= PV issuer =============================================
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, A3_NO, 0, 1);
      // MN-major B: first MN atom = the 32 head dims of V, second atom (leading byte offset) = the ones block.
      // Low word = address | LBO << 16 with LBO = ones - address: moving the tile by d (16-byte units) adds
      // d * (1 - 65536) (no field overflows: shared-memory offsets stay below 2^18 bytes).
      const uint32_t hi = (uint32_t)(make_smem_desc(0, 0, L::SBO, L::LAYOUT) >> 32);
      const uint32_t v_lo = (uint32_t)make_smem_desc(smem_u32(sV), smem_u32(sOnes) - smem_u32(sV), L::SBO, L::LAYOUT);
      constexpr uint32_t STEP = 1u - 65536u;
      const bool leader = elect_one();
      for (int u = 0; u < A3_NU; ++u) {
        const int t = u / A3_NT, j = u - t * A3_NT, b = u % A3_NB, n = u / A3_NB;      // t: processing index of the tile
        mbar_wait(&p_full[b], n & 1, 23);
        if (j == 0 && t >= 1) mbar_wait(o_free, (t - 1) & 1, 24);
        if (t == 0) mbar_wait(&v_full[j], 0, 25);
        ATT_TRACE3(4, 12, u % A3_G, t, j);
        tc_fence_after();
        if (leader) {
          const uint32_t tO = tmem_base + A3_NB * KT;
          const uint32_t tP = tmem_base + b * KT;
          const uint32_t lo_j = v_lo + (uint32_t)j * (KT * 64 >> 4) * STEP;
#pragma unroll
          for (int s = 0; s < KT / 16; ++s)
            umma_ts(tO, tP + s * 8, ((uint64_t)hi << 32) | (lo_j + (uint32_t)s * (16 * 64 >> 4) * STEP), idesc_pv,
                    (j != 0) || (s != 0));
          umma_commit(&pv_done[b]);
          if (j == A3_NT - 1) umma_commit(tile_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ========================================= softmax warpgroups =========================================
    reg_inc<A3_REGS_SOFTMAX>();
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;

    const int nWw = p.W / WS;
    const int wr = (bwin % ((p.H / WS) * nWw)) / nWw;
    const int wc = (bwin % ((p.H / WS) * nWw)) % nWw;
    const bool rowflag = p.shift > 0 && (wr == p.H / WS - 1);
    const bool colflag = p.shift > 0 && (wc == nWw - 1);
    const float NEG100 = -100.0f * 1.4426950408889634f;
    const float m_ref = q_norm + bmax;

    for (int u = g; u < A3_NU; u += A3_G) {
      const int pt = u / A3_NT, j = u - pt * A3_NT;
      const int t = a3_tile(pt);
      const int buf = u % A3_NB;
      const uint32_t par = (uint32_t)(u / A3_NB) & 1u;
      const uint32_t tS = tmem_base + buf * KT + lane_off;
      const bool split_t = (t == A3_NT - 1);
      const int i = t * ATT_BM + (split_t ? (r & (ATT_SPLIT_ROWS - 1)) : r);
      int hi = i / WS;
      const int wi = i - hi * WS;
      if (hi > WS - 1) hi = WS - 1;
      const bool ri = hi >= SPLIT, ci = wi >= SPLIT;

      if (split_t && (j >> 1) != quarter) {
        // split remainder tile, key tile owned by another warp's lanes: P = 0
        mbar_wait(&s_full[buf], par, 30);
        tc_fence_after();
        uint32_t zero[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) zero[q] = 0u;
        tmem_st32p(tS, zero);
        tmem_st16p(tS + 32, zero);
        tmem_st8p(tS + 48, zero);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      } else {
        float csm[NSEG];
        const float* tb[ROWS_PER_TILE];
#pragma unroll
        for (int rr = 0; rr < ROWS_PER_TILE; ++rr) {
          const int hj = j * ROWS_PER_TILE + rr;
          const bool rdiff = rowflag && ((hj >= SPLIT) != ri);
          csm[2 * rr] = ((rdiff || (colflag && ci)) ? NEG100 : 0.f) - m_ref;          // wj <  SPLIT
          csm[2 * rr + 1] = ((rdiff || (colflag && !ci)) ? NEG100 : 0.f) - m_ref;     // wj >= SPLIT
          tb[rr] = sTab + (hi - hj + WS - 1) * TSTRIDE + (WS - 1 - wi);
        }
        if (split_t && (r >> 4) != j) {
          // the other 16-lane group of this warp owns key tile j: every term flushes to 0
#pragma unroll
          for (int s = 0; s < NSEG; ++s) csm[s] = -INFINITY;
        }

        if (r == 0) ATT_TRACE3(g, 0, g, t, j);
        mbar_wait(&s_full[buf], par, 30);
        tc_fence_after();
        if (r == 0) ATT_TRACE3(g, 1, g, t, j);
        uint32_t sv[KT];
        float bb[3][8];
        uint32_t pw[PW];
        tmem_ld32p(tS, sv);
        tmem_ld32p(tS + 32, sv + 32);
        tmem_ld_wait();
        if (r == 0) ATT_TRACE3(g, 6, g, t, j);
        // Software pipeline over 8-column chunks (see attn_fwd_kernel): A bias LDS | B adds | C ex2 | D bf16 pack.
        // S columns 64.. are pulled in while the first chunks are in flight; P words go back to TMEM eight at a time,
        // always into columns whose S values are already in registers.
#pragma unroll
        for (int st = 0; st < NCH + 3; ++st) {
          if (st == 4) tmem_ld32p(tS + 64, sv + 64);
          if (st == 8) {
            tmem_ld_wait();
            tmem_ld16p(tS + 96, sv + 96);
          }
          if (st == 12) tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (st >= 2 && st < NCH + 2) {                   // stage C
              const int c = (st - 2) * 8 + q;
              sv[c] = __float_as_uint(ex2_approx(__uint_as_float(sv[c])));
            }
            if (st < NCH) {                                  // stage A
              const int c = st * 8 + q;
              bb[st % 3][q] = tb[c / WS][c % WS];
            }
            if (st >= 1 && st < NCH + 1 && (q & 1) == 0) {   // stage B
              const int k = st - 1;
              const int c = k * 8 + q;
              const int seg = (c / WS) * 2 + ((c % WS) >= SPLIT ? 1 : 0);
              const int seg1 = ((c + 1) / WS) * 2 + (((c + 1) % WS) >= SPLIT ? 1 : 0);
              uint64_t x = add2(pack2u(sv[c], sv[c + 1]), pack2f(csm[seg], csm[seg1]));
              x = add2(x, pack2f(bb[k % 3][q], bb[k % 3][q + 1]));
              unpack2u(x, sv[c], sv[c + 1]);
            }
            if (st >= 3 && (q & 1) == 1) {                   // stage D
              const int c = (st - 3) * 8 + q - 1;
              pw[c >> 1] = pack_bf16x2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
            }
          }
          if (st >= 4 && (st & 1) == 0) tmem_st8p(tS + 4 * (st - 2) - 8, pw + 4 * (st - 2) - 8);   // words [4(st-2)-8, 4(st-2))
        }
        if (r == 0) ATT_TRACE3(g, 4, g, t, j);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
        if (r == 0) ATT_TRACE3(g, 7, g, t, j);
      }

      if (j == A3_NT - 1) {
        // ---- output of query tile t: O / l -> bf16, token-major store (window_reverse + inverse shift folded in) ----
        mbar_wait(tile_done, pt & 1, 31);
        tc_fence_after();
        const uint32_t tO = tmem_base + A3_NB * KT + lane_off;
        uint32_t o[32], o2[16];
        tmem_ld32(tO, o);
        tmem_ld16(tO + 32, o2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        const int b = bwin / ((p.H / WS) * nWw);
        if (split_t) {
          float* mrow = sMerge + r * ATT_MERGE_LD;
#pragma unroll
          for (int q = 0; q < 32; ++q) mrow[q] = __uint_as_float(o[q]);
          mrow[32] = __uint_as_float(o2[0]);
          named_bar_sync(1 + g, 128);
          const int qq = r & (ATT_SPLIT_ROWS - 1), cg = (r >> 4) * 4;   // this thread: query qq, O columns cg .. cg + 3
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, lsum = 0.f;
#pragma unroll
          for (int v8 = 0; v8 < ATT_BM / ATT_SPLIT_ROWS; ++v8) {
            const float* src = sMerge + (v8 * ATT_SPLIT_ROWS + qq) * ATT_MERGE_LD;
            a0 += src[cg];
            a1 += src[cg + 1];
            a2 += src[cg + 2];
            a3 += src[cg + 3];
            lsum += src[32];
          }
          const float invs = lsum > 0.f ? 1.0f / lsum : 0.f;
          if (p.lse != nullptr && cg == 0) p.lse[(size_t)bh * p.Nq + i] = m_ref + log2f(lsum);
          const int hl = i / WS, wl = i - hl * WS;
          int hh = wr * WS + hl + p.shift;
          if (hh >= p.H) hh -= p.H;
          int ww = wc * WS + wl + p.shift;
          if (ww >= p.W) ww -= p.W;
          const size_t orow_s = (size_t)b * p.H * p.W + (size_t)hh * p.W + ww;
          uint2 w2;
          w2.x = pack_bf16x2(a0 * invs, a1 * invs);
          w2.y = pack_bf16x2(a2 * invs, a3 * invs);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + orow_s * p.C + head * HD + cg) = w2;
        } else {
          const float lsum = __uint_as_float(o2[0]);
          const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
          if (p.lse != nullptr) p.lse[(size_t)bh * p.Nq + i] = m_ref + log2f(lsum);
          const int hl = i / WS, wl = i - hl * WS;
          int hh = wr * WS + hl + p.shift;
          if (hh >= p.H) hh -= p.H;
          int ww = wc * WS + wl + p.shift;
          if (ww >= p.W) ww -= p.W;
          const size_t orow = (size_t)b * p.H * p.W + (size_t)hh * p.W + ww;
          bf16* op = reinterpret_cast<bf16*>(p.out) + orow * p.C + head * HD;
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[q]) * inv, __uint_as_float(o[q + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[q + 2]) * inv, __uint_as_float(o[q + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[q + 4]) * inv, __uint_as_float(o[q + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[q + 6]) * inv, __uint_as_float(o[q + 7]) * inv);
            *reinterpret_cast<uint4*>(op + q) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int WS>
static int launch_attn_swin3(const void* q, const void* k, const void* v, int n_bh, const AttnParams& p,
                             cudaStream_t stream) {
  using L = AttnLayout<32>;
  constexpr int TBL = (2 * WS - 1) * att_tab_stride(WS);
  const int smem = a3_smem_bytes(TBL);
  MV_CHECK_ARG(smem <= 232448, "attention: %d B shared memory needed, 232448 available", smem);
  CUtensorMap tmQ, tmK, tmV, tmQ16;
  uint64_t dq[3] = {32, (uint64_t)p.Nq, (uint64_t)n_bh};
  uint64_t sq[2] = {64, (uint64_t)p.Nq * 64};
  uint32_t bq[3] = {32, ATT_BM, 1};
  int rc = make_tmap_16b(&tmQ, q, 3, dq, sq, bq, L::SWZ);
  if (rc) return rc;
  uint32_t bk[3] = {32, (uint32_t)A3_KT, 1};
  rc = make_tmap_16b(&tmK, k, 3, dq, sq, bk, L::SWZ);
  if (rc) return rc;
  rc = make_tmap_16b(&tmV, v, 3, dq, sq, bk, L::SWZ);
  if (rc) return rc;
  uint32_t bq16[3] = {32, ATT_SPLIT_ROWS, 1};
  rc = make_tmap_16b(&tmQ16, q, 3, dq, sq, bq16, L::SWZ);
  if (rc) return rc;
  auto kern = attn_swin3_kernel<WS>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<n_bh, A3_THREADS, smem, stream>>>(tmQ, tmK, tmV, tmQ16, p);
  MV_LAUNCH_OK();
  return 0;
}

template <int MODE, int HD, int WS, int KT, bool QK_FP16>
static int launch_attn(const void* q, const void* k, const void* v, int n_bh, const AttnParams& p,
                       cudaStream_t stream) {
  using L = AttnLayout<HD>;
  const int nkt = (p.Nkv + KT - 1) / KT;
  MV_CHECK_ARG(nkt <= ATT_MAX_KT, "attention: %d kv tiles exceed the resident maximum %d", nkt, ATT_MAX_KT);
  constexpr int TBL = (MODE == MODE_SWIN) ? (2 * WS - 1) * att_tab_stride(WS) : 0;
  const int smem = att_smem_bytes(HD, KT, nkt, TBL, (MODE == MODE_SWIN) ? ATT_MERGE_BYTES : 0);
  MV_CHECK_ARG(smem <= 232448, "attention: %d B shared memory needed, 232448 available", smem);
  CUtensorMap tmQ, tmK, tmV, tmQ16;
  uint64_t dq[3] = {(uint64_t)HD, (uint64_t)p.Nq, (uint64_t)n_bh};
  uint64_t sq[2] = {(uint64_t)HD * 2, (uint64_t)p.Nq * HD * 2};
  uint32_t bq[3] = {(uint32_t)HD, ATT_BM, 1};
  int rc = make_tmap_16b(&tmQ, q, 3, dq, sq, bq, L::SWZ);
  if (rc) return rc;
  uint64_t dk[3] = {(uint64_t)HD, (uint64_t)p.Nkv, (uint64_t)n_bh};
  uint64_t sk[2] = {(uint64_t)HD * 2, (uint64_t)p.Nkv * HD * 2};
  uint32_t bk[3] = {(uint32_t)HD, (uint32_t)KT, 1};
  rc = make_tmap_16b(&tmK, k, 3, dk, sk, bk, L::SWZ);
  if (rc) return rc;
  rc = make_tmap_16b(&tmV, v, 3, dk, sk, bk, L::SWZ);
  if (rc) return rc;
  uint32_t bq16[3] = {(uint32_t)HD, ATT_SPLIT_ROWS, 1};      // 16-row Q box of the split remainder tile
  rc = make_tmap_16b(&tmQ16, q, 3, dq, sq, bq16, L::SWZ);
  if (rc) return rc;
  auto kern = attn_fwd_kernel<MODE, HD, WS, KT, QK_FP16>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<n_bh, ATT_THREADS, smem, stream>>>(tmQ, tmK, tmV, tmQ16, p);
  MV_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Continuous-position-bias table (swin_transformer_v2.py:98-111,159,163): for every relative offset
// (dh, dw) in [-(ws-1), ws-1]^2:  16*sigmoid(W2 relu(W1 coords + b1)) * log2(e), stored per head with the w axis
// REVERSED (entry [dh_idx][x] holds dw_idx = 2ws-2-x) so a query row reads consecutive addresses as wj grows.
// Also emits the per-head maximum.  Runs once per weight version, not per forward.
// ---------------------------------------------------------------------------------------------------------
__global__ void cpb_table_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, int nH, int ws, int pretrained_ws,
                                 float* __restrict__ table_rev, float* __restrict__ table_ref) {
  const int side = 2 * ws - 1;
  const int e = blockIdx.x;                   // one table entry per block
  const int dhi = e / side, dwi = e % side;
  const float denom = (float)((pretrained_ws > 0 ? pretrained_ws : ws) - 1);
  auto coord = [&](int idx) {
    float t = (float)(idx - (ws - 1)) / denom * 8.0f;
    float a = log2f(fabsf(t) + 1.0f) / 3.0f;          // log2(8) = 3
    return t > 0.f ? a : (t < 0.f ? -a : 0.f);
  };
  const float ch = coord(dhi), cw = coord(dwi);
  extern __shared__ float hid[];              // [512]
  for (int u = threadIdx.x; u < 512; u += blockDim.x) {
    float a = w1[2 * u] * ch + w1[2 * u + 1] * cw + b1[u];
    hid[u] = a > 0.f ? a : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < nH; h += blockDim.x >> 5) {
    float acc = 0.f;
    for (int u = lane; u < 512; u += 32) acc += w2[h * 512 + u] * hid[u];
    acc = warp_sum(acc);
    if (lane == 0) {
      const float val = 16.0f / (1.0f + expf(-acc));
      table_ref[(size_t)h * side * side + e] = val;                                   // reference order, natural units
      table_rev[(size_t)h * side * side + dhi * side + (side - 1 - dwi)] = val * 1.4426950408889634f;
    }
  }
}

__global__ void table_max_kernel(const float* __restrict__ table_rev, int n, float* __restrict__ out) {
  const float* t = table_rev + (size_t)blockIdx.x * n;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, t[i]);
  m = warp_max(m);
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = sm[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sm[i]);
    out[blockIdx.x] = r;
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_cpb_table(const float* w1, const float* b1, const float* w2, int nH, int ws, int pretrained_ws,
                               float* table_rev, float* table_ref, float* table_max, cudaStream_t stream) {
  MV_CHECK_ARG(nH > 0 && ws > 1, "cpb_table: bad geometry");
  const int side = 2 * ws - 1;
  cpb_table_kernel<<<side * side, 256, 512 * sizeof(float), stream>>>(w1, b1, w2, nH, ws, pretrained_ws, table_rev,
                                                                      table_ref);
  MV_LAUNCH_OK();
  table_max_kernel<<<nH, 256, 0, stream>>>(table_rev, side * side, table_max);
  MV_LAUNCH_OK();
  return 0;
}

#ifdef MV_ATT_TRACE
extern "C" int mvuld_debug_att_trace(long long* host_out, int max_records) {
  unsigned int n[5] = {0, 0, 0, 0, 0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(n, g_att_trace_n, sizeof(n));
  static long long tmp[5 * 2 * 4096];
  cudaMemcpyFromSymbol(tmp, g_att_trace, sizeof(tmp));
  int out = 0;
  for (int w = 0; w < 5; ++w)
    for (unsigned int i = 0; i < n[w] && i < 4096u && out < max_records; ++i, ++out) {
      host_out[4 * out] = tmp[(w * 4096 + i) * 2];
      host_out[4 * out + 1] = tmp[(w * 4096 + i) * 2 + 1];
    }
  unsigned int zero[5] = {0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_att_trace_n, zero, sizeof(zero));
  return out;
}
#endif

static int swin_window_attention(const void* q, const void* k, const void* v, const float* bias_rev,
                                 const float* bias_max, const float* q_norm, void* out, float* lse, int B, int H, int W,
                                 int C, int nH, int ws, int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C == nH * 32, "swin attention: head_dim must be 32");
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0, "swin attention: window must tile the token grid");
  MV_CHECK_ARG(shift == 0 || shift == ws / 2, "swin attention: shift must be 0 or ws/2");
  AttnParams p{};
  p.Nq = p.Nkv = ws * ws;
  p.nH = nH;
  p.bias_rev = bias_rev;
  p.bias_max = bias_max;
  p.q_norm = q_norm;
  p.H = H; p.W = W; p.shift = shift; p.C = C;
  p.kv_len = nullptr;
  p.out = out;
  p.lse = lse;
  p.no_split = false;
  const int n_bh = B * (H / ws) * (W / ws) * nH;
  switch (ws) {
    case 28: return launch_attn<MODE_SWIN, 32, 28, 112, true>(q, k, v, n_bh, p, stream);
    case 14: return launch_attn<MODE_SWIN, 32, 14, 112, true>(q, k, v, n_bh, p, stream);
    case 7: return launch_attn<MODE_SWIN, 32, 7, 112, true>(q, k, v, n_bh, p, stream);
    default: return mv::fail(-1, "swin attention: window %d not instantiated (7, 14, 28)", ws);
  }
}
extern "C" int mvuld_swin_window_attention(const void* q, const void* k, const void* v, const float* bias_rev,
                                           const float* bias_max, const float* q_norm, void* out, int B, int H, int W,
                                           int C, int nH, int ws, int shift, cudaStream_t stream) {
  return swin_window_attention(q, k, v, bias_rev, bias_max, q_norm, out, nullptr, B, H, W, C, nH, ws, shift, stream);
}

static int swin_window_attention_fixed(const void* q, const void* k, const void* v, const float* bias_rev,
                                       const float* bias_max, const float* q_norm, void* out, float* lse, int B, int H,
                                       int W, int C, int nH, int ws, int shift, cudaStream_t stream);
// Training forward: the same kernels, plus the log2-domain log-sum-exp of every score row (fp32 [windows * nH, ws^2],
// window-major like q) that mvuld_swin_attention_bwd recomputes the probabilities from.  fixed != 0 selects the
// constant-reference kernel (ws == 28, caller-checked as for mvuld_swin_window_attention_fixed).
extern "C" int mvuld_swin_window_attention_train(const void* q, const void* k, const void* v, const float* bias_rev,
                                                 const float* bias_max, const float* q_norm, void* out, float* lse,
                                                 int fixed, int B, int H, int W, int C, int nH, int ws, int shift,
                                                 cudaStream_t stream) {
  MV_CHECK_ARG(lse != nullptr, "swin attention (train): lse is null");
  if (fixed) return swin_window_attention_fixed(q, k, v, bias_rev, bias_max, q_norm, out, lse, B, H, W, C, nH, ws, shift, stream);
  return swin_window_attention(q, k, v, bias_rev, bias_max, q_norm, out, lse, B, H, W, C, nH, ws, shift, stream);
}

// Same operation for launches whose heads ALL satisfy the constant-reference condition 2 |q^| + max bias <= 100 (log2
// units; the caller checks it once per weight version -- a violating head traps): the three-group kernel above.
static int swin_window_attention_fixed(const void* q, const void* k, const void* v, const float* bias_rev,
                                       const float* bias_max, const float* q_norm, void* out, float* lse, int B, int H,
                                       int W, int C, int nH, int ws, int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C == nH * 32, "swin attention: head_dim must be 32");
  MV_CHECK_ARG(ws == 28, "swin attention (constant reference): 28x28 windows only, use mvuld_swin_window_attention");
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0, "swin attention: window must tile the token grid");
  MV_CHECK_ARG(shift == 0 || shift == ws / 2, "swin attention: shift must be 0 or ws/2");
  MV_CHECK_ARG(q_norm != nullptr && bias_max != nullptr, "swin attention (constant reference): q_norm / bias_max are null");
  AttnParams p{};
  p.Nq = p.Nkv = ws * ws;
  p.nH = nH;
  p.bias_rev = bias_rev;
  p.bias_max = bias_max;
  p.q_norm = q_norm;
  p.H = H; p.W = W; p.shift = shift; p.C = C;
  p.out = out;
  p.lse = lse;
  const int n_bh = B * (H / ws) * (W / ws) * nH;
  return launch_attn_swin3<28>(q, k, v, n_bh, p, stream);
}
extern "C" int mvuld_swin_window_attention_fixed(const void* q, const void* k, const void* v, const float* bias_rev,
                                                 const float* bias_max, const float* q_norm, void* out, int B, int H,
                                                 int W, int C, int nH, int ws, int shift, cudaStream_t stream) {
  return swin_window_attention_fixed(q, k, v, bias_rev, bias_max, q_norm, out, nullptr, B, H, W, C, nH, ws, shift, stream);
}

static int seq_attention(const void* q, const void* k, const void* v, const int* kv_len, const int* seg_lo,
                         const int* seg_hi, const int* tile_lo, const int* tile_hi, void* out, float* lse, int B, int L,
                         int nH, int hd, cudaStream_t stream) {
  MV_CHECK_ARG(hd == 64, "seq attention: head_dim 64 only");
  MV_CHECK_ARG(L <= 512 && L % 8 == 0, "seq attention: L must be <= 512 and a multiple of 8");
  AttnParams p{};
  p.Nq = p.Nkv = L;
  p.nH = nH;
  p.C = nH * hd;
  p.kv_len = kv_len;
  p.seg_lo = seg_lo;
  p.seg_hi = seg_hi;
  p.tile_lo = tile_lo;
  p.tile_hi = tile_hi;
  p.out = out;
  p.lse = lse;
  return launch_attn<MODE_SEQ, 64, 1, 128, false>(q, k, v, B * nH, p, stream);
}
extern "C" int mvuld_seq_attention(const void* q, const void* k, const void* v, const int* kv_len, void* out, int B,
                                   int L, int nH, int hd, cudaStream_t stream) {
  return seq_attention(q, k, v, kv_len, nullptr, nullptr, nullptr, nullptr, out, nullptr, B, L, nH, hd, stream);
}
// training forward: also writes the log2-domain log-sum-exp of every score row (fp32 [B * nH, L]) for
// mvuld_seq_attention_bwd
extern "C" int mvuld_seq_attention_train(const void* q, const void* k, const void* v, const int* kv_len, void* out,
                                         float* lse, int B, int L, int nH, int hd, cudaStream_t stream) {
  MV_CHECK_ARG(lse != nullptr, "seq attention (train): lse is null");
  return seq_attention(q, k, v, kv_len, nullptr, nullptr, nullptr, nullptr, out, lse, B, L, nH, hd, stream);
}
extern "C" int mvuld_seq_attention_packed(const void* q, const void* k, const void* v, const int* kv_len,
                                          const int* seg_lo, const int* seg_hi, const int* tile_lo, const int* tile_hi,
                                          void* out, int B, int L, int nH, int hd, cudaStream_t stream) {
  MV_CHECK_ARG(seg_lo && seg_hi, "packed seq attention: segment bounds are null");
  MV_CHECK_ARG((tile_lo == nullptr) == (tile_hi == nullptr), "packed seq attention: give both tile bounds or neither");
  return seq_attention(q, k, v, kv_len, seg_lo, seg_hi, tile_lo, tile_hi, out, nullptr, B, L, nH, hd, stream);
}
