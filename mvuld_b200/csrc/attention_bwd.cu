// Backward of SwinV2 scaled-cosine window attention (swin_transformer_v2.py:155-176 under autograd, as trained by
// mvuld/main.py:251-300) for sm_100a: tcgen05 + TMEM + TMA, one CTA per (window, head), nothing but the gradients and
// the per-window dS matrix (for the bias-table gradient) ever leaves the chip.
//
// With P = softmax(S), S = q^ k^T + bias + mask (log2 units, q^ = s q / |q| with the logit scale folded in, k^ = k / |k|):
//   dP = dO V^T,  D = rowsum(dO o O),  G = P o (dP - D)          (= dL / d natural-unit logits)
//   dV = P^T dO,  dK^ = G^T Q^,  dQ^ = G K^,  d bias[h, rel(q, k)] += G[q, k]
// The kernel works on TRANSPOSED score tiles (TMEM lane = key, column = query), FlashAttention-backward style:
//   S^T = K^ Q^T and dP^T = V dO^T are SS MMAs into TMEM; each math thread owns one key row, recomputes
//   P^T = 2^(S^T + bias + mask - LSE[q]) from the forward's log-sum-exp and forms G^T; P^T and G^T go back into TMEM as
//   bf16 (over the S^T / dP^T columns they came from) and are the A operands of dV += P^T dO and dK^ += G^T Q^ straight
//   from there; G^T is also written to shared memory once ([key][query], the MN-major A operand of dQ^ += G K^) and to
//   global memory (bf16 [window * head, key, query]) for the bias-table reduction (mvuld_swin_bias_grad).
// Loop: key tile j (128 keys) outer, query tile i (112 queries) inner.  TMEM (512 columns): S^T 112 | dP^T 112 | dV 32 |
// dK^ 32 | dQ^ of all 7 query tiles 7 x 32 -- dQ^ accumulates over j on chip, no atomics anywhere: deterministic.
// Warp roles: 0 TMA producer, 1 MMA issuer (uniform control flow, one elected lane), 2 TMEM owner, 4-11 two math
// warpgroups that split a tile's 16-query MMA steps 4 : 3 (thread == key row == TMEM lane in both).
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int AB_THREADS = 384;
constexpr int AB_KT = 128;        // keys per tile (TMEM lanes)
constexpr int AB_QT = 112;        // queries per tile (TMEM columns)
constexpr int AB_NSTEP = AB_QT / 16;
constexpr int AB_STEPS_WG0 = 4;   // math warpgroup 0 takes query steps [0, 4), warpgroup 1 steps [4, 7)

__host__ __device__ constexpr int ab_tab_stride(int ws) {
  // lane l owns key k0 + l; its table address for a fixed query is base - kx (descending); when the key wraps to the
  // next window row the address moves by -stride + (ws - 1): stride = ws (mod 32) keeps the descending bank sequence
  int s = 2 * ws - 1;
  while ((s - ws) % 32 != 0) ++s;
  return s;
}

struct AttnBwdParams {
  int nH, H, W, shift;
  const float* bias_rev;   // [nH, (2ws-1)^2] log2 units, w axis reversed (mvuld_cpb_table)
  const float2* ld;        // [n_bh, ntok] (LSE in log2 units, D = rowsum(dO o O)) from mvuld_swin_attention_bwd_prep
  float* dq;               // [n_bh, ntok, 32] fp32: sum_k G[q, k] k^[k]
  float* dk;               // [n_bh, ntok, 32] fp32: sum_q G[q, k] q^[q]   (q^ carries the logit scale and log2 e)
  float* dv;               // [n_bh, ntok, 32] fp32
  bf16* gt;                // [n_bh, ntok, ntok_pad] G^T (key major), or null
  int ntok_pad;
};

template <int WS>
struct AbCfg {
  static constexpr int NTOK = WS * WS;
  static constexpr int NKT = (NTOK + AB_KT - 1) / AB_KT;
  static constexpr int NQT = (NTOK + AB_QT - 1) / AB_QT;
  static constexpr int SIDE = 2 * WS - 1;
  static constexpr int TSTRIDE = ab_tab_stride(WS);
  static constexpr int TAB_ROWS = SIDE + AB_QT / WS;          // rows past the table are read (and discarded) by padding queries
  static constexpr int TAB_BYTES = ((TAB_ROWS * TSTRIDE * 4 + 1023) / 1024) * 1024;
  static constexpr int K_BYTES = AB_KT * 64;                  // one [128 x 32] 16-bit operand tile
  static constexpr int Q_BYTES = AB_QT * 64;
  static constexpr int G_BYTES = 2 * AB_KT * 128;             // [128 keys][2 x 64 queries] bf16, 128-byte swizzle atoms
  static constexpr int LD_BYTES = ((NQT * AB_QT * 8 + 1023) / 1024) * 1024;
  static constexpr int SMEM = 3 * K_BYTES + 2 * 3 * Q_BYTES + G_BYTES + TAB_BYTES + LD_BYTES + 256 + 1024;
  static constexpr int COL_ST = 0, COL_DP = AB_QT, COL_DV = 2 * AB_QT, COL_DK = 2 * AB_QT + 32, COL_DQ = 2 * AB_QT + 64;
  static constexpr int TMEM_COLS = (COL_DQ + 32 * NQT <= 256) ? 256 : 512;
  static_assert(COL_DQ + 32 * NQT <= 512, "TMEM budget");
};

template <int WS>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQb,
                const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKb,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, AttnBwdParams p) {
  using Cf = AbCfg<WS>;
  constexpr int NTOK = Cf::NTOK, NKT = Cf::NKT, NQT = Cf::NQT, TS = Cf::TSTRIDE;
  constexpr int RPT = AB_QT / WS;                   // query window rows per tile
  constexpr int SPLIT = WS - WS / 2;
  constexpr float NEG100 = -100.0f * 1.4426950408889634f;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sG = smem;                               // 1024-aligned (128-byte swizzle)
  uint8_t* sKh = sG + Cf::G_BYTES;                  // fp16 K^   (A of S^T, K-major)
  uint8_t* sKb = sKh + Cf::K_BYTES;                 // bf16 K^   (B of dQ^, MN-major)
  uint8_t* sV = sKb + Cf::K_BYTES;                  // bf16 V    (A of dP^T, K-major)
  uint8_t* sQh = sV + Cf::K_BYTES;                  // [2] fp16 Q^  (B of S^T)
  uint8_t* sQb = sQh + 2 * Cf::Q_BYTES;             // [2] bf16 Q^  (B of dK^, MN-major)
  uint8_t* sdO = sQb + 2 * Cf::Q_BYTES;             // [2] bf16 dO  (B of dP^T K-major, B of dV MN-major)
  float* sTab = reinterpret_cast<float*>(sdO + 2 * Cf::Q_BYTES);
  float2* sLD = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(sTab) + Cf::TAB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLD) + Cf::LD_BYTES);
  uint64_t* kv_full = bars;          // K^ / V of key tile j landed
  uint64_t* kv_empty = bars + 1;     // every MMA of key tile j retired
  uint64_t* q_full = bars + 2;       // [2]
  uint64_t* q_empty = bars + 4;      // [2]
  uint64_t* sdp_full = bars + 6;     // S^T and dP^T of the unit are in TMEM
  uint64_t* pg_full = bars + 7;      // P^T / G^T are in TMEM and G^T in shared memory
  uint64_t* dkv_full = bars + 8;     // dV / dK^ of key tile j complete
  uint64_t* dkv_free = bars + 9;     // ... and read out
  uint64_t* dq_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int head = bh % p.nH;
  const int bwin = bh / p.nH;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQh); prefetch_tmap(&tmQb); prefetch_tmap(&tmKh); prefetch_tmap(&tmKb);
    prefetch_tmap(&tmV); prefetch_tmap(&tmdO);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pg_full, 8);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_free, 8);
    mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cf::TMEM_COLS);
    tmem_relinquish();
  }
  {
    // bias table in natural orientation (entry [dy][dx], dx = qx - kx + ws - 1) from the reversed global layout
    const float* src = p.bias_rev + (size_t)head * Cf::SIDE * Cf::SIDE;
    for (int i = threadIdx.x; i < Cf::TAB_ROWS * TS; i += AB_THREADS) {
      const int dy = i / TS, dx = i - dy * TS;
      sTab[i] = (dy < Cf::SIDE && dx < Cf::SIDE) ? __ldg(src + dy * Cf::SIDE + (Cf::SIDE - 1 - dx)) : 0.f;
    }
    const float2* ldg = p.ld + (size_t)bh * NTOK;
    for (int i = threadIdx.x; i < NQT * AB_QT; i += AB_THREADS)
      sLD[i] = i < NTOK ? __ldg(ldg + i) : make_float2(1.0e30f, 0.f);      // padding queries: P = 0, G = 0
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int n = 0;
      for (int j = 0; j < NKT; ++j) {
        if (j > 0) mbar_wait(kv_empty, (j - 1) & 1, 1);
        mbar_arrive_expect_tx(kv_full, 3 * Cf::K_BYTES);
        tma_load_3d(sKh, &tmKh, kv_full, 0, j * AB_KT, bh);
        tma_load_3d(sKb, &tmKb, kv_full, 0, j * AB_KT, bh);
        tma_load_3d(sV, &tmV, kv_full, 0, j * AB_KT, bh);
        for (int i = 0; i < NQT; ++i, ++n) {
          const int st = n & 1;
          if (n >= 2) mbar_wait(&q_empty[st], ((n >> 1) - 1) & 1, 2);
          mbar_arrive_expect_tx(&q_full[st], 3 * Cf::Q_BYTES);
          tma_load_3d(sQh + st * Cf::Q_BYTES, &tmQh, &q_full[st], 0, i * AB_QT, bh);
          tma_load_3d(sQb + st * Cf::Q_BYTES, &tmQb, &q_full[st], 0, i * AB_QT, bh);
          tma_load_3d(sdO + st * Cf::Q_BYTES, &tmdO, &q_full[st], 0, i * AB_QT, bh);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    constexpr uint32_t idesc_st = make_idesc_bf16(AB_KT, AB_QT, 0, 0) & ~((1u << 7) | (1u << 10));   // fp16 x fp16
    constexpr uint32_t idesc_dp = make_idesc_bf16(AB_KT, AB_QT, 0, 0);                               // bf16 x bf16
    constexpr uint32_t idesc_dv = make_idesc_bf16(AB_KT, 32, 0, 1);       // A from TMEM, B MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(AB_KT, 32, 1, 1);       // A MN-major (shared memory), B MN-major
    // 64-byte-swizzle operand tiles (rows of 32 16-bit elements): 8-row groups 512 B apart
    const uint32_t hi64 = (uint32_t)(make_smem_desc(0, 16, 512, 4) >> 32);
    // G^T tile: [128 keys][2 atoms of 64 queries], 128-byte swizzle; MN atoms 16 KB apart (LBO), 8-key groups 1 KB (SBO)
    const uint32_t hi128 = (uint32_t)(make_smem_desc(0, 0, 1024, 2) >> 32);
    const uint32_t g_lo = (uint32_t)make_smem_desc(smem_u32(sG), AB_KT * 128, 1024, 2);
    const uint32_t kh_lo = (uint32_t)make_smem_desc(smem_u32(sKh), 16, 512, 4);
    const uint32_t kb_lo = (uint32_t)make_smem_desc(smem_u32(sKb), 16, 512, 4);
    const uint32_t v_lo = (uint32_t)make_smem_desc(smem_u32(sV), 16, 512, 4);
    const uint32_t qh_lo = (uint32_t)make_smem_desc(smem_u32(sQh), 16, 512, 4);
    const uint32_t qb_lo = (uint32_t)make_smem_desc(smem_u32(sQb), 16, 512, 4);
    const uint32_t do_lo = (uint32_t)make_smem_desc(smem_u32(sdO), 16, 512, 4);
    const bool leader = elect_one();
    auto d64 = [&](uint32_t lo) { return ((uint64_t)hi64 << 32) | lo; };
    int n = 0;
    for (int j = 0; j < NKT; ++j) {
      mbar_wait(kv_full, j & 1, 3);
      for (int i = 0; i < NQT; ++i, ++n) {
        const int st = n & 1;
        mbar_wait(&q_full[st], (n >> 1) & 1, 4);
        if (i == 0 && j > 0) mbar_wait(dkv_free, (j - 1) & 1, 5);
        tc_fence_after();
        const uint32_t qoff = (uint32_t)st * (Cf::Q_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 2; ++k)       // head dim 32 = two K16 steps, 32 bytes apart inside the swizzled row
            umma_ss(tmem_base + Cf::COL_ST, d64(kh_lo + k * 2), d64(qh_lo + qoff + k * 2), idesc_st, k != 0);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_ss(tmem_base + Cf::COL_DP, d64(v_lo + k * 2), d64(do_lo + qoff + k * 2), idesc_dp, k != 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(pg_full, n & 1, 6);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int s = 0; s < AB_NSTEP; ++s)      // dV += P^T dO: K = 16 queries per step, B rows = those queries
            umma_ts(tmem_base + Cf::COL_DV, tmem_base + Cf::COL_ST + 16 * s, d64(do_lo + qoff + s * (16 * 64 >> 4)),
                    idesc_dv, (i != 0) || (s != 0));
#pragma unroll
          for (int s = 0; s < AB_NSTEP; ++s)      // dK^ += G^T Q^
            umma_ts(tmem_base + Cf::COL_DK, tmem_base + Cf::COL_DP + 16 * s, d64(qb_lo + qoff + s * (16 * 64 >> 4)),
                    idesc_dv, (i != 0) || (s != 0));
#pragma unroll
          for (int s = 0; s < AB_KT / 16; ++s)    // dQ^_i += G K^: K = 16 keys per step
            umma_ss(tmem_base + Cf::COL_DQ + 32 * i, ((uint64_t)hi128 << 32) | (g_lo + s * (16 * 128 >> 4)),
                    d64(kb_lo + s * (16 * 64 >> 4)), idesc_dq, (j != 0) || (s != 0));
          umma_commit(&q_empty[st]);
          if (i == NQT - 1) {
            umma_commit(dkv_full);
            umma_commit(kv_empty);
          }
        }
        __syncwarp();
      }
    }
    if (leader) umma_commit(dq_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================ math warpgroups ============================================
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                         // key row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int s_lo = wg == 0 ? 0 : AB_STEPS_WG0, s_hi = wg == 0 ? AB_STEPS_WG0 : AB_NSTEP;

    const int nWw = p.W / WS;
    const int wr = (bwin % ((p.H / WS) * nWw)) / nWw;
    const int wc = (bwin % ((p.H / WS) * nWw)) % nWw;
    const bool rowflag = p.shift > 0 && (wr == p.H / WS - 1);
    const bool colflag = p.shift > 0 && (wc == nWw - 1);

    int n = 0;
    for (int j = 0; j < NKT; ++j) {
      const int kk = j * AB_KT + r;
      const bool kvalid = kk < NTOK;
      const int kc = kvalid ? kk : NTOK - 1;
      const int ky = kc / WS, kx = kc - ky * WS;
      const bool rk = ky >= SPLIT, ck = kx >= SPLIT;
      bf16* grow = p.gt ? p.gt + ((size_t)bh * NTOK + kc) * p.ntok_pad : nullptr;
      // G^T row of this key in shared memory: 16-byte chunk c of atom a sits at a * 16 KB + (r / 8) KB + (r % 8) * 128 +
      // ((c ^ (r % 8)) * 16)
      uint8_t* srow = sG + (r >> 3) * 1024 + (r & 7) * 128;

      for (int i = 0; i < NQT; ++i, ++n) {
        // table row of query window row (i * RPT + qy') against this key: base + qy' * TS + qx
        const float* tb = sTab + (i * RPT - ky + WS - 1) * TS + (WS - 1 - kx);
        mbar_wait(sdp_full, n & 1, 10);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < AB_NSTEP; ++s) {
          if (s < s_lo || s >= s_hi) continue;
          uint32_t sv[16], dp[16];
          tmem_ld16(tmem_base + lane_off + Cf::COL_ST + 16 * s, sv);
          tmem_ld16(tmem_base + lane_off + Cf::COL_DP + 16 * s, dp);
          tmem_ld_wait();
          uint32_t pw[8], gw[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float pv[2], gv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cq = 16 * s + c + e;                   // query column inside the tile (compile time)
              const int qyl = cq / WS, qx = cq - qyl * WS;
              const float2 ldq = sLD[i * AB_QT + cq];          // (LSE, D) of the query: one broadcast load
              const int qy = i * RPT + qyl;
              const bool masked = (rowflag && ((qy >= SPLIT) != rk)) || (colflag && ((qx >= SPLIT) != ck));
              float x = __uint_as_float(sv[c + e]) + tb[qyl * TS + qx] + (masked ? NEG100 : 0.f) - ldq.x;
              const float pe = ex2_approx(x);
              pv[e] = pe;
              gv[e] = pe * (__uint_as_float(dp[c + e]) - ldq.y);
            }
            pw[c >> 1] = pack_bf16x2(pv[0], pv[1]);
            gw[c >> 1] = pack_bf16x2(gv[0], gv[1]);
          }
          tmem_st8p(tmem_base + lane_off + Cf::COL_ST + 16 * s, pw);
          tmem_st8p(tmem_base + lane_off + Cf::COL_DP + 16 * s, gw);
          // shared-memory copy (A operand of dQ^): queries [16 s, 16 s + 16) = chunks 2 s, 2 s + 1 of the key's row
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = 2 * s + h;
            const uint4 val = make_uint4(gw[4 * h], gw[4 * h + 1], gw[4 * h + 2], gw[4 * h + 3]);
            *reinterpret_cast<uint4*>(srow + (c >> 3) * (AB_KT * 128) + (((c & 7) ^ (r & 7)) << 4)) = val;
            const int q0 = i * AB_QT + 16 * s + 8 * h;
            if (grow != nullptr && kvalid && q0 < p.ntok_pad) *reinterpret_cast<uint4*>(grow + q0) = val;
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pg_full);
      }

      // ---- dV (warpgroup 0) / dK^ (warpgroup 1) of key tile j ----
      mbar_wait(dkv_full, j & 1, 11);
      tc_fence_after();
      {
        uint32_t o[32];
        tmem_ld32(tmem_base + lane_off + (wg == 0 ? Cf::COL_DV : Cf::COL_DK), o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);
        if (kvalid) {
          float* dst = (wg == 0 ? p.dv : p.dk) + ((size_t)bh * NTOK + kk) * 32;
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<uint4*>(dst + q) = make_uint4(o[q], o[q + 1], o[q + 2], o[q + 3]);
        }
      }
    }
    // ---- dQ^ ----
    mbar_wait(dq_full, 0, 12);
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < NQT; ++i) {
      if ((i & 1) != wg) continue;
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_off + Cf::COL_DQ + 32 * i, o);
      tmem_ld_wait();
      const int q = i * AB_QT + r;
      if (r < AB_QT && q < NTOK) {
        float* dst = p.dq + ((size_t)bh * NTOK + q) * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<uint4*>(dst + c) = make_uint4(o[c], o[c + 1], o[c + 2], o[c + 3]);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cf::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// Preparation: one thread per (token, head).  Gathers dO (token-major, gradient of the attention output before proj)
// into the window-major head-major order of q / k / v with the cyclic shift applied, D = rowsum(dO o O), packs
// (LSE, D), and makes bf16 copies of q^ / k^ (the fp16 originals feed the score recomputation, the bf16 copies are the
// MN-major B operands of dK^ += G^T Q^ and dQ^ += G K^ next to the bf16 G).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O, const float* __restrict__ lse,
                     const __half* __restrict__ qh, const __half* __restrict__ kh, bf16* __restrict__ dOw,
                     float2* __restrict__ ld, bf16* __restrict__ qb, bf16* __restrict__ kb, int B, int H, int W, int C,
                     int nH, int ws, int shift) {
  // Threads walk the WINDOW-major order (consecutive threads = consecutive slots of one (window, head)): six of the eight
  // tensors touched here are window-major, so a warp reads / writes 2 KB runs instead of 32 pieces 50 KB apart; the two
  // token-major operands (dO, O) are then 64-byte pieces one token row apart.
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * nH;
  if (idx >= total) return;
  const int N = ws * ws;
  const int slot = (int)(idx % N);
  const long long bwh = idx / N;
  const int head = (int)(bwh % nH);
  const long long bw = bwh / nH;
  const int nWw = W / ws, nW = (H / ws) * nWw;
  const int win = (int)(bw % nW), b = (int)(bw / nW);
  int hh = (win / nWw) * ws + slot / ws + shift, ww = (win % nWw) * ws + slot % ws + shift;   // undo the cyclic shift
  if (hh >= H) hh -= H;
  if (ww >= W) ww -= W;
  const long long row = (long long)b * H * W + (long long)hh * W + ww;                         // token
  const size_t wrow = (size_t)idx;
  const uint4* dp = reinterpret_cast<const uint4*>(dO + (size_t)row * C + head * 32);
  const uint4* op = reinterpret_cast<const uint4*>(O + (size_t)row * C + head * 32);
  uint4* dst = reinterpret_cast<uint4*>(dOw + wrow * 32);
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 a = __ldg(dp + i), o = __ldg(op + i);
    d += bf16_lo(a.x) * bf16_lo(o.x) + bf16_hi(a.x) * bf16_hi(o.x) + bf16_lo(a.y) * bf16_lo(o.y) +
         bf16_hi(a.y) * bf16_hi(o.y) + bf16_lo(a.z) * bf16_lo(o.z) + bf16_hi(a.z) * bf16_hi(o.z) +
         bf16_lo(a.w) * bf16_lo(o.w) + bf16_hi(a.w) * bf16_hi(o.w);
    dst[i] = a;
  }
  ld[wrow] = make_float2(lse[wrow], d);
  // fp16 -> bf16 copies of this (token, head)'s 32 q^ and 32 k^ values: four 16-byte loads and stores per tensor (as 16
  // half2 accesses per thread every warp instruction touched 32 sectors for 4 bytes each)
  const uint4* qs = reinterpret_cast<const uint4*>(qh + wrow * 32);
  const uint4* ks = reinterpret_cast<const uint4*>(kh + wrow * 32);
  uint4* qd = reinterpret_cast<uint4*>(qb + wrow * 32);
  uint4* kd = reinterpret_cast<uint4*>(kb + wrow * 32);
  auto cvt = [](uint32_t h2) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
    return pack_bf16x2(f.x, f.y);
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 a = __ldg(qs + i), c = __ldg(ks + i);
    qd[i] = make_uint4(cvt(a.x), cvt(a.y), cvt(a.z), cvt(a.w));
    kd[i] = make_uint4(cvt(c.x), cvt(c.y), cvt(c.z), cvt(c.w));
  }
}

template <int WS>
static int launch_attn_bwd(const void* qh, const void* qb, const void* kh, const void* kb, const void* v, const void* dOw,
                           int n_bh, const AttnBwdParams& p, cudaStream_t stream) {
  using Cf = AbCfg<WS>;
  CUtensorMap tmQh, tmQb, tmKh, tmKb, tmV, tmdO;
  uint64_t dims[3] = {32, (uint64_t)Cf::NTOK, (uint64_t)n_bh};
  uint64_t str[2] = {64, (uint64_t)Cf::NTOK * 64};
  uint32_t bq[3] = {32, AB_QT, 1}, bk[3] = {32, AB_KT, 1};
  int rc;
  if ((rc = make_tmap_16b(&tmQh, qh, 3, dims, str, bq, 64))) return rc;
  if ((rc = make_tmap_16b(&tmQb, qb, 3, dims, str, bq, 64))) return rc;
  if ((rc = make_tmap_16b(&tmdO, dOw, 3, dims, str, bq, 64))) return rc;
  if ((rc = make_tmap_16b(&tmKh, kh, 3, dims, str, bk, 64))) return rc;
  if ((rc = make_tmap_16b(&tmKb, kb, 3, dims, str, bk, 64))) return rc;
  if ((rc = make_tmap_16b(&tmV, v, 3, dims, str, bk, 64))) return rc;
  auto kern = attn_bwd_kernel<WS>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM));
  kern<<<n_bh, AB_THREADS, Cf::SMEM, stream>>>(tmQh, tmQb, tmKh, tmKb, tmV, tmdO, p);
  MV_LAUNCH_OK();
  return 0;
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_swin_attention_bwd_prep(const void* dO, const void* O, const float* lse, const void* qh,
                                             const void* kh, void* dOw, void* ld, void* qb, void* kb, int B, int H,
                                             int W, int C, int nH, int ws, int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C == nH * 32, "attention_bwd_prep: head_dim must be 32");
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "attention_bwd_prep: bad window geometry");
  const long long total = (long long)B * H * W * nH;
  if (total <= 0) return 0;
  attn_bwd_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dO), reinterpret_cast<const bf16*>(O), lse, reinterpret_cast<const __half*>(qh),
      reinterpret_cast<const __half*>(kh), reinterpret_cast<bf16*>(dOw), reinterpret_cast<float2*>(ld),
      reinterpret_cast<bf16*>(qb), reinterpret_cast<bf16*>(kb), B, H, W, C, nH, ws, shift);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_swin_attention_bwd(const void* qh, const void* qb, const void* kh, const void* kb, const void* v,
                                        const void* dOw, const void* ld, const float* bias_rev, float* dq, float* dk,
                                        float* dv, void* gt, int ntok_pad, int B, int H, int W, int nH, int ws,
                                        int shift, cudaStream_t stream) {
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0, "attention_bwd: window must tile the token grid");
  MV_CHECK_ARG(shift == 0 || shift == ws / 2, "attention_bwd: shift must be 0 or ws/2");
  MV_CHECK_ARG(gt == nullptr || (ntok_pad >= ws * ws && ntok_pad % 8 == 0), "attention_bwd: ntok_pad must cover ws^2 in multiples of 8");
  AttnBwdParams p{};
  p.nH = nH; p.H = H; p.W = W; p.shift = shift;
  p.bias_rev = bias_rev;
  p.ld = reinterpret_cast<const float2*>(ld);
  p.dq = dq; p.dk = dk; p.dv = dv;
  p.gt = reinterpret_cast<bf16*>(gt);
  p.ntok_pad = ntok_pad;
  const int n_bh = B * (H / ws) * (W / ws) * nH;
  switch (ws) {
    case 28: return launch_attn_bwd<28>(qh, qb, kh, kb, v, dOw, n_bh, p, stream);
    case 14: return launch_attn_bwd<14>(qh, qb, kh, kb, v, dOw, n_bh, p, stream);
    case 7: return launch_attn_bwd<7>(qh, qb, kh, kb, v, dOw, n_bh, p, stream);
    default: return mv::fail(-1, "attention_bwd: window %d not instantiated (7, 14, 28)", ws);
  }
}

// =========================================================================================================
// Backward of the RoBERTa self-attention of the text encoder (HF RobertaSelfAttention under autograd, as fine-tuned
// through mvuld/models/unixcoder.py:33-54): head dim 64, sequences of up to 512 tokens, keys past kv_len masked.
// Same structure as attn_bwd_kernel (transposed score tiles, P^T / G^T as TMEM A operands, G^T once through shared
// memory for dQ), with 128-query tiles and without a bias table.  dQ of all query tiles does not fit TMEM next to the
// 64-wide accumulators, so each unit's dQ contribution is added to the fp32 dq rows in global memory by the CTA that
// owns the (sequence, head) -- sequential over the key tiles, hence still deterministic.
// q is stored pre-scaled by log2(e) / sqrt(hd) (mvuld_heads_qkv): outputs dq = G k, dk = G^T q_stored, dv = P^T dO with
// G = dL / d(natural logits).
// =========================================================================================================
namespace mv {

constexpr int SB_T = 128;            // keys per tile (TMEM lanes) and queries per tile (TMEM columns)
constexpr int SB_HD = 64;
constexpr int SB_TILE = SB_T * SB_HD * 2;        // 16 KB operand tile, 128-byte rows
struct SeqBwdParams {
  int nH, L;
  const int* kv_len;       // [B]
  const float2* ld;        // [B * nH, L] (LSE log2 units, D)
  float* dq;               // [B * nH, L, 64] fp32 (zero-filled by the caller: tiles past kv_len are skipped)
  float* dk;
  float* dv;
};
constexpr int SB_SMEM = 2 * SB_T * 128 /*G: 2 atoms of 64 queries*/ + 2 * SB_TILE /*K, V*/ + 4 * SB_TILE /*Q, dO x2*/ +
                        4096 /*LD*/ + 256 + 1024;

__global__ void __launch_bounds__(AB_THREADS, 1)
seq_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, SeqBwdParams p) {
  constexpr int COL_ST = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
  constexpr int NSTEP = SB_T / 16;                   // 8 query steps per tile, 4 per math warpgroup
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sG = smem;
  uint8_t* sK = sG + 2 * SB_T * 128;
  uint8_t* sV = sK + SB_TILE;
  uint8_t* sQ = sV + SB_TILE;                        // [2]
  uint8_t* sdO = sQ + 2 * SB_TILE;                   // [2]
  float2* sLD = reinterpret_cast<float2*>(sdO + 2 * SB_TILE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLD) + 4096);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;       // [2]
  uint64_t* q_empty = bars + 4;      // [2]
  uint64_t* sdp_full = bars + 6;
  uint64_t* pg_full = bars + 7;
  uint64_t* unit_done = bars + 8;    // the unit's dV / dK / dQ products retired (dQ of the unit is readable)
  uint64_t* dq_free = bars + 9;      // ... and its dQ tile has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int b = bh / p.nH;
  int len = __ldg(p.kv_len + b);
  len = len < 1 ? 1 : (len > p.L ? p.L : len);
  const int nt = (len + SB_T - 1) / SB_T;            // key tiles == query tiles that hold valid tokens

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmdO);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pg_full, 8);
    mbar_init(unit_done, 1);
    mbar_init(dq_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  {
    const float2* ldg = p.ld + (size_t)bh * p.L;
    for (int i = threadIdx.x; i < 512; i += AB_THREADS) sLD[i] = i < p.L ? __ldg(ldg + i) : make_float2(1.0e30f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int j = 0; j < nt; ++j) {
        if (j > 0) mbar_wait(kv_empty, (j - 1) & 1, 1);
        mbar_arrive_expect_tx(kv_full, 2 * SB_TILE);
        tma_load_3d(sK, &tmK, kv_full, 0, j * SB_T, bh);
        tma_load_3d(sV, &tmV, kv_full, 0, j * SB_T, bh);
        for (int i = 0; i < nt; ++i, ++n) {
          const int st = n & 1;
          if (n >= 2) mbar_wait(&q_empty[st], ((n >> 1) - 1) & 1, 2);
          mbar_arrive_expect_tx(&q_full[st], 2 * SB_TILE);
          tma_load_3d(sQ + st * SB_TILE, &tmQ, &q_full[st], 0, i * SB_T, bh);
          tma_load_3d(sdO + st * SB_TILE, &tmdO, &q_full[st], 0, i * SB_T, bh);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_st = make_idesc_bf16(SB_T, SB_T, 0, 0);
    constexpr uint32_t idesc_dv = make_idesc_bf16(SB_T, SB_HD, 0, 1);        // A from TMEM, B MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(SB_T, SB_HD, 1, 1);        // A MN-major (shared memory), B MN-major
    // 128-byte-swizzle tiles with 128-byte rows: K-major = MN-major atom layout, 8-row groups 1 KB apart
    const uint32_t hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    const uint32_t g_lo = (uint32_t)make_smem_desc(smem_u32(sG), SB_T * 128, 1024, 2);
    const uint32_t k_lo = (uint32_t)make_smem_desc(smem_u32(sK), 16, 1024, 2);
    const uint32_t v_lo = (uint32_t)make_smem_desc(smem_u32(sV), 16, 1024, 2);
    const uint32_t q_lo = (uint32_t)make_smem_desc(smem_u32(sQ), 16, 1024, 2);
    const uint32_t do_lo = (uint32_t)make_smem_desc(smem_u32(sdO), 16, 1024, 2);
    const bool leader = elect_one();
    auto d = [&](uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    int n = 0;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(kv_full, j & 1, 3);
      for (int i = 0; i < nt; ++i, ++n) {
        const int st = n & 1;
        mbar_wait(&q_full[st], (n >> 1) & 1, 4);
        tc_fence_after();
        const uint32_t qoff = (uint32_t)st * (SB_TILE >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < SB_HD / 16; ++k)
            umma_ss(tmem_base + COL_ST, d(k_lo + k * 2), d(q_lo + qoff + k * 2), idesc_st, k != 0);
#pragma unroll
          for (int k = 0; k < SB_HD / 16; ++k)
            umma_ss(tmem_base + COL_DP, d(v_lo + k * 2), d(do_lo + qoff + k * 2), idesc_st, k != 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(pg_full, n & 1, 6);
        if (n > 0) mbar_wait(dq_free, (n - 1) & 1, 7);        // the previous unit's dQ tile has left TMEM
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int s = 0; s < NSTEP; ++s)
            umma_ts(tmem_base + COL_DV, tmem_base + COL_ST + 16 * s, d(do_lo + qoff + s * (16 * 128 >> 4)), idesc_dv,
                    (i != 0) || (s != 0));
#pragma unroll
          for (int s = 0; s < NSTEP; ++s)
            umma_ts(tmem_base + COL_DK, tmem_base + COL_DP + 16 * s, d(q_lo + qoff + s * (16 * 128 >> 4)), idesc_dv,
                    (i != 0) || (s != 0));
#pragma unroll
          for (int s = 0; s < SB_T / 16; ++s)
            umma_ss(tmem_base + COL_DQ, d(g_lo + s * (16 * 128 >> 4)), d(k_lo + s * (16 * 128 >> 4)), idesc_dq, s != 0);
          umma_commit(&q_empty[st]);
          umma_commit(unit_done);
          if (i == nt - 1) umma_commit(kv_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int s_lo = wg * (NSTEP / 2), s_hi = s_lo + NSTEP / 2;
    int n = 0;
    for (int j = 0; j < nt; ++j) {
      const int kk = j * SB_T + r;
      const bool kvalid = kk < len;
      uint8_t* srow = sG + (r >> 3) * 1024 + (r & 7) * 128;
      for (int i = 0; i < nt; ++i, ++n) {
        mbar_wait(sdp_full, n & 1, 10);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < NSTEP; ++s) {
          if (s < s_lo || s >= s_hi) continue;
          uint32_t sv[16], dp[16];
          tmem_ld16(tmem_base + lane_off + COL_ST + 16 * s, sv);
          tmem_ld16(tmem_base + lane_off + COL_DP + 16 * s, dp);
          tmem_ld_wait();
          uint32_t pw[8], gw[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float pv[2], gv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float2 ldq = sLD[i * SB_T + 16 * s + c + e];
              const float pe = kvalid ? ex2_approx(__uint_as_float(sv[c + e]) - ldq.x) : 0.f;
              pv[e] = pe;
              gv[e] = pe * (__uint_as_float(dp[c + e]) - ldq.y);
            }
            pw[c >> 1] = pack_bf16x2(pv[0], pv[1]);
            gw[c >> 1] = pack_bf16x2(gv[0], gv[1]);
          }
          tmem_st8p(tmem_base + lane_off + COL_ST + 16 * s, pw);
          tmem_st8p(tmem_base + lane_off + COL_DP + 16 * s, gw);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = 2 * s + h;
            *reinterpret_cast<uint4*>(srow + (c >> 3) * (SB_T * 128) + (((c & 7) ^ (r & 7)) << 4)) =
                make_uint4(gw[4 * h], gw[4 * h + 1], gw[4 * h + 2], gw[4 * h + 3]);
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pg_full);
        // ---- this unit's dQ tile (rows = queries): add into the fp32 rows of query tile i (warpgroup = column half) ----
        mbar_wait(unit_done, n & 1, 11);
        tc_fence_after();
        {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + COL_DQ + 32 * wg, o);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_free);
          const int q = i * SB_T + r;
          if (q < len) {
            float4* dst = reinterpret_cast<float4*>(p.dq + ((size_t)bh * p.L + q) * SB_HD + 32 * wg);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 a = make_float4(__uint_as_float(o[4 * c]), __uint_as_float(o[4 * c + 1]),
                                     __uint_as_float(o[4 * c + 2]), __uint_as_float(o[4 * c + 3]));
              if (j > 0) {
                const float4 old = dst[c];
                a.x += old.x; a.y += old.y; a.z += old.z; a.w += old.w;
              }
              dst[c] = a;
            }
          }
        }
      }
      // ---- dV (warpgroup 0) / dK (warpgroup 1) of key tile j: the last unit_done wait above covered these products ----
      {
        float* base = (wg == 0 ? p.dv : p.dk) + ((size_t)bh * p.L + kk) * SB_HD;
#pragma unroll
        for (int c0 = 0; c0 < SB_HD; c0 += 32) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_off + (wg == 0 ? COL_DV : COL_DK) + c0, o);
          tmem_ld_wait();
          if (kvalid) {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              *reinterpret_cast<uint4*>(base + c0 + q) = make_uint4(o[q], o[q + 1], o[q + 2], o[q + 3]);
          }
        }
        tc_fence_before();
      }
      // the next key tile's first dV / dK products must not start before every warp has read these accumulators: the
      // issuer's next pg_full wait needs all eight warps, and each arrives only after this point
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// one thread per (token, head): dO token-major [B*L, nH*64] -> head-major [B, nH, L, 64]; ld = (lse, rowsum(dO o O))
__global__ void __launch_bounds__(256)
seq_attn_bwd_prep_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O, const float* __restrict__ lse,
                         bf16* __restrict__ dOh, float2* __restrict__ ld, int B, int L, int nH) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * L * nH) return;
  const int head = (int)(idx % nH);
  const long long row = idx / nH;
  const int b = (int)(row / L), t = (int)(row - (long long)b * L);
  const size_t hrow = ((size_t)b * nH + head) * L + t;
  const uint4* dp = reinterpret_cast<const uint4*>(dO + (size_t)row * nH * SB_HD + head * SB_HD);
  const uint4* op = reinterpret_cast<const uint4*>(O + (size_t)row * nH * SB_HD + head * SB_HD);
  uint4* dst = reinterpret_cast<uint4*>(dOh + hrow * SB_HD);
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a = __ldg(dp + i), o = __ldg(op + i);
    d += bf16_lo(a.x) * bf16_lo(o.x) + bf16_hi(a.x) * bf16_hi(o.x) + bf16_lo(a.y) * bf16_lo(o.y) +
         bf16_hi(a.y) * bf16_hi(o.y) + bf16_lo(a.z) * bf16_lo(o.z) + bf16_hi(a.z) * bf16_hi(o.z) +
         bf16_lo(a.w) * bf16_lo(o.w) + bf16_hi(a.w) * bf16_hi(o.w);
    dst[i] = a;
  }
  ld[hrow] = make_float2(lse[hrow], d);
}

// (dq, dk, dv) fp32 head-major -> d(x Wqkv^T + b) bf16 token-major [B*L, 3*Hd] (q | k | v): dq / sqrt(hd), ln2 dk, dv
__global__ void __launch_bounds__(256)
seq_qkv_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                   bf16* __restrict__ dqkv, int B, int L, int nH, float q_mul, float k_mul) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * L * nH) return;
  const int head = (int)(idx % nH);
  const long long row = idx / nH;
  const int b = (int)(row / L), t = (int)(row - (long long)b * L);
  const size_t hrow = ((size_t)b * nH + head) * L + t;
  const int Hd = nH * SB_HD;
  const float* src[3] = {dq + hrow * SB_HD, dk + hrow * SB_HD, dv + hrow * SB_HD};
  const float mul[3] = {q_mul, k_mul, 1.0f};
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    uint4* o = reinterpret_cast<uint4*>(dqkv + (size_t)row * 3 * Hd + w * Hd + head * SB_HD);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src[w]) + 2 * i);
      const float4 c = __ldg(reinterpret_cast<const float4*>(src[w]) + 2 * i + 1);
      o[i] = make_uint4(pack_bf16x2(a.x * mul[w], a.y * mul[w]), pack_bf16x2(a.z * mul[w], a.w * mul[w]),
                        pack_bf16x2(c.x * mul[w], c.y * mul[w]), pack_bf16x2(c.z * mul[w], c.w * mul[w]));
    }
  }
}

}  // namespace mv

extern "C" int mvuld_seq_attention_bwd_prep(const void* dO, const void* O, const float* lse, void* dOh, void* ld, int B,
                                            int L, int nH, cudaStream_t stream) {
  const long long total = (long long)B * L * nH;
  if (total <= 0) return 0;
  seq_attn_bwd_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dO), reinterpret_cast<const bf16*>(O), lse, reinterpret_cast<bf16*>(dOh),
      reinterpret_cast<float2*>(ld), B, L, nH);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_seq_attention_bwd(const void* q, const void* k, const void* v, const void* dOh, const void* ld,
                                       const int* kv_len, float* dq, float* dk, float* dv, int B, int L, int nH, int hd,
                                       cudaStream_t stream) {
  MV_CHECK_ARG(hd == SB_HD, "seq attention backward: head_dim 64 only");
  MV_CHECK_ARG(L <= 512 && L % 8 == 0, "seq attention backward: L must be <= 512 and a multiple of 8");
  SeqBwdParams p{};
  p.nH = nH; p.L = L; p.kv_len = kv_len;
  p.ld = reinterpret_cast<const float2*>(ld);
  p.dq = dq; p.dk = dk; p.dv = dv;
  const int n_bh = B * nH;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  uint64_t dims[3] = {SB_HD, (uint64_t)L, (uint64_t)n_bh};
  uint64_t str[2] = {SB_HD * 2, (uint64_t)L * SB_HD * 2};
  uint32_t box[3] = {SB_HD, SB_T, 1};
  int rc;
  if ((rc = make_tmap_16b(&tmQ, q, 3, dims, str, box, 128))) return rc;
  if ((rc = make_tmap_16b(&tmK, k, 3, dims, str, box, 128))) return rc;
  if ((rc = make_tmap_16b(&tmV, v, 3, dims, str, box, 128))) return rc;
  if ((rc = make_tmap_16b(&tmdO, dOh, 3, dims, str, box, 128))) return rc;
  MV_CUDA_OK(cudaFuncSetAttribute(seq_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM));
  seq_attn_bwd_kernel<<<n_bh, AB_THREADS, SB_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_seq_qkv_bwd(const float* dq, const float* dk, const float* dv, void* dqkv, int B, int L, int nH,
                                 int hd, cudaStream_t stream) {
  MV_CHECK_ARG(hd == SB_HD, "seq_qkv_bwd: head_dim 64 only");
  const long long total = (long long)B * L * nH;
  if (total <= 0) return 0;
  // q is stored as (x Wq + b) log2e / sqrt(hd): d(x Wq + b) = ln2 (G k) log2e / sqrt(hd) = dq / sqrt(hd); d k = ln2 G^T q_stored
  seq_qkv_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dq, dk, dv, reinterpret_cast<bf16*>(dqkv), B,
                                                                         L, nH, 1.0f / sqrtf((float)hd),
                                                                         0.6931471805599453f);
  MV_LAUNCH_OK();
  return 0;
}
