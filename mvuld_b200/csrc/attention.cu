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
// query tiles.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (single thread), warp 2 TMEM allocator,
// warps 4-7 / 8-11 two softmax warpgroups (thread == query row, no cross-thread reductions) that ping-pong so one
// group's exp/bias work overlaps the other's QK^T / PV tensor-core work.  S and O accumulators live in TMEM;
// P goes back through 128B-swizzled shared memory as the A operand of the PV MMA; V is consumed MN-major so no
// transpose is ever written.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int ATT_THREADS = 384;
constexpr int ATT_BM = 128;          // query rows per tile (= TMEM lanes)
constexpr int ATT_MAX_KT = 8;        // max kv tiles resident
constexpr int ATT_REGS_CTRL = 88;      // setmaxnreg for warps 0-3 (128 threads)
constexpr int ATT_REGS_SOFTMAX = 208;  // 8 softmax warps: 128*88 + 256*208 = 64512 = 384 threads * 168 regs at launch

enum { MODE_SWIN = 0, MODE_SEQ = 1 };

struct AttnParams {
  int Nq, Nkv;            // tokens per window / sequence
  int nH;                 // heads
  // MODE_SWIN
  const float* bias_rev;  // [nH, (2ws-1)^2] fp32, = 16*sigmoid(.)*log2e, w-axis reversed (see cpb kernel)
  const float* bias_max;  // [nH] max of the head's table (softmax reference bound)
  int H, W, shift;        // token grid and cyclic shift of this block
  int C;                  // channels (= nH * HD)
  // MODE_SEQ
  const int* kv_len;      // [B] valid keys per sequence
  void* out;              // bf16 [tokens, C]
};

template <int HD>
struct AttnLayout {
  static constexpr int ROW_BYTES = HD * 2;                    // 64 (SW64) or 128 (SW128)
  static constexpr int SWZ = ROW_BYTES;                       // swizzle span == row
  static constexpr int LAYOUT = (HD == 32) ? 4 : 2;           // UMMA layout_type
  static constexpr int SBO = 8 * ROW_BYTES;                   // 8-row core-matrix group
  static constexpr int Q_BYTES = ATT_BM * ROW_BYTES;
  static constexpr int P_BYTES = ATT_BM * 128 * 2;            // 2 chunks of [128 rows x 128 B]
};

__host__ __device__ constexpr int att_smem_bytes(int HD, int KT, int nkt, int table_floats) {
  return 2 * nkt * KT * HD * 2      // K, V
         + 2 * ATT_BM * HD * 2      // Q x2
         + 2 * ATT_BM * 128 * 2     // P x2
         + ((table_floats * 4 + 1023) / 1024) * 1024 + 512 /*barriers*/ + 1024 /*align*/;
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
                const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  using L = AttnLayout<HD>;
  constexpr int SIDE = 2 * WS - 1;
  constexpr int TSTRIDE = att_tab_stride(WS);
  constexpr int TBL = (MODE == MODE_SWIN) ? SIDE * TSTRIDE : 0;          // shared-memory floats (padded rows)
  constexpr int ROWS_PER_TILE = (MODE == MODE_SWIN) ? KT / WS : 1;
  constexpr int SPLIT = WS - WS / 2;                 // first column / row of the "shifted-in" band
  constexpr int NSEG = (MODE == MODE_SWIN) ? ROWS_PER_TILE * 2 : 4;     // independent max chains
  static_assert(MODE != MODE_SWIN || KT % WS == 0, "kv tile must hold whole window rows");
  static_assert(KT % 16 == 0 && KT <= 128, "kv tile");

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

  uint8_t* sK = smem;
  uint8_t* sV = sK + nkt_all * KT * L::ROW_BYTES;
  uint8_t* sQ = sV + nkt_all * KT * L::ROW_BYTES;
  uint8_t* sP = sQ + 2 * L::Q_BYTES;
  float* sTab = reinterpret_cast<float*>(sP + 2 * L::P_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sTab) + ((TBL * 4 + 1023) / 1024) * 1024);
  uint64_t* k_full = bars;                    // [ATT_MAX_KT]
  uint64_t* v_full = k_full + ATT_MAX_KT;     // [ATT_MAX_KT]
  uint64_t* q_full = v_full + ATT_MAX_KT;     // [2]
  uint64_t* q_empty = q_full + 2;
  uint64_t* s_full = q_empty + 2;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* pv_done = p_full + 2;
  uint64_t* o_free = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    for (int i = 0; i < ATT_MAX_KT; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1);
      mbar_init(&q_empty[g], 1);
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
      mbar_init(&o_free[g], 128);
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
  const uint32_t tmem_S = tmem_base;           // S[g] at columns g*128
  const uint32_t tmem_O = tmem_base + 256;     // O[g] at columns 256 + g*64

  if (warp < 4) {
    reg_dec<ATT_REGS_CTRL>();                  // producer / MMA / allocator warps need few registers
    if (warp == 0) {
      // =========================================== TMA producer ===========================================
      if (lane == 0) {
        auto load_q = [&](int g, int t) {
          mbar_arrive_expect_tx(&q_full[g], L::Q_BYTES);
          tma_load_3d(sQ + g * L::Q_BYTES, &tmQ, &q_full[g], 0, t * ATT_BM, bh);
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
    } else if (warp == 1) {
      // ============================================ MMA issuer ============================================
      if (lane == 0) {
        constexpr uint32_t idesc_s =
            make_idesc_bf16(ATT_BM, KT, 0, 0) & ~((QK_FP16 ? 1u : 0u) * ((1u << 7) | (1u << 10)));
        constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, HD, 0, 1);
        const uint32_t q0 = smem_u32(sQ), k0 = smem_u32(sK), v0 = smem_u32(sV), p0 = smem_u32(sP);
        int s_count[2] = {0, 0};               // S tiles issued per group: S[g] is reused once its reader freed it
        auto issue_s = [&](int g, int j, bool first_pass) {
          if (s_count[g] > 0) mbar_wait(&s_free[g], (s_count[g] - 1) & 1, 26);
          if (first_pass) mbar_wait(&k_full[j], 0, 21);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint64_t ad = make_smem_desc(q0 + g * L::Q_BYTES + k * 32, 16, L::SBO, L::LAYOUT);
            const uint64_t bd = make_smem_desc(k0 + j * KT * L::ROW_BYTES + k * 32, 16, L::SBO, L::LAYOUT);
            umma_ss(tmem_S + g * 128, ad, bd, idesc_s, k != 0);
          }
          umma_commit(&s_full[g]);
          ++s_count[g];
        };
        auto issue_pv = [&](int g, int j) {
#pragma unroll
          for (int s = 0; s < KT / 16; ++s) {
            const uint64_t ad = make_smem_desc(p0 + g * L::P_BYTES + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(v0 + (j * KT + s * 16) * L::ROW_BYTES, 16, L::SBO, L::LAYOUT);
            umma_ss(tmem_O + g * 64, ad, bd, idesc_pv, (j | s) != 0);
          }
          umma_commit(&pv_done[g]);
        };
        uint32_t ph_q[2] = {0, 0}, ph_p[2] = {0, 0}, ph_o[2] = {0, 0};
        const int n_it = (nq + 1) >> 1;
        for (int it = 0; it < n_it; ++it) {
          const bool valid[2] = {true, 2 * it + 1 < nq};
          for (int g = 0; g < 2; ++g) {
            if (!valid[g]) continue;
            mbar_wait(&q_full[g], ph_q[g], 20);
            ph_q[g] ^= 1;
            issue_s(g, 0, it == 0);
            if (nkt == 1) umma_commit(&q_empty[g]);
          }
          for (int j = 0; j < nkt; ++j) {
            // S(j+1) goes out as soon as the softmax group has pulled S(j) into registers ...
            for (int g = 0; g < 2; ++g) {
              if (!valid[g] || j + 1 >= nkt) continue;
              issue_s(g, j + 1, it == 0);
              if (j + 2 == nkt) umma_commit(&q_empty[g]);   // last read of this Q tile is in flight
            }
            // ... and P(j) V(j) once the group has written P(j)
            for (int g = 0; g < 2; ++g) {
              if (!valid[g]) continue;
              mbar_wait(&p_full[g], ph_p[g], 22);
              ph_p[g] ^= 1;
              if (j == 0 && it > 0) {
                mbar_wait(&o_free[g], ph_o[g], 23);
                ph_o[g] ^= 1;
              }
              if (it == 0) mbar_wait(&v_full[j], 0, 24);
              tc_fence_after();
              issue_pv(g, j);
            }
          }
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
    const uint32_t tS = tmem_S + g * 128 + lane_off;
    const uint32_t tO = tmem_O + g * 64 + lane_off;
    uint8_t* myP = sP + g * L::P_BYTES + r * 128;
    uint32_t ph_s = 0, ph_pv = 0;

    // window coordinates (MODE_SWIN)
    const int nWw = (MODE == MODE_SWIN) ? p.W / WS : 1;
    const int wr = (MODE == MODE_SWIN) ? (bwin % ((p.H / WS) * nWw)) / nWw : 0;
    const int wc = (MODE == MODE_SWIN) ? (bwin % ((p.H / WS) * nWw)) % nWw : 0;
    const bool rowflag = (MODE == MODE_SWIN) && p.shift > 0 && (wr == p.H / WS - 1);
    const bool colflag = (MODE == MODE_SWIN) && p.shift > 0 && (wc == nWw - 1);
    const float NEG100 = -100.0f * 1.4426950408889634f;
    const float bmax = (MODE == MODE_SWIN) ? __ldg(p.bias_max + head) : 0.f;

    for (int t = g; t < nq; t += 2) {
      const int i = t * ATT_BM + r;                          // slot inside the window / position in the sequence
      int hi = 0, wi = 0;
      if (MODE == MODE_SWIN) {
        hi = i / WS;
        wi = i - hi * WS;
        if (hi > WS - 1) hi = WS - 1;                        // rows past the window only exist as padding
      }
      const bool ri = hi >= SPLIT, ci = wi >= SPLIT;
      float m_run = -INFINITY;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};                    // split row sum: four independent add chains

      for (int j = 0; j < nkt; ++j) {
        const int ncols = min(KT, kv_valid - j * KT);        // valid kv columns in this tile

        // per-segment additive constants (shift mask) and bias-table row bases
        float cseg[NSEG];
        int tb[ROWS_PER_TILE];
#pragma unroll
        for (int rr = 0; rr < ROWS_PER_TILE; ++rr) {
          if (MODE == MODE_SWIN) {
            int hj = j * ROWS_PER_TILE + rr;
            if (hj > WS - 1) hj = WS - 1;
            const bool rdiff = rowflag && ((hj >= SPLIT) != ri);
            cseg[2 * rr] = (rdiff || (colflag && ci)) ? NEG100 : 0.f;           // wj <  SPLIT
            cseg[2 * rr + 1] = (rdiff || (colflag && !ci)) ? NEG100 : 0.f;      // wj >= SPLIT
            tb[rr] = (hi - hj + WS - 1) * TSTRIDE + (WS - 1 - wi);
          } else {
            cseg[0] = 0.f;
            tb[0] = 0;
          }
        }

        // ---- pull the whole S row into registers, then hand the TMEM buffer back to the MMA warp ----
        mbar_wait(&s_full[g], ph_s, 30);
        ph_s ^= 1;
        tc_fence_after();
        uint32_t sv[KT];
#pragma unroll
        for (int c0 = 0; c0 < KT; c0 += 32) {
          if (c0 + 32 <= KT) tmem_ld32p(tS + c0, sv + c0);
          else tmem_ld16p(tS + c0, sv + c0);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[g]);

        auto tile = [&](auto partial_c) {
          constexpr bool PARTIAL = decltype(partial_c)::value;
          // sweep 1: s += bias (all shared-memory loads of the tile are independent and issued back to back: no
          // store sits between them, so the scheduler can keep dozens of LDS in flight)
          if (MODE == MODE_SWIN) {
#pragma unroll
            for (int c = 0; c < KT; ++c) {
              const int rr = c / WS, wj = c % WS;
              float s = __uint_as_float(sv[c]) + sTab[tb[rr] + wj];
              if (PARTIAL) s = (c < ncols) ? s : -INFINITY;
              sv[c] = __float_as_uint(s);
            }
          } else if (PARTIAL) {
#pragma unroll
            for (int c = 0; c < KT; ++c) sv[c] = (c < ncols) ? sv[c] : 0xff800000u;   // -inf
          }
          // sweep 2: exact row max (segment constants added once per segment)
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

          // running max with lazy rescale (only when the reference point moves by more than 2^8)
          const float m_new = fmaxf(m_run, m_tile);
          const bool need = (j > 0) && (m_new > m_run + 8.0f);
          if (j == 0) m_run = m_new;
          if (j > 0) {
            mbar_wait(&pv_done[g], ph_pv, 31);               // PV(j-1) retired: P buffer free, O up to date
            ph_pv ^= 1;
            tc_fence_after();
            if (__any_sync(0xffffffffu, need)) {
              const float f = need ? ex2_approx(m_run - m_new) : 1.0f;
              if (need) m_run = m_new;
#pragma unroll
              for (int q = 0; q < 4; ++q) l4[q] *= f;
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

          // sweep 3: p = 2^(s + cseg - m_run); row sums; P -> 128B-swizzled shared memory (A operand of PV).
          // Masked / out-of-range columns carry -inf and come out as exactly 0.
          float csm[NSEG];
#pragma unroll
          for (int s = 0; s < NSEG; ++s) csm[s] = ((MODE == MODE_SWIN) ? cseg[s] : 0.f) - m_run;
          // the exponentials are issued as one long run of independent MUFU ops (the XU pipe takes one warp
          // instruction every 8 cycles; the other warpgroup's FADD / LDS / STS work fills the issue slots in between)
#pragma unroll
          for (int c = 0; c < KT; ++c) {
            const int seg = (MODE == MODE_SWIN) ? ((c / WS) * 2 + ((c % WS) >= SPLIT ? 1 : 0)) : 0;
            sv[c] = __float_as_uint(__uint_as_float(sv[c]) + csm[seg]);
          }
#pragma unroll
          for (int c = 0; c < KT; ++c) sv[c] = __float_as_uint(ex2_approx(__uint_as_float(sv[c])));
#pragma unroll
          for (int c0 = 0; c0 < KT; c0 += 8) {
            float pv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float e = __uint_as_float(sv[c0 + q]);
              l4[q & 3] += e;
              pv[q] = e;
            }
            uint4 w;
            w.x = pack_bf16x2(pv[0], pv[1]);
            w.y = pack_bf16x2(pv[2], pv[3]);
            w.z = pack_bf16x2(pv[4], pv[5]);
            w.w = pack_bf16x2(pv[6], pv[7]);
            // 8 columns = one 16-byte unit; unit u of row r sits at ((u ^ (r & 7)) * 16) inside its 128-byte row
            const int chunk = c0 >> 6, unit = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(myP + chunk * 16384 + ((unit ^ (r & 7)) << 4)) = w;
          }
        };
        if (ncols < KT) tile(std::true_type{});
        else tile(std::false_type{});

        fence_proxy_async_smem();      // P (generic proxy) -> visible to the tensor core (async proxy)
        tc_fence_before();             // orders our tcgen05.ld/st of S and O before the MMA warp's next issue
        mbar_arrive(&p_full[g]);
      }

      // ---- epilogue: O / l -> bf16, token-major store ----
      mbar_wait(&pv_done[g], ph_pv, 32);
      ph_pv ^= 1;
      tc_fence_after();
      const float inv = 1.0f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
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

template <int MODE, int HD, int WS, int KT, bool QK_FP16>
static int launch_attn(const void* q, const void* k, const void* v, int n_bh, const AttnParams& p,
                       cudaStream_t stream) {
  using L = AttnLayout<HD>;
  const int nkt = (p.Nkv + KT - 1) / KT;
  MV_CHECK_ARG(nkt <= ATT_MAX_KT, "attention: %d kv tiles exceed the resident maximum %d", nkt, ATT_MAX_KT);
  constexpr int TBL = (MODE == MODE_SWIN) ? (2 * WS - 1) * att_tab_stride(WS) : 0;
  const int smem = att_smem_bytes(HD, KT, nkt, TBL);
  MV_CHECK_ARG(smem <= 232448, "attention: %d B shared memory needed, 232448 available", smem);
  CUtensorMap tmQ, tmK, tmV;
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
  auto kern = attn_fwd_kernel<MODE, HD, WS, KT, QK_FP16>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<n_bh, ATT_THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
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

extern "C" int mvuld_swin_window_attention(const void* q, const void* k, const void* v, const float* bias_rev,
                                           const float* bias_max, void* out, int B, int H, int W, int C, int nH, int ws,
                                           int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C == nH * 32, "swin attention: head_dim must be 32");
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0, "swin attention: window must tile the token grid");
  MV_CHECK_ARG(shift == 0 || shift == ws / 2, "swin attention: shift must be 0 or ws/2");
  AttnParams p{};
  p.Nq = p.Nkv = ws * ws;
  p.nH = nH;
  p.bias_rev = bias_rev;
  p.bias_max = bias_max;
  p.H = H; p.W = W; p.shift = shift; p.C = C;
  p.kv_len = nullptr;
  p.out = out;
  const int n_bh = B * (H / ws) * (W / ws) * nH;
  switch (ws) {
    case 28: return launch_attn<MODE_SWIN, 32, 28, 112, true>(q, k, v, n_bh, p, stream);
    case 14: return launch_attn<MODE_SWIN, 32, 14, 112, true>(q, k, v, n_bh, p, stream);
    case 7: return launch_attn<MODE_SWIN, 32, 7, 112, true>(q, k, v, n_bh, p, stream);
    default: return mv::fail(-1, "swin attention: window %d not instantiated (7, 14, 28)", ws);
  }
}

extern "C" int mvuld_seq_attention(const void* q, const void* k, const void* v, const int* kv_len, void* out, int B,
                                   int L, int nH, int hd, cudaStream_t stream) {
  MV_CHECK_ARG(hd == 64, "seq attention: head_dim 64 only");
  MV_CHECK_ARG(L <= 512 && L % 8 == 0, "seq attention: L must be <= 512 and a multiple of 8");
  AttnParams p{};
  p.Nq = p.Nkv = L;
  p.nH = nH;
  p.C = nH * hd;
  p.kv_len = kv_len;
  p.out = out;
  return launch_attn<MODE_SEQ, 64, 1, 128, false>(q, k, v, B * nH, p, stream);
}
