// tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (bf16 in, fp32 accumulate in TMEM).
//
// Replaces the cuBLAS calls behind every nn.Linear / 1x1 Conv1d on the MVulD hot path (SURVEY.md K1, K6, K7,
// K9, K12-K14, K16, K17, K19): swin_transformer_v2.py:150,177,27-30,361; unixcoder.py:36 (HF RobertaModel);
// GraphModel.py:167-177,186-187; Rs_GCN.py:57-71; DGL GATConv.fc / GatedGraphConv.linears / GRUCell.
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0      TMA producer   : A/W tiles (128B-swizzled, K-major) -> STAGES-deep smem ring, mbarrier full/empty
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage, commits to mbarriers
//   warps 2..9  epilogue       : tcgen05.ld 32 lanes x 32 columns, fused epilogue, 128-bit global stores
//                                (two warps per TMEM lane quarter, each draining half of the tile's columns)
//   TMEM        2 x BN fp32 columns: the epilogue of tile i overlaps the mainloop of tile i+1
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;   // two per TMEM lane quarter (16 warps measured 15 % slower on the GELU GEMMs: the
                                    // register cap of 576 threads costs more than the extra latency hiding gains)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

// ------------------------------------------------------------------------------------------------
// Epilogues.  Called once per (row, 32-column chunk) by the thread that owns the row.
// ------------------------------------------------------------------------------------------------
// 256-bit global store of eight 32-bit words (sm_100 PTX): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void st_global_v8_b32(void* p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                                                 uint32_t w5, uint32_t w6, uint32_t w7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3),
               "r"(w4), "r"(w5), "r"(w6), "r"(w7)
               : "memory");
}

struct EpiGeneric {
  static constexpr bool kStaged = false;
  const float* bias;   // [N] or null
  const float* res;    // fp32 [M, ldr] or null; added after the activation
  bf16* out_b;         // bf16 [M, ldc] or null
  float* out_f;        // fp32 [M, ldc] or null
  int ldr, ldc, act;   // act: 0 none, 1 GELU(erf), 2 ELU
  int M, N;
  __device__ __forceinline__ void operator()(int row, int col0, const uint32_t (&r)[32]) const {
    if (row >= M || col0 >= N) return;
    float v[32];
    const bool full = (col0 + 32 <= N);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (bias) {
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
      } else {
        for (int i = 0; i < 32; ++i) if (col0 + i < N) v[i] += bias[col0 + i];
      }
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf2(v[i], v[i + 1]);
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = elu1(v[i]);
    }
    if (res) {
      const float* rp = res + (size_t)row * ldr + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(rp + i));
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
      } else {
        for (int i = 0; i < 32; ++i) if (col0 + i < N) v[i] += rp[i];
      }
    }
    if (out_b) {
      bf16* op = out_b + (size_t)row * ldc + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(v[i], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
          u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      } else {
        for (int i = 0; i < 32; ++i) if (col0 + i < N) op[i] = __float2bfloat16(v[i]);
      }
    }
    if (out_f) {
      float* op = out_f + (size_t)row * ldc + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
        for (int i = 0; i < 32; ++i) if (col0 + i < N) op[i] = v[i];
      }
    }
  }
};

// bias + activation -> bf16, written through shared memory and a TMA store.  The thread-per-row epilogues above store
// 16 bytes per lane into 32 different rows per instruction (32 L1 transactions); for the K <= 512 layers (fc1, proj,
// RoBERTa dense) that store stream, not the MMA, set the tile time.  Here each warp stages a [32 rows x 64 columns]
// bf16 slab (128-byte rows, 128B-swizzled: conflict-free STS) and one lane hands it to the TMA engine, which also
// clips the M / N tails.
struct EpiBf16Tma {
  static constexpr bool kStaged = true;
  const float* bias;   // [N] or null
  int act;             // 0 none, 1 GELU(erf), 2 ELU
  int M, N;
  // values of one 32-column chunk -> packed bf16 (16 words)
  __device__ __forceinline__ void compute(int col0, const uint32_t (&r)[32], uint32_t (&w)[16]) const {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (bias) {
      if (col0 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (col0 + i < N) v[i] += bias[col0 + i];
      }
    }
    if (act == 1) {
#ifdef MV_GELU_SCALAR
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
#else
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf2(v[i], v[i + 1]);
#endif
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = elu1(v[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  }
};

// SwinV2 qkv epilogue (swin_transformer_v2.py:147-157 + the window partition / cyclic shift of :276-286):
//   + cat(q_bias, 0, v_bias); q,k L2-normalised per head (F.normalize eps 1e-12); q additionally scaled by
//   exp(min(logit_scale, ln 100)) * log2(e) so the attention kernel's exp2 needs no multiply;
//   rows scattered to window-major head-major [B*nW, nH, ws*ws, 32]; q,k stored fp16 (|q| <= 145), v bf16.
struct EpiQkvSwin {
  static constexpr bool kStaged = false;
  const float* q_bias;   // [C]
  const float* v_bias;   // [C]
  const float* qscale;   // [nH]
  __half* q;
  __half* k;
  bf16* v;
  float* rq = nullptr;   // training: 1 / max(|q|, eps) of every (token, head) in the window-major order of q (or null)
  float* rk = nullptr;
  int C, nH, H, W, ws, shift, M;
  __device__ __forceinline__ void operator()(int row, int col0, const uint32_t (&r)[32]) const {
    if (row >= M) return;
    const int which = col0 / C;
    const int cc = col0 - which * C;
    const int head = cc >> 5;
    float v32[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v32[i] = __uint_as_float(r[i]);
    if (which != 1) {
      const float* bp = (which == 0 ? q_bias : v_bias) + cc;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(bp + i));
        v32[i] += b.x; v32[i + 1] += b.y; v32[i + 2] += b.z; v32[i + 3] += b.w;
      }
    }
    float rnorm = 0.f;
    if (which < 2) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) ss += v32[i] * v32[i];
      float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      rnorm = inv;
      if (which == 0) inv *= __ldg(qscale + head);
#pragma unroll
      for (int i = 0; i < 32; ++i) v32[i] *= inv;
    }
    // token -> (window, slot) of the cyclically shifted image
    const int HW = H * W;
    const int b = row / HW;
    const int t = row - b * HW;
    int hh = t / W, ww = t - hh * W;
    hh -= shift; if (hh < 0) hh += H;
    ww -= shift; if (ww < 0) ww += W;
    const int nWw = W / ws;
    const int win = (hh / ws) * nWw + (ww / ws);
    const int slot = (hh % ws) * ws + (ww % ws);
    const int nW = (H / ws) * nWw;
    const size_t dst = ((((size_t)b * nW + win) * nH + head) * (size_t)(ws * ws) + slot) * 32;
    if (which == 0 && rq) rq[dst >> 5] = rnorm;
    if (which == 1 && rk) rk[dst >> 5] = rnorm;
    // one head of one token = 64 contiguous bytes = two 256-bit stores
    uint32_t w[16];
    if (which == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v32[2 * i], v32[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __half2 h = __floats2half2_rn(v32[2 * i], v32[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    uint8_t* op = which == 2 ? reinterpret_cast<uint8_t*>(v + dst)
                             : reinterpret_cast<uint8_t*>((which == 0 ? q : k) + dst);
    st_global_v8_b32(op, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
    st_global_v8_b32(op + 32, w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]);
  }
};

// Head-major qkv epilogue for the RoBERTa encoder (HF RobertaSelfAttention as called from unixcoder.py:36):
// columns [0,Hd) = query, [Hd,2Hd) = key, [2Hd,3Hd) = value; + bias; q scaled by log2(e)/sqrt(hd);
// rows scattered to [B, nH, L, hd] (bf16).
struct EpiQkvHeads {
  static constexpr bool kStaged = false;
  const float* bias;   // [3*Hd]
  bf16* q;
  bf16* k;
  bf16* v;
  float qmul;
  int Hd, nH, hd, L, M;
  __device__ __forceinline__ void operator()(int row, int col0, const uint32_t (&r)[32]) const {
    if (row >= M) return;
    const int which = col0 / Hd;
    const int cc = col0 - which * Hd;
    const int head = cc / hd;
    const int d0 = cc - head * hd;
    float v32[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
      v32[i] = __uint_as_float(r[i]) + b.x; v32[i + 1] = __uint_as_float(r[i + 1]) + b.y;
      v32[i + 2] = __uint_as_float(r[i + 2]) + b.z; v32[i + 3] = __uint_as_float(r[i + 3]) + b.w;
    }
    if (which == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v32[i] *= qmul;
    }
    const int b = row / L, t = row - b * L;
    bf16* base = which == 0 ? q : (which == 1 ? k : v);
    uint8_t* op = reinterpret_cast<uint8_t*>(base + ((((size_t)b * nH + head) * L + t) * hd + d0));
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v32[2 * i], v32[2 * i + 1]);
    st_global_v8_b32(op, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
    st_global_v8_b32(op + 32, w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]);
  }
};

// GRUCell epilogue of DGL GatedGraphConv (torch.nn.GRUCell, gate order r, z, n): the GEMM computes, for node row n and
// feature j, the four pre-activations in ADJACENT columns 4j .. 4j+3 of  [a | h] (K = 2D) x Wg^T:
//   4j   r : W_ir a + W_hr h      4j+1  z : W_iz a + W_hz h      4j+2  i_n : W_in a      4j+3  h_n : W_hn h
// (bias4 holds b_ir + b_hr, b_iz + b_hz, b_in, b_hn interleaved the same way), so one thread owns all it needs for
//   h' = (1 - z) * tanh(i_n + r * h_n) + z * h.
// h32 (fp32 state) is updated in place; the bf16 shadow goes to ANOTHER buffer than the one the A operand is read from
// (other tiles of the same rows are still loading it).  Replaces two GEMMs writing gi / gh (2 x N x 3D bf16) and the
// element-wise gate kernel that read them back.
// 256-bit global accesses (sm_100 PTX): one full 32-byte sector per lane and instruction.  Two 128-bit stores to the
// halves of a sector reached L2 as partial writes (ncu: 0.66 GB of extra DRAM reads per launch to fill them).
__device__ __forceinline__ void ld_global_v8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_global_v8(float* p, const float* o) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]),
               "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
               : "memory");
}

struct EpiGru {
  static constexpr bool kStaged = false;
  static constexpr bool kPrefetch = true;
  struct Pre { float4 h0, h1; };
  const float* bias4;  // [4D]
  float* h32;          // [M, D]
  bf16* hb;            // [M, ldhb]
  int ldhb, M, D;
  // the old state of the 8 features of this chunk, requested BEFORE the wait on the accumulator: the HBM latency of
  // this read would otherwise sit in the epilogue's critical path twice per tile
  __device__ __forceinline__ void prefetch(int row, int col0, Pre& p) const {
    if (row >= M || col0 >= 4 * D) return;
    ld_global_v8(h32 + (size_t)row * D + (col0 >> 2), p.h0, p.h1);
  }
  __device__ __forceinline__ void operator()(int row, int col0, const uint32_t (&r)[32], const Pre& p) const {
    if (row >= M || col0 >= 4 * D) return;
    const int j0 = col0 >> 2;
    float* hp = h32 + (size_t)row * D + j0;
    const float h[8] = {p.h0.x, p.h0.y, p.h0.z, p.h0.w, p.h1.x, p.h1.y, p.h1.z, p.h1.w};
    float o[8];
    constexpr float L2E = 1.4426950408889634f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias4 + col0) + q);
      const float pr = __uint_as_float(r[4 * q]) + b.x, pz = __uint_as_float(r[4 * q + 1]) + b.y;
      const float pi = __uint_as_float(r[4 * q + 2]) + b.z, ph = __uint_as_float(r[4 * q + 3]) + b.w;
      const float rg = rcp_approx(1.0f + ex2_approx(-pr * L2E));
      const float zg = rcp_approx(1.0f + ex2_approx(-pz * L2E));
      const float nn = 1.0f - 2.0f * rcp_approx(1.0f + ex2_approx(2.0f * L2E * (pi + rg * ph)));   // tanh
      o[q] = (1.0f - zg) * nn + zg * h[q];
    }
    st_global_v8(hp, o);
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(hb + (size_t)row * ldhb + j0) = w;
  }
};

// epilogues that want operands fetched before the accumulator is ready declare kPrefetch + Pre + prefetch()
template <class E, class = void>
struct epi_prefetch : std::false_type { struct Pre {}; };
template <class E>
struct epi_prefetch<E, std::void_t<decltype(E::kPrefetch)>> : std::true_type { using Pre = typename E::Pre; };

// ------------------------------------------------------------------------------------------------
// BKB > 0 selects the WEIGHT-STATIONARY schedule for skinny-K problems (K <= 64 BKB): a CTA keeps its BN x K weight
// panel resident in shared memory (loaded once), owns one column block and walks row blocks, so only A streams through
// the ring.  With the streaming schedule a 128 x 128 tile at K = 448 pulls 114 KB of A and 114 KB of W through L2 for
// 32 KB of output; the GGNN GEMMs (M = 825 k, K = 200 / 400) sat at 8.4 TB/s of L2 traffic, not at HBM or MMA limits.
//
// CG = 2 is the CTA-PAIR form (tcgen05 cta_group::2): the two CTAs of a cluster (= the two SMs of a TPC) own one
// 256 x BN tile; each CTA loads its own 128 rows of A and HALF of the tile's weight rows, the leader issues one
// M = 256 MMA over both shared memories, each CTA drains its own 128 accumulator rows.  Why: these GEMMs are bound by
// the L2 -> SM path, not by the tensor pipe or the epilogue -- a 128 x 256 tile pulls 48 KB per 64-wide k block through
// L2 for 512 cycles of MMA = 96 B/clk/SM, the chip's L2 delivers ~43 B/clk/SM (profiles/r2_ncu_gemm_fc1.md: 8.4 k cycles
// per tile against 4.1 k of MMA, = 384 KB / 45 B/clk).  The pair form moves 32 KB per CTA for the same MMA time.
template <int BN, int STAGES, int BKB = 0, bool STAGED_C = true, int CG = 1>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN / CG * GEMM_BK * 2;   // this CTA's share of the weight rows
  static constexpr int STAGE_BYTES = BKB > 0 ? A_BYTES : A_BYTES + B_BYTES;    // bytes one ring slot receives
  static constexpr int B_TOTAL = (BKB > 0 ? BKB : STAGES) * B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SLAB_BYTES = 32 * 128;                                  // [32 rows x 64 bf16] per epilogue warp
  static constexpr int SLABS_BYTES = STAGED_C ? GEMM_EPI_WARPS * SLAB_BYTES : 0;   // only the TMA-store epilogues stage C
  static constexpr int SMEM_BYTES = STAGES * A_BYTES + B_TOTAL + SLABS_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;                                    // 256 or 512 (power of two)
};

template <int BN, int STAGES, class Epi, int BKB = 0, int CG = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int K, Epi epi) {
  static_assert(CG == 1 || (CG == 2 && BKB == 0), "the pair form uses the streaming schedule");
  using Cfg = GemmCfg<BN, STAGES, BKB, Epi::kStaged, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sC = smem + STAGES * Cfg::A_BYTES + Cfg::B_TOTAL;            // staged epilogue slabs (kStaged only)
  uint64_t* full = reinterpret_cast<uint64_t*>(sC + Cfg::SLABS_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bfull = tempty + 2;                                          // weight panel resident (BKB > 0)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = CG == 2 ? (int)blockIdx.x >> 1 : (int)blockIdx.x;      // a CTA, or a CTA pair
  const int units = CG == 2 ? (int)gridDim.x >> 1 : (int)gridDim.x;
  const int m_tiles = (M + GEMM_BM * CG - 1) / (GEMM_BM * CG);            // row blocks of 128 * CG rows
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  // tile walk: streaming = tiles blockIdx.x, + gridDim.x, ... in row-major (m, n) order; weight-stationary = fixed
  // column block blockIdx.x % n_tiles, row blocks blockIdx.x / n_tiles, + gridDim.x / n_tiles, ... (the host makes
  // gridDim.x a multiple of n_tiles).  `it` counts this CTA's tiles.
  const int ws_groups = BKB > 0 ? (int)gridDim.x / n_tiles : 1;
  const int ws_n = BKB > 0 ? (int)blockIdx.x % n_tiles : 0;
  auto tile_at = [&](int it, int& m_blk, int& n_blk) -> bool {
    if (BKB > 0) {
      m_blk = (int)blockIdx.x / n_tiles + it * ws_groups;
      n_blk = ws_n;
      return m_blk < m_tiles;
    }
    const int tile = unit + it * units;
    const int mb = tile / n_tiles;
    n_blk = tile - mb * n_tiles;
    m_blk = mb * CG + cta_rank;                                           // in 128-row blocks
    return tile < num_tiles;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], (GEMM_THREADS - 64) * CG);       // pair form: both CTAs' epilogues arrive on the leader's
    }
    mbar_init(bfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int m_blk, n_blk;
      if (BKB > 0 && tile_at(0, m_blk, n_blk)) {                         // the weight panel, once
        mbar_arrive_expect_tx(bfull, num_kb * Cfg::B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_2d(sB + kb * Cfg::B_BYTES, &tmB, bfull, kb * GEMM_BK, n_blk * BN);
      }
      const uint32_t lead_full = CG == 2 ? mapa_u32(smem_u32(full), 0) : 0;   // the leader CTA's full[] barriers
      for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          if (CG == 2) {
            // both halves of the stage complete on the LEADER's barrier, which expects the pair's bytes
            if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_2d_cg2(sA + stage * Cfg::A_BYTES, &tmA, lead_full + stage * 8, kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, &tmB, lead_full + stage * 8, kb * GEMM_BK,
                            n_blk * BN + cta_rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            if (BKB == 0) tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * GEMM_BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // The whole warp runs the (uniform) control flow and one elected lane issues.  Descriptors: the high word is a
    // constant, the low word (address >> 4 | LBO << 16) advances by plain adds -- assembling each descriptor from
    // scratch inside a one-lane branch put a ~100-cycle dependent chain (shift / mask / or / elect loop) in front of
    // every MMA, as long as a 128 x 128 x 16 MMA takes in the tensor pipe.
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM * CG, BN, 0, 0);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_smem_desc(smem_u32(sA), 16, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(sB), 16, 1024, 2);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int m_blk, n_blk;
    if (BKB > 0 && tile_at(0, m_blk, n_blk)) mbar_wait(bfull, 0, 5);
    for (int local = 0; tile_at(local, m_blk, n_blk); ++local) {
      const int acc = local & 1;
      if (CG == 2) mbar_wait_cluster(&tempty[acc], ((local >> 1) & 1) ^ 1, 2);
      else mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase, 3);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)(BKB > 0 ? kb : stage) * (Cfg::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if (CG == 2)
              umma_ss2(d_tmem, ((uint64_t)desc_hi << 32) | (a_lo + k * 2), ((uint64_t)desc_hi << 32) | (b_lo + k * 2),
                       idesc, (kb | k) != 0);
            else
              umma_ss(d_tmem, ((uint64_t)desc_hi << 32) | (a_lo + k * 2), ((uint64_t)desc_hi << 32) | (b_lo + k * 2),
                      idesc, (kb | k) != 0);
          }
          if (CG == 2) umma_commit2_mc(&empty[stage], 3);   // frees the slot in BOTH CTAs when these MMAs retire
          else umma_commit(&empty[stage]);                  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) {                               // accumulator ready for the epilogue (of both CTAs)
        if (CG == 2) umma_commit2_mc(&tfull[acc], 3);
        else umma_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;    // which slice of the tile's columns this warp drains
    constexpr int PARTS = GEMM_EPI_WARPS / 4;
    constexpr int CPP = (BN / 32) / PARTS;   // 32-column chunks per warp
    int m_blk, n_blk;
    const uint32_t lead_tempty = CG == 2 ? mapa_u32(smem_u32(tempty), 0) : 0;
    for (int local = 0; tile_at(local, m_blk, n_blk); ++local) {
      const int acc = local & 1;
      const int row = m_blk * GEMM_BM + quarter * 32 + lane;
      typename epi_prefetch<Epi>::Pre pre[CPP];
      if constexpr (epi_prefetch<Epi>::value) {
#pragma unroll
        for (int i = 0; i < CPP; ++i) epi.prefetch(row, n_blk * BN + (part * CPP + i) * 32, pre[i]);
      }
      mbar_wait(&tfull[acc], (local >> 1) & 1, 4);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      // the TMEM load of the next chunk is in flight while the epilogue math of this one runs
      uint32_t r[2][32];
      tmem_ld32(t0 + part * CPP * 32, r[0]);
      if constexpr (Epi::kStaged) {
        uint8_t* slab = sC + (warp - 2) * Cfg::SLAB_BYTES;
        const int rl = lane;                                            // row inside this warp's slab
#pragma unroll
        for (int i = 0; i < CPP; ++i) {
          const int c = part * CPP + i;
          tmem_ld_wait();
          if (i + 1 < CPP) tmem_ld32(t0 + (c + 1) * 32, r[(i + 1) & 1]);
          uint32_t w[16];
          epi.compute(n_blk * BN + c * 32, r[i & 1], w);
          if ((i & 1) == 0) {
            // the previous TMA store of this warp must have finished reading the slab
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)                                   // 16-byte unit (i & 1) * 4 + u of the 128-byte row
            *reinterpret_cast<uint4*>(slab + rl * 128 + (((((i & 1) << 2) + u) ^ (rl & 7)) << 4)) =
                make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
          if ((i & 1) == 1 || i + 1 == CPP) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, slab, n_blk * BN + (c & ~1) * 32, m_blk * GEMM_BM + quarter * 32);
              tma_store_commit();
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPP; ++i) {
          const int c = part * CPP + i;
          tmem_ld_wait();
          if (i + 1 < CPP) tmem_ld32(t0 + (c + 1) * 32, r[(i + 1) & 1]);
          if constexpr (epi_prefetch<Epi>::value) epi(row, n_blk * BN + c * 32, r[i & 1], pre[i]);
          else epi(row, n_blk * BN + c * 32, r[i & 1]);
        }
      }
      tc_fence_before();
      if (CG == 2) mbar_arrive_cluster(lead_tempty + acc * 8);
      else mbar_arrive(&tempty[acc]);
    }
    if (Epi::kStaged && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores landed
  }
  tc_fence_before();
  if (CG == 2) {
    cluster_sync_all();                   // no CTA leaves (or frees TMEM) while its peer can still reach it
    if (warp == 1) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int STAGES, class Epi, int BKB = 0, int CG = 1>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const Epi& epi,
                       cudaStream_t stream, void* out_bf16 = nullptr, int ldc = 0) {
  using Cfg = GemmCfg<BN, STAGES, BKB, Epi::kStaged, CG>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "GEMM configuration exceeds the 227 KB of shared memory per CTA");
  MV_CHECK_ARG(BKB == 0 || K <= BKB * GEMM_BK, "gemm: weight-stationary schedule holds K <= %d (K=%d)", BKB * GEMM_BK, K);
  MV_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MV_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements (16 B): %d %d", lda, ldw);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap_16b(&tmA, A, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(BN / CG)};
    int rc = make_tmap_16b(&tmB, W, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  CUtensorMap tmC = tmA;          // placeholder for the thread-per-row epilogues (never dereferenced)
  if (Epi::kStaged) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldc * 2};
    uint32_t box[2] = {64, 32};
    int rc = make_tmap_16b(&tmC, out_bf16, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  auto kern = gemm_tn_kernel<BN, STAGES, Epi, BKB, CG>;
  static unsigned long long attr_set = 0;   // per template instantiation, one bit per device
  if (first_use_on_current_device(&attr_set)) {
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int n_tiles_h = (N + BN - 1) / BN;
  const int tiles = ((M + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * n_tiles_h;
  if (CG == 2) {                                   // one CTA pair per TPC, persistent over 256 x BN tiles
    const int pairs_max = num_sms() / 2;
    const int pairs = tiles < pairs_max ? tiles : pairs_max;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, M, N, K, epi));
    return 0;
  }
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (BKB > 0) {                                   // a multiple of the column blocks, at least one CTA per column block
    MV_CHECK_ARG(n_tiles_h <= num_sms(), "gemm: weight-stationary schedule needs N / %d <= SM count", BN);
    grid = (grid / n_tiles_h) * n_tiles_h;
    if (grid == 0) grid = n_tiles_h;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, M, N, K, epi);
  MV_LAUNCH_OK();
  return 0;
}

}  // namespace mv

using namespace mv;

// The pair form (CG = 2) pays above K = 512 (tools/time_gemm_ksweep.py, M = 50 176, N = 2 048: K = 1 024 162 -> 150 us,
// K = 2 048 297 -> 281 us); at K <= 512 a tile's time is its epilogue and HBM writes, and coupling two CTAs' epilogues
// to one accumulator hand-off costs more than the halved weight traffic returns (K = 256: 61 -> 82 us).
static inline bool pair_form(int K) { return K >= 1024; }

extern "C" int mvuld_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                               int act, const float* res_f32, int ldr, void* out_bf16, float* out_f32, int ldc,
                               cudaStream_t stream) {
  MV_CHECK_ARG(out_bf16 || out_f32, "gemm: no output");
  MV_CHECK_ARG(ldc % 8 == 0, "gemm: ldc must be a multiple of 8 elements");
  MV_CHECK_ARG(!res_f32 || ldr % 4 == 0, "gemm: ldr must be a multiple of 4 elements");
  const bool big = (N % 256 == 0 && (long long)M * N >= 256ll * 256 * 148);
  if (out_bf16 && !out_f32 && !res_f32 && ((uintptr_t)out_bf16 % 16) == 0) {
    EpiBf16Tma t;
    t.bias = bias; t.act = act; t.M = M; t.N = N;
    // skinny-K, tall-M problems keep the weight panel resident (see GemmCfg)
    const bool ws = K <= 256 && (long long)M >= 128ll * 2 * num_sms();
    // (measured neutral for the 128 x 256 tiles of the Swin stage-0 / 1 layers, whose GELU epilogue sets the pace)
    if (big && pair_form(K)) return launch_gemm<256, 6, EpiBf16Tma, 0, 2>(A, lda, W, ldw, M, N, K, t, stream, out_bf16, ldc);
    if (big) return launch_gemm<256, 4, EpiBf16Tma>(A, lda, W, ldw, M, N, K, t, stream, out_bf16, ldc);
    if (ws) return launch_gemm<128, 8, EpiBf16Tma, 4>(A, lda, W, ldw, M, N, K, t, stream, out_bf16, ldc);
    return launch_gemm<128, 6, EpiBf16Tma>(A, lda, W, ldw, M, N, K, t, stream, out_bf16, ldc);
  }
  EpiGeneric e;
  e.bias = bias; e.res = res_f32; e.out_b = reinterpret_cast<bf16*>(out_bf16); e.out_f = out_f32;
  e.ldr = ldr; e.ldc = ldc; e.act = act; e.M = M; e.N = N;
  // 128x256 tiles move 27 % fewer operand bytes per MAC through L2 than 128x128; use them when N fills them
  if (big && pair_form(K)) return launch_gemm<256, 7, EpiGeneric, 0, 2>(A, lda, W, ldw, M, N, K, e, stream);
  if (big) return launch_gemm<256, 4, EpiGeneric>(A, lda, W, ldw, M, N, K, e, stream);
  return launch_gemm<128, 6, EpiGeneric>(A, lda, W, ldw, M, N, K, e, stream);
}

extern "C" int mvuld_gemm_gru(const void* A, int lda, const void* Wg, int ldw, int M, int D, int K, const float* bias4,
                              float* h32, void* hb_out, int ldhb, cudaStream_t stream) {
  MV_CHECK_ARG(D % 8 == 0 && ldhb % 8 == 0, "gemm_gru: D and ldhb must be multiples of 8 (D=%d ldhb=%d)", D, ldhb);
  MV_CHECK_ARG(bias4 && h32 && hb_out, "gemm_gru: null pointer");
  MV_CHECK_ARG((const void*)hb_out != A, "gemm_gru: the bf16 state must be written to another buffer than A");
  EpiGru e;
  e.bias4 = bias4; e.h32 = h32; e.hb = reinterpret_cast<bf16*>(hb_out); e.ldhb = ldhb; e.M = M; e.D = D;
  if (K <= 448 && (long long)M >= 128ll * 2 * num_sms())
    return launch_gemm<128, 6, EpiGru, 7>(A, lda, Wg, ldw, M, 4 * D, K, e, stream);
  return launch_gemm<128, 6, EpiGru>(A, lda, Wg, ldw, M, 4 * D, K, e, stream);
}

static int swin_qkv(const void* X, const void* Wqkv, const float* q_bias, const float* v_bias, const float* qscale,
                    void* q, void* k, void* v, float* rq, float* rk, int B, int H, int W, int C, int nH, int ws,
                    int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C % 128 == 0 && C / nH == 32, "swin_qkv: need C %% 128 == 0 and head_dim 32 (C=%d nH=%d)", C, nH);
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "swin_qkv: bad window geometry");
  EpiQkvSwin e;
  e.q_bias = q_bias; e.v_bias = v_bias; e.qscale = qscale;
  e.q = reinterpret_cast<__half*>(q); e.k = reinterpret_cast<__half*>(k); e.v = reinterpret_cast<bf16*>(v);
  e.rq = rq; e.rk = rk;
  e.C = C; e.nH = nH; e.H = H; e.W = W; e.ws = ws; e.shift = shift; e.M = B * H * W;
  // (the weight-stationary schedule was tried for the narrow stages, K = C <= 256: 288 vs 284 us at C = 128, 159 vs 150
  // at C = 256 -- these launches are bound by the normalise + scatter epilogue at ~2.9 TB/s of q / k / v writes, not by
  // the operand path; tools/time_qkv.py)
  if ((3 * C) % 256 == 0) return launch_gemm<256, 4, EpiQkvSwin>(X, C, Wqkv, C, B * H * W, 3 * C, C, e, stream);
  return launch_gemm<128, 6, EpiQkvSwin>(X, C, Wqkv, C, B * H * W, 3 * C, C, e, stream);
}
extern "C" int mvuld_swin_qkv(const void* X, const void* Wqkv, const float* q_bias, const float* v_bias,
                              const float* qscale, void* q, void* k, void* v, int B, int H, int W, int C, int nH,
                              int ws, int shift, cudaStream_t stream) {
  return swin_qkv(X, Wqkv, q_bias, v_bias, qscale, q, k, v, nullptr, nullptr, B, H, W, C, nH, ws, shift, stream);
}
// training forward: also keeps 1 / max(|q|, eps) and 1 / max(|k|, eps) per (token, head) (fp32, window-major like q / k)
// for the backward of F.normalize (swin_transformer_v2.py:155)
extern "C" int mvuld_swin_qkv_train(const void* X, const void* Wqkv, const float* q_bias, const float* v_bias,
                                    const float* qscale, void* q, void* k, void* v, float* rq, float* rk, int B, int H,
                                    int W, int C, int nH, int ws, int shift, cudaStream_t stream) {
  MV_CHECK_ARG(rq && rk, "swin_qkv_train: rq / rk are null");
  return swin_qkv(X, Wqkv, q_bias, v_bias, qscale, q, k, v, rq, rk, B, H, W, C, nH, ws, shift, stream);
}

extern "C" int mvuld_heads_qkv(const void* X, const void* Wqkv, const float* bias, void* q, void* k, void* v, int B,
                               int L, int Hd, int nH, float qmul, cudaStream_t stream) {
  MV_CHECK_ARG(Hd % nH == 0 && (Hd / nH) % 32 == 0 && Hd % 32 == 0, "heads_qkv: head_dim must be a multiple of 32");
  EpiQkvHeads e;
  e.bias = bias; e.q = reinterpret_cast<bf16*>(q); e.k = reinterpret_cast<bf16*>(k); e.v = reinterpret_cast<bf16*>(v);
  e.qmul = qmul; e.Hd = Hd; e.nH = nH; e.hd = Hd / nH; e.L = L; e.M = B * L;
  if ((3 * Hd) % 256 == 0) return launch_gemm<256, 4, EpiQkvHeads>(X, Hd, Wqkv, Hd, B * L, 3 * Hd, Hd, e, stream);
  return launch_gemm<128, 6, EpiQkvHeads>(X, Hd, Wqkv, Hd, B * L, 3 * Hd, Hd, e, stream);
}
