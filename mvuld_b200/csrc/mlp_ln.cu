// Fused SwinV2 Mlp + res-post-norm for the narrow stages (C = 128, 256) on sm_100a:
//
//     x = x + LayerNorm(fc2(GELU(fc1(xb) + b1)) + b2) * gamma + beta
//
// (swin_transformer_v2.py:26-32 Mlp, :304 `x = x + self.drop_path(self.norm2(self.mlp(x)))`)  As two
// kernels (gemm_tn_kernel<EpiBf16Tma> + gemm_ln_kernel) the hidden activation [M, 4C] bf16 is written to HBM by fc1 and
// read back by fc2: at stage 0 of a 64-image batch (M = 802 816, C = 128) that is 2 x 822 MB of the 2.9 GB the pair
// moves.  Here a CTA owns 128 rows and the hidden activation never leaves the SM:
//
//   for each 128-wide chunk j of the hidden dimension:
//     G1(j): acc1[j & 1]  = X[128, C] . W1[128 j .. 128 j + 128, :]^T                  (TMEM, 128 fp32 columns, x2)
//     GELU : acc1 -> + b1 -> erf-GELU -> bf16 -> H[j & 1] in shared memory, written directly in the K-major
//            128B-swizzled layout a tcgen05 A operand is read in (thread = row: its 64 columns ARE one 128-byte row)
//     G2(j): acc2 += H[j & 1][128, 128] . W2[:, 128 j .. 128 j + 128]^T                (TMEM, C fp32 columns)
//   LayerNorm epilogue on acc2 (the gemm_ln_kernel epilogue: statistics on the fp32 accumulator, affine, + shortcut).
//
//   warp 0      TMA producer : the X tile (resident for the whole row tile) and the weights as 16 KB units
//                              [128 rows x 64 k] through an NW-deep ring, in the order the MMA warp consumes them
//   warp 1      MMA issuer   : issues G1(g + 1) BEFORE G2(g), so the tensor pipe works on the next chunk while the
//                              epilogue warps run the GELU of this one
//   warps 2..9  GELU warps   : acc1 -> bias + GELU -> bf16 -> H, one chunk after the other
//   warps 10..17 LN warps    : the LayerNorm + residual epilogue of row tile t while the GELU warps are already on row
//                              tile t + 1 (with one set of eight warps doing both, every phase -- each latency bound on
//                              two warps per scheduler -- added up: 31 k cycles per row tile against 27 k for the two
//                              kernels)
//
// Measured (64 images; tools/time_mlp.py): C = 128 454 us against 624 us for the two kernels, C = 256 302 against 350.
// What bounds it (variant builds): with the GELU arithmetic removed 317 us, with the LayerNorm stores removed 421, with
// both 258 -- the skeleton itself (TMEM reads, H hand-off through the proxy fence, and 256 KB of weights per row tile
// streamed from L2 by every SM = 6 TB/s of L2 -> SM traffic against the ~12 TB/s the fabric gives) is the floor, and the
// GELU warps' time adds to it because G2(g) cannot be issued before GELU(g) is done.  Tried and not kept: the GELU warps
// as two groups of four on alternate chunks (each group then waits for a G1 that sits behind the other group's G2 in the
// MMA warp's in-order issue: 505 us); one set of eight warps for GELU and LayerNorm (667 us); sixteen GELU warps at 32
// columns per thread, with eight LayerNorm warps (72-register cap: 467 / 374 us) or four whole-row LayerNorm warps
// (603 / 441 us: the LayerNorm epilogue becomes the longest stage).
//
// The accumulation order over K is that of the two-kernel path (k blocks of 64 in ascending order), and the GELU / LN
// arithmetic is the same code, so the result is bit-identical to gemm + gemm_ln (tests/test_gpu_kernels.py).
// Algorithmic HBM bytes per row: 2 C (xb in) + 4 C (shortcut) + 4 C (x32 out) + 2 C (xb out) = 12 C, against
// 12 C + 2 * 8 C = 28 C for the two kernels.
#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int ML_BM = 128;
constexpr int ML_HC = 128;                     // hidden columns per chunk
constexpr int ML_THREADS = 64 + 256 + 256;     // TMA, MMA, 8 GELU warps, 8 LayerNorm warps
constexpr int ML_UNIT = 128 * 64 * 2;          // 16 KB: [128 rows x 64 k] bf16, K-major, 128B swizzle

template <int C>
struct MlpCfg {
  static constexpr int HID = 4 * C;
  static constexpr int NCH = HID / ML_HC;      // chunks per row tile
  static constexpr int KB1 = C / 64;           // k blocks of G1 = weight units of G1 per chunk
  static constexpr int NH = C / 128;           // 128-column halves of the G2 accumulator
  static constexpr int NACC2 = (C == 128) ? 2 : 1;
  static constexpr int X_BYTES = KB1 * ML_UNIT;
  static constexpr int H_BYTES = 2 * ML_UNIT;  // one H buffer: 128 x 128 bf16 = two k blocks
  static constexpr int NW = 5 - 2 * (NH - 1);  // weight ring depth: what is left of the 227 KB (5 at C = 128, 3 at 256)
  static constexpr int STATS_BYTES = 2 * 2 * ML_BM * 8;
  static constexpr int XPOSE_BYTES = 8 * 32 * 128;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = X_BYTES + 2 * H_BYTES + NW * ML_UNIT + STATS_BYTES + XPOSE_BYTES + BAR_BYTES + 1024;
  static constexpr int ACC2_COL = 2 * ML_HC;   // acc1[0], acc1[1] at columns 0 / 128, acc2 from 256
};

struct MlpParams {
  const float* b1;         // [4C]
  const float* b2;         // [C]
  const float* gamma;      // [C]
  const float* beta;       // [C]
  const float* shortcut;   // fp32 [M, C] (may alias x32)
  float* x32;              // fp32 [M, C] or null
  bf16* xb;                // bf16 [M, C] or null (may alias the X operand: a row tile is read before it is written,
                           // and by no other CTA)
  float eps;
};

template <int C>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_ln_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, int M, MlpParams ep) {
  using Cfg = MlpCfg<C>;
  constexpr int NCH = Cfg::NCH, KB1 = Cfg::KB1, NH = Cfg::NH, NW = Cfg::NW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sX = smem;
  uint8_t* sH = sX + Cfg::X_BYTES;
  uint8_t* sW = sH + 2 * Cfg::H_BYTES;
  float2* sStats = reinterpret_cast<float2*>(sW + NW * ML_UNIT);                   // [parity][half][row]
  uint8_t* sT = reinterpret_cast<uint8_t*>(sStats) + Cfg::STATS_BYTES;             // [warp][32 rows][8 x 16 B, swizzled]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sT + Cfg::XPOSE_BYTES);
  uint64_t* w_empty = w_full + NW;
  uint64_t* x_full = w_empty + NW;
  uint64_t* x_empty = x_full + 1;
  uint64_t* a1_full = x_empty + 1;
  uint64_t* a1_empty = a1_full + 2;
  uint64_t* h_full = a1_empty + 2;
  uint64_t* h_empty = h_full + 2;
  uint64_t* a2_full = h_empty + 2;
  uint64_t* a2_empty = a2_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = (M + ML_BM - 1) / ML_BM;
  const int nt = ((int)blockIdx.x < num_tiles) ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int G = nt * NCH;                       // chunks this CTA runs
  constexpr int EPI_THREADS = 256;              // GELU threads = LayerNorm threads

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    for (int s = 0; s < NW; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&a1_full[a], 1);
      mbar_init(&a1_empty[a], EPI_THREADS);
      mbar_init(&h_full[a], EPI_THREADS);
      mbar_init(&h_empty[a], 1);
      mbar_init(&a2_full[a], 1);
      mbar_init(&a2_empty[a], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      auto load_w = [&](const CUtensorMap* tm, int c0, int c1) {
        mbar_wait(&w_empty[slot], ph ^ 1, 1);
        mbar_arrive_expect_tx(&w_full[slot], ML_UNIT);
        tma_load_2d(sW + slot * ML_UNIT, tm, &w_full[slot], c0, c1);
        if (++slot == NW) { slot = 0; ph ^= 1; }
      };
      auto load_g1 = [&](int g) {
        const int tl = g / NCH, j = g - tl * NCH;
        if (j == 0) {                                               // the row tile's X, once the previous tile's G1s are done
          const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
          mbar_wait(x_empty, (tl & 1) ^ 1, 5);
          mbar_arrive_expect_tx(x_full, Cfg::X_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sX + kb * ML_UNIT, &tmX, x_full, kb * 64, tile * ML_BM);
        }
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb) load_w(&tmW1, kb * 64, j * ML_HC);
      };
      auto load_g2 = [&](int g) {
        const int j = g % NCH;
#pragma unroll
        for (int kb2 = 0; kb2 < 2; ++kb2)
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) load_w(&tmW2, j * ML_HC + kb2 * 64, nh * 128);
      };
      if (G > 0) load_g1(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) load_g1(g + 1);
        load_g2(g);
      }
    }
  } else if (warp == 1) {
    // uniform control flow, one elected lane issues, descriptors advanced by adds (see gemm.cu)
    constexpr uint32_t idesc = make_idesc_bf16(ML_BM, 128, 0, 0);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    const uint32_t x_lo0 = (uint32_t)make_smem_desc(smem_u32(sX), 16, 1024, 2);
    const uint32_t h_lo0 = (uint32_t)make_smem_desc(smem_u32(sH), 16, 1024, 2);
    const uint32_t w_lo0 = (uint32_t)make_smem_desc(smem_u32(sW), 16, 1024, 2);
    const bool leader = elect_one();
    int slot = 0;
    uint32_t ph = 0;
    // one 128 x 128 x 64 product: A block at a_lo, B = the next weight unit of the ring
    auto unit_mma = [&](uint32_t d_tmem, uint32_t a_lo, bool first) {
      mbar_wait(&w_full[slot], ph, 3);
      tc_fence_after();
      if (leader) {
        const uint32_t b_lo = w_lo0 + (uint32_t)slot * (ML_UNIT >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(d_tmem, ((uint64_t)desc_hi << 32) | (a_lo + k * 2), ((uint64_t)desc_hi << 32) | (b_lo + k * 2), idesc,
                  !(first && k == 0));
        umma_commit(&w_empty[slot]);
      }
      __syncwarp();
      if (++slot == NW) { slot = 0; ph ^= 1; }
    };
    auto issue_g1 = [&](int g) {
      const int tl = g / NCH, j = g - tl * NCH;
      const int b = g & 1;
      if (j == 0) mbar_wait(x_full, tl & 1, 6);
      mbar_wait(&a1_empty[b], ((g >> 1) & 1) ^ 1, 2);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB1; ++kb) unit_mma(tmem_base + b * ML_HC, x_lo0 + (uint32_t)kb * (ML_UNIT >> 4), kb == 0);
      if (leader) {
        umma_commit(&a1_full[b]);
        if (j == NCH - 1) umma_commit(x_empty);                     // the X tile may be replaced
      }
      __syncwarp();
    };
    auto issue_g2 = [&](int g) {
      const int tl = g / NCH, j = g - tl * NCH;
      const int b = g & 1;
      const int a = Cfg::NACC2 == 2 ? (tl & 1) : 0;
      const uint32_t apar = Cfg::NACC2 == 2 ? ((tl >> 1) & 1) : (tl & 1);
      mbar_wait(&h_full[b], (g >> 1) & 1, 7);
      if (j == 0) mbar_wait(&a2_empty[a], apar ^ 1, 8);
      tc_fence_after();
#pragma unroll
      for (int kb2 = 0; kb2 < 2; ++kb2)
#pragma unroll
        for (int nh = 0; nh < NH; ++nh)
          unit_mma(tmem_base + Cfg::ACC2_COL + a * C + nh * 128,
                   h_lo0 + (uint32_t)(b * 2 + kb2) * (ML_UNIT >> 4), j == 0 && kb2 == 0);
      if (leader) {
        umma_commit(&h_empty[b]);
        if (j == NCH - 1) umma_commit(&a2_full[a]);
      }
      __syncwarp();
    };
    if (G > 0) issue_g1(0);
    for (int g = 0; g < G; ++g) {
      if (g + 1 < G) issue_g1(g + 1);
      issue_g2(g);
    }
  } else if (warp < 10) {
    // ---- GELU warps: acc1[g & 1] (this thread's row, 64 columns) -> H[g & 1] ----
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;            // row within the tile
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const int b = g & 1;
      const int j = g % NCH;
      const float* bias = ep.b1 + j * ML_HC + half * 64;
      mbar_wait(&a1_full[b], (g >> 1) & 1, 4);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + b * ML_HC + half * 64;
      uint32_t v[2][32];
      tmem_ld32(t0, v[0]);
      tmem_ld32(t0 + 32, v[1]);
      mbar_wait(&h_empty[b], ((g >> 1) & 1) ^ 1, 9);   // G2(g - 2) has read this H buffer
      uint8_t* dst = sH + b * Cfg::H_BYTES + half * ML_UNIT + r * 128;
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&a1_empty[b]);                       // both halves are in registers: G1(g + 2) may overwrite
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {                  // 16-byte unit hh * 4 + u of the 128-byte row: 8 columns
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + hh * 32 + u * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + hh * 32 + u * 8 + 4));
          float x0 = __uint_as_float(v[hh][8 * u]) + b0.x, x1 = __uint_as_float(v[hh][8 * u + 1]) + b0.y;
          float x2 = __uint_as_float(v[hh][8 * u + 2]) + b0.z, x3 = __uint_as_float(v[hh][8 * u + 3]) + b0.w;
          float x4 = __uint_as_float(v[hh][8 * u + 4]) + b1.x, x5 = __uint_as_float(v[hh][8 * u + 5]) + b1.y;
          float x6 = __uint_as_float(v[hh][8 * u + 6]) + b1.z, x7 = __uint_as_float(v[hh][8 * u + 7]) + b1.w;
          gelu_erf2(x0, x1);
          gelu_erf2(x2, x3);
          gelu_erf2(x4, x5);
          gelu_erf2(x6, x7);
          *reinterpret_cast<uint4*>(dst + ((((hh << 2) + u) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3), pack_bf16x2(x4, x5), pack_bf16x2(x6, x7));
        }
      }
      fence_proxy_async_smem();                        // generic-proxy writes -> visible to the tensor core's reads
      mbar_arrive(&h_full[b]);
    }
  } else {
    // ---- LayerNorm warps: LayerNorm + residual epilogue of row tile tl (the epilogue of gemm_ln_kernel, N = C) ----
    const int quarter = warp & 3;
    const int half = (warp - 10) >> 2;
    const int r = quarter * 32 + lane;
    constexpr int HCOLS = C / 2;                       // columns per thread
    const float invN = 1.0f / (float)C;
    // lane l handles float4 column c4 = l % 8 of rows rl = 4 k + l / 8, k = 0..7, of this warp's 32 rows
    const int c4 = lane & 7, rsub = lane >> 3;
    uint8_t* myT = sT + (warp - 10) * 4096;
    const int colbase = half * HCOLS;
    // The residual rows are the only HBM operand of this epilogue and do not depend on the MMAs: the lines of row tile
    // tl + 1 are pulled into L2 while the warps wait for the accumulator of row tile tl (no registers held; the loads
    // under the TMEM read then cost an L2 hit instead of a DRAM round trip per 32-column chunk).
    auto prefetch_shortcut = [&](int tl) {
      if (!ep.shortcut || tl >= nt || c4 != 0) return;
      const int row0 = ((int)blockIdx.x + tl * (int)gridDim.x) * ML_BM + quarter * 32;
#pragma unroll
      for (int cc = 0; cc < HCOLS / 32; ++cc)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int grow = row0 + 4 * k + rsub;
          if (grow < M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.shortcut + (size_t)grow * C + colbase + cc * 32));
        }
    };
    prefetch_shortcut(0);
#pragma unroll 1
    for (int tl = 0; tl < nt; ++tl) {
      const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
      prefetch_shortcut(tl + 1);
      const int a = Cfg::NACC2 == 2 ? (tl & 1) : 0;
      const uint32_t apar = Cfg::NACC2 == 2 ? ((tl >> 1) & 1) : (tl & 1);
      const int wrow0 = tile * ML_BM + quarter * 32;
      mbar_wait(&a2_full[a], apar, 10);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + Cfg::ACC2_COL + a * C + half * HCOLS;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < HCOLS; c += 32) {
        uint32_t v[32];
        tmem_ld32(t0 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.b2 + colbase + c + i));
          const float x0 = __uint_as_float(v[i]) + b.x, x1 = __uint_as_float(v[i + 1]) + b.y;
          const float x2 = __uint_as_float(v[i + 2]) + b.z, x3 = __uint_as_float(v[i + 3]) + b.w;
          s1 += (x0 + x1) + (x2 + x3);
          s2 = fmaf(x0, x0, s2); s2 = fmaf(x1, x1, s2); s2 = fmaf(x2, x2, s2); s2 = fmaf(x3, x3, s2);
        }
      }
      float2* st = sStats + (tl & 1) * 2 * ML_BM;
      st[half * ML_BM + r] = make_float2(s1, s2);
      named_bar_sync(1 + quarter, 64);
      const float2 other = st[(half ^ 1) * ML_BM + r];
      // half 0 adds (own, other), half 1 (other, own): the same order in both threads of a row
      const float t1 = half == 0 ? s1 + other.x : other.x + s1, t2 = half == 0 ? s2 + other.y : other.y + s2;
      const float mean = t1 * invN;
      const float var = fmaxf(t2 * invN - mean * mean, 0.f);
      const float rstd = rsqrtf(var + ep.eps);
      const float nmr = -mean * rstd;
#pragma unroll 1
      for (int cc = 0; cc < HCOLS / 32; ++cc) {
        const int col = colbase + cc * 32;
        uint32_t v[32];
        tmem_ld32(t0 + cc * 32, v);
        float4 sc[8];                                  // the residual rows, requested under the TMEM load
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int grow = wrow0 + 4 * k + rsub;
          sc[k] = (ep.shortcut && grow < M) ? *reinterpret_cast<const float4*>(ep.shortcut + (size_t)grow * C + col + 4 * c4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        if (cc + 1 == HCOLS / 32) {
          tc_fence_before();
          mbar_arrive(&a2_empty[a]);                   // the accumulator is in registers: a later tile's G2 may start
        }
        __syncwarp();                                  // previous chunk's reads of myT are done
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.b2 + col + 4 * i));
          float4 o;
          o.x = fmaf(__uint_as_float(v[4 * i]) + b.x, rstd, nmr);
          o.y = fmaf(__uint_as_float(v[4 * i + 1]) + b.y, rstd, nmr);
          o.z = fmaf(__uint_as_float(v[4 * i + 2]) + b.z, rstd, nmr);
          o.w = fmaf(__uint_as_float(v[4 * i + 3]) + b.w, rstd, nmr);
          *reinterpret_cast<float4*>(myT + lane * 128 + ((i ^ (lane & 7)) << 4)) = o;
        }
        __syncwarp();
        const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + col + 4 * c4));
        const float4 be = __ldg(reinterpret_cast<const float4*>(ep.beta + col + 4 * c4));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rl = 4 * k + rsub;
          const int grow = wrow0 + rl;
          const float4 a4 = *reinterpret_cast<const float4*>(myT + rl * 128 + ((c4 ^ (rl & 7)) << 4));
          if (grow < M) {
            const size_t off = (size_t)grow * C + col + 4 * c4;
            const float4 o = make_float4(fmaf(a4.x, g.x, be.x) + sc[k].x, fmaf(a4.y, g.y, be.y) + sc[k].y,
                                         fmaf(a4.z, g.z, be.z) + sc[k].z, fmaf(a4.w, g.w, be.w) + sc[k].w);
            if (ep.x32) *reinterpret_cast<float4*>(ep.x32 + off) = o;
            if (ep.xb) *reinterpret_cast<uint2*>(ep.xb + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int C>
static int launch_mlp_ln(const void* X, const void* W1, const void* W2, int M, const MlpParams& ep, cudaStream_t stream) {
  using Cfg = MlpCfg<C>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(Cfg::ACC2_COL + Cfg::NACC2 * C <= 512, "TMEM budget");
  CUtensorMap tmX, tmW1, tmW2;
  const uint32_t box[2] = {64, 128};
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)C * 2};
    int rc = make_tmap_16b(&tmX, X, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)Cfg::HID};
    uint64_t str[1] = {(uint64_t)C * 2};
    int rc = make_tmap_16b(&tmW1, W1, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Cfg::HID, (uint64_t)C};
    uint64_t str[1] = {(uint64_t)Cfg::HID * 2};
    int rc = make_tmap_16b(&tmW2, W2, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  auto kern = mlp_ln_kernel<C>;
  static unsigned long long attr_set = 0;   // per template instantiation, one bit per device
  if (first_use_on_current_device(&attr_set)) {
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int tiles = (M + ML_BM - 1) / ML_BM;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, ML_THREADS, Cfg::SMEM_BYTES, stream>>>(tmX, tmW1, tmW2, M, ep);
  MV_LAUNCH_OK();
  return 0;
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_mlp_ln_bf16(const void* X, const void* W1, const float* b1, const void* W2, const float* b2,
                                 const float* gamma, const float* beta, float eps, const float* shortcut_f32, float* x32,
                                 void* xb, int M, int C, cudaStream_t stream) {
  MV_CHECK_ARG(M > 0, "mlp_ln: empty problem M=%d", M);
  MV_CHECK_ARG(X && W1 && W2 && b1 && b2 && gamma && beta && (x32 || xb), "mlp_ln: null pointer");
  MV_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)W1 % 16) == 0 && ((uintptr_t)W2 % 16) == 0,
               "mlp_ln: operands must be 16-byte aligned");
  MlpParams ep;
  ep.b1 = b1; ep.b2 = b2; ep.gamma = gamma; ep.beta = beta; ep.shortcut = shortcut_f32; ep.x32 = x32;
  ep.xb = reinterpret_cast<bf16*>(xb); ep.eps = eps;
  switch (C) {
    case 128: return launch_mlp_ln<128>(X, W1, W2, M, ep, stream);
    case 256: return launch_mlp_ln<256>(X, W1, W2, M, ep, stream);
    default: return mv::fail(-1, "mlp_ln: C = %d not instantiated (128, 256)", C);
  }
}
