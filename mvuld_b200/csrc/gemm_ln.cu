// tcgen05 GEMM with a full-row LayerNorm (+ residual) epilogue for sm_100a:
//
//     x[M,N] = shortcut + LayerNorm(A[M,K] W[N,K]^T + bias) * gamma + beta        (shortcut / bias optional)
//
// The SwinV2 res-post-norm pattern  x = x + norm(proj(attn))  /  x = x + norm(fc2(h))  (swin_transformer_v2.py:301,304)
// and PatchMerging's  norm(reduction(x))  (:361-362) in ONE kernel: the LayerNorm needs whole output rows, so a CTA
// owns 128 rows x all N <= 512 columns -- the fp32 accumulator is exactly the 512 TMEM columns of an SM (two
// accumulators when N <= 256, so the epilogue of tile i overlaps the mainloop of tile i+1).  Statistics are taken on
// the fp32 accumulator (the separate-kernel path rounded the GEMM output to bf16 first).  This removes one full
// HBM round trip of the activation (bf16 write + read) and one launch per LayerNorm.
//
//   warp 0      TMA producer : A box [128 x BK], W boxes [<=256 x BK] -> STAGES-deep ring (128B / 64B swizzle)
//   warp 1      MMA issuer   : tcgen05.mma M128 x N{128,256} x K16, N = 512 as two instructions per K step
//   warps 2..9  epilogue     : thread = output row; the two warps of a TMEM lane quarter split the columns, exchange
//                              (sum, sum of squares) through shared memory, then normalise on a second TMEM read
//                              with the residual row prefetched one chunk ahead
#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int GLN_BM = 128;
constexpr int GLN_THREADS = 320;

template <int N, int BK, int STAGES>
struct GemmLnCfg {
  static constexpr int ROW_BYTES = BK * 2;                          // 128 (SW128) or 64 (SW64)
  static constexpr int LAYOUT = (BK == 64) ? 2 : 4;                 // UMMA layout_type
  static constexpr int SBO = 8 * ROW_BYTES;
  static constexpr int A_BYTES = GLN_BM * ROW_BYTES;
  static constexpr int B_BYTES = N * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NACC = (2 * N <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = (NACC * N <= 256) ? 256 : 512;
  static constexpr int NI = (N > 256) ? 256 : N;                    // columns per MMA instruction
  static constexpr int STATS_BYTES = 2 /*tile parity*/ * 2 /*column half*/ * GLN_BM * 8;
  static constexpr int XPOSE_BYTES = 8 /*epilogue warps*/ * 32 * 128;   // one [32 rows x 32 fp32] chunk per warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STATS_BYTES + XPOSE_BYTES + 256 + 1024;
};

struct LnEpiParams {
  const float* bias;       // [N] or null
  const float* gamma;      // [N]
  const float* beta;       // [N]
  const float* shortcut;   // fp32 [M, N] or null (may alias x32)
  float* x32;              // fp32 [M, N] or null
  bf16* xb;                // bf16 [M, N] or null
  float eps;
};

template <int N, int BK, int STAGES>
__global__ void __launch_bounds__(GLN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int K,
               LnEpiParams ep) {
  using Cfg = GemmLnCfg<N, BK, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  float2* sStats = reinterpret_cast<float2*>(smem + STAGES * Cfg::STAGE_BYTES);     // [parity][half][row]
  uint8_t* sX = smem + STAGES * Cfg::STAGE_BYTES + Cfg::STATS_BYTES;                // [warp][32 rows][8 x 16 B, swizzled]
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + Cfg::XPOSE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = (M + GLN_BM - 1) / GLN_BM;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], GLN_THREADS - 64);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * BK, tile * GLN_BM);
#pragma unroll
          for (int h = 0; h < N / Cfg::NI; ++h)
            tma_load_2d(sB + stage * Cfg::B_BYTES + h * Cfg::NI * Cfg::ROW_BYTES, &tmB, &full[stage], kb * BK,
                        h * Cfg::NI);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // uniform control flow, one elected lane issues, descriptors advanced by adds (see gemm.cu)
    constexpr uint32_t idesc = make_idesc_bf16(GLN_BM, Cfg::NI, 0, 0);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 16, Cfg::SBO, Cfg::LAYOUT) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_smem_desc(smem_u32(sA), 16, Cfg::SBO, Cfg::LAYOUT);
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(sB), 16, Cfg::SBO, Cfg::LAYOUT);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = (Cfg::NACC == 2) ? (local & 1) : 0;
      const uint32_t par = (Cfg::NACC == 2) ? ((local >> 1) & 1) : (local & 1);
      mbar_wait(&tempty[acc], par ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase, 3);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (Cfg::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int h = 0; h < N / Cfg::NI; ++h)
              umma_ss(d_tmem + h * Cfg::NI, ((uint64_t)desc_hi << 32) | (a_lo + k * 2),
                      ((uint64_t)desc_hi << 32) | (b_lo + ((h * Cfg::NI * Cfg::ROW_BYTES) >> 4) + k * 2), idesc,
                      (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = N / 2;                     // columns per epilogue thread
    const int r = quarter * 32 + lane;            // row within the tile
    const float invN = 1.0f / (float)N;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = (Cfg::NACC == 2) ? (local & 1) : 0;
      const uint32_t par = (Cfg::NACC == 2) ? ((local >> 1) & 1) : (local & 1);
      mbar_wait(&tfull[acc], par, 4);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * N + half * HC;
      const int colbase = half * HC;
      // Global accesses of the epilogue use a transposed lane mapping -- lane = (row-in-warp % 4 ... see below) -- so
      // that one warp instruction covers four complete 128-byte row segments (thread-per-row float4 accesses touch
      // 32 different lines with 16 bytes each, which costs partial-sector traffic on both loads and stores).
      //   lane l handles float4 column c4 = l % 8 of rows rl = 4 k + l / 8, k = 0..7, of this warp's 32 rows
      const int c4 = lane & 7, rsub = lane >> 3;
      const int wrow0 = tile * GLN_BM + quarter * 32;               // first global row of this warp
      uint8_t* myX = sX + (warp - 2) * 4096;
      // the residual is the only long-latency (HBM) operand: chunk 0 is requested before the statistics pass, every
      // later chunk one iteration ahead of its use
      float4 sc[2][8];
      auto load_sc = [&](int buf, int c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int grow = wrow0 + 4 * k + rsub;
          sc[buf][k] = (ep.shortcut && grow < M)
                           ? *reinterpret_cast<const float4*>(ep.shortcut + (size_t)grow * N + colbase + c + 4 * c4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_sc(0, 0);
      // pass 1: sum and sum of squares of (acc + bias) over this thread's half of the row, two chunks per TMEM wait
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < HC; c += 64) {
        uint32_t v[64];
        tmem_ld32p(t0 + c, v);
        tmem_ld32p(t0 + c + 32, v + 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias) b = __ldg(reinterpret_cast<const float4*>(ep.bias + colbase + c + i));
          const float x0 = __uint_as_float(v[i]) + b.x, x1 = __uint_as_float(v[i + 1]) + b.y;
          const float x2 = __uint_as_float(v[i + 2]) + b.z, x3 = __uint_as_float(v[i + 3]) + b.w;
          s1 += (x0 + x1) + (x2 + x3);
          s2 = fmaf(x0, x0, s2); s2 = fmaf(x1, x1, s2); s2 = fmaf(x2, x2, s2); s2 = fmaf(x3, x3, s2);
        }
      }
      float2* st = sStats + (local & 1) * 2 * GLN_BM;
      st[half * GLN_BM + r] = make_float2(s1, s2);
      named_bar_sync(1 + quarter, 64);
      const float2 other = st[(half ^ 1) * GLN_BM + r];
      const float mean = (s1 + other.x) * invN;
      const float var = fmaxf((s2 + other.y) * invN - mean * mean, 0.f);
      const float rstd = rsqrtf(var + ep.eps);
      const float nmr = -mean * rstd;
      // pass 2: normalise in the row-owner layout, transpose the 32x32 chunk through shared memory (16-byte units
      // XOR-swizzled by row: conflict free both ways), then affine + residual + coalesced stores
      uint32_t v[2][32];
      tmem_ld32(t0, v[0]);
#pragma unroll
      for (int cc = 0; cc < HC / 32; ++cc) {
        const int c = cc * 32;
        const int col = colbase + c;
        tmem_ld_wait();
        if (cc + 1 < HC / 32) {
          tmem_ld32(t0 + c + 32, v[(cc + 1) & 1]);
          load_sc((cc + 1) & 1, c + 32);
        }
        __syncwarp();                                               // previous chunk's reads of myX are done
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4 * i));
          float4 o;
          o.x = fmaf(__uint_as_float(v[cc & 1][4 * i]) + b.x, rstd, nmr);
          o.y = fmaf(__uint_as_float(v[cc & 1][4 * i + 1]) + b.y, rstd, nmr);
          o.z = fmaf(__uint_as_float(v[cc & 1][4 * i + 2]) + b.z, rstd, nmr);
          o.w = fmaf(__uint_as_float(v[cc & 1][4 * i + 3]) + b.w, rstd, nmr);
          *reinterpret_cast<float4*>(myX + lane * 128 + ((i ^ (lane & 7)) << 4)) = o;
        }
        __syncwarp();
        const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + col + 4 * c4));
        const float4 be = __ldg(reinterpret_cast<const float4*>(ep.beta + col + 4 * c4));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rl = 4 * k + rsub;
          const int grow = wrow0 + rl;
          const float4 a = *reinterpret_cast<const float4*>(myX + rl * 128 + ((c4 ^ (rl & 7)) << 4));
          if (grow < M) {
            const size_t off = (size_t)grow * N + col + 4 * c4;
            const float4 s = sc[cc & 1][k];
            const float4 o = make_float4(fmaf(a.x, g.x, be.x) + s.x, fmaf(a.y, g.y, be.y) + s.y,
                                         fmaf(a.z, g.z, be.z) + s.z, fmaf(a.w, g.w, be.w) + s.w);
            if (ep.x32) *reinterpret_cast<float4*>(ep.x32 + off) = o;
            if (ep.xb) *reinterpret_cast<uint2*>(ep.xb + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int N, int BK, int STAGES>
static int launch_gemm_ln(const void* A, int lda, const void* W, int ldw, int M, int K, const LnEpiParams& ep,
                          cudaStream_t stream) {
  using Cfg = GemmLnCfg<N, BK, STAGES>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "shared memory budget");
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)BK, GLN_BM};
    int rc = make_tmap_16b(&tmA, A, 2, dims, str, box, Cfg::ROW_BYTES);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)Cfg::NI};
    int rc = make_tmap_16b(&tmB, W, 2, dims, str, box, Cfg::ROW_BYTES);
    if (rc) return rc;
  }
  auto kern = gemm_ln_kernel<N, BK, STAGES>;
  static unsigned long long attr_set = 0;   // per template instantiation, one bit per device
  if (first_use_on_current_device(&attr_set)) {
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int tiles = (M + GLN_BM - 1) / GLN_BM;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GLN_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, K, ep);
  MV_LAUNCH_OK();
  return 0;
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_gemm_ln_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                  const float* bias, const float* gamma, const float* beta, float eps,
                                  const float* shortcut_f32, float* x32, void* xb, cudaStream_t stream) {
  MV_CHECK_ARG(M > 0 && K > 0, "gemm_ln: empty problem M=%d K=%d", M, K);
  MV_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0, "gemm_ln: lda/ldw must be multiples of 8 elements: %d %d", lda, ldw);
  MV_CHECK_ARG(gamma && beta && (x32 || xb), "gemm_ln: gamma, beta and an output are required");
  LnEpiParams ep;
  ep.bias = bias; ep.gamma = gamma; ep.beta = beta; ep.shortcut = shortcut_f32; ep.x32 = x32;
  ep.xb = reinterpret_cast<bf16*>(xb); ep.eps = eps;
  switch (N) {
    case 128: return launch_gemm_ln<128, 64, 5>(A, lda, W, ldw, M, K, ep, stream);
    case 256: return launch_gemm_ln<256, 64, 3>(A, lda, W, ldw, M, K, ep, stream);
    case 512: return launch_gemm_ln<512, 32, 4>(A, lda, W, ldw, M, K, ep, stream);
    default: return mv::fail(-1, "gemm_ln: N = %d not instantiated (128, 256, 512: the row must fit the 512 TMEM columns)", N);
  }
}
