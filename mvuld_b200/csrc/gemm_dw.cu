// Weight-gradient product dW[N_out, K_in] = dY^T X for sm_100a: dY bf16 [M, N_out], X bf16 [M, K_in], both ROW-major
// as the forward / backward passes leave them, reduced over the M token rows (autograd's grad of F.linear's weight,
// swin_transformer_v2.py / GraphModel.py under mvuld/main.py:251-300, main_bigvul.py:294-342).
//
// The general GEMM (gemm.cu) wants K-major operands, so this product used to cost two transposes of [M, *] activations
// per weight (16 ms of a 115 ms SwinV2 training step) and then ran on ONE CTA when dW is a single 128 x 128 tile with
// K = 401 408.  Here both operands are consumed MN-major straight from their row-major tensors (TMA boxes of 64 rows x
// 64 columns land as 128-byte-swizzle MN-major atoms; tcgen05 takes the transposes through the descriptor major bits),
// and the M rows are split over CTAs: partial tiles (fp32) + a fixed-order reduction -- no atomics, bit-reproducible.
// Warp roles: 0 TMA producer, 1 MMA issuer (uniform control flow, elected lane), 2 TMEM owner, 4-7 epilogue.
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int DW_THREADS = 256;
constexpr int DW_BM = 128;       // rows of dW per tile  (columns of dY)
constexpr int DW_BK = 64;        // token rows per pipeline stage
constexpr int DW_STAGES = 4;

template <int BN>
struct DwCfg {
  static constexpr int A_BYTES = DW_BM * DW_BK * 2;       // two [64 rows x 128 B] atoms
  static constexpr int B_BYTES = BN * DW_BK * 2;
  static constexpr int SMEM = DW_STAGES * (A_BYTES + B_BYTES) + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(DW_THREADS, 1)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
               int ld_out, long long split_stride, int n_out, int k_in, int M, int rows_per_split, int tiles_n) {
  using Cf = DwCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + DW_STAGES * Cf::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_STAGES * (Cf::A_BYTES + Cf::B_BYTES));
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + DW_STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * DW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DW_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
  const int row0 = split * rows_per_split;
  const int row1 = min(M, row0 + rows_per_split);
  const int num_kb = (row1 - row0 + DW_BK - 1) / DW_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < DW_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);
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
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1, 1);
        mbar_arrive_expect_tx(&full[stage], Cf::A_BYTES + Cf::B_BYTES);
        const int r = row0 + kb * DW_BK;
        // rows past row1 inside the last box belong to the next split (or lie past M: zero filled): the MMA below runs
        // whole 16-row steps, so a split boundary must be a multiple of 64 rows (the host guarantees it)
#pragma unroll
        for (int a = 0; a < DW_BM / 64; ++a)
          tma_load_2d(sA + stage * Cf::A_BYTES + a * (DW_BK * 128), &tmA, &full[stage], m_blk * DW_BM + a * 64, r);
#pragma unroll
        for (int a = 0; a < BN / 64; ++a)
          tma_load_2d(sB + stage * Cf::B_BYTES + a * (DW_BK * 128), &tmB, &full[stage], n_blk * BN + a * 64, r);
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(DW_BM, BN, 1, 1);            // both operands MN-major
    // MN-major, 128-byte swizzle: 64-element MN atoms DW_BK * 128 bytes apart (LBO), 8-row K groups 1 KB apart (SBO)
    const uint32_t hi = (uint32_t)(make_smem_desc(0, 0, 1024, 2) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_smem_desc(smem_u32(sA), DW_BK * 128, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(sB), DW_BK * 128, 1024, 2);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[stage], phase, 2);
      tc_fence_after();
      if (leader) {
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cf::A_BYTES >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)stage * (Cf::B_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < DW_BK / 16; ++k)       // 16 token rows per MMA: 16 * 128 bytes further into every atom
          umma_ss(tmem_base, ((uint64_t)hi << 32) | (a_lo + k * (16 * 128 >> 4)),
                  ((uint64_t)hi << 32) | (b_lo + k * (16 * 128 >> 4)), idesc, (kb | k) != 0);
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // row of the tile == TMEM lane
    const int orow = m_blk * DW_BM + r;
    float* dst = out + (long long)split * split_stride + (long long)orow * ld_out + n_blk * BN;
    if (num_kb > 0) {
      mbar_wait(acc_full, 0, 3);
      tc_fence_after();
    }
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (orow < n_out) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (n_blk * BN + c0 + i < k_in)
            *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

// out[i] = sum_s partial[s][i] in split order (i over the [n_out, ld] rows of dW)
__global__ void dw_reduce_kernel(const float* __restrict__ partial, long long split_stride, int splits,
                                 float* __restrict__ out, int ld_out, int n_out, int k_in) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // float4 index over [n_out, k_in / 4]
  const int k4 = k_in >> 2;
  if (i >= (long long)n_out * k4) return;
  const int r = (int)(i / k4), c = (int)(i - (long long)r * k4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(partial + s * split_stride + (long long)r * k_in + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + (long long)r * ld_out + c) = acc;
}

}  // namespace mv

using namespace mv;

// How the host splits the M rows.  One CTA per SM (192 KB of shared memory), so a launch runs in ceil(CTAs / SMs) waves;
// a CTA's time is its k blocks (bound by the L2 -> SM path: ~0.6 us per 64-row block of a 128 x 256 tile, 0.4 us at
// 128 x 128) plus ~3 us of prologue / epilogue, and every split adds one partial tile set to write and
// read back.  Take the split count that minimises that estimate -- "two CTAs per SM's worth of splits" put the SwinV2
// stage-2 products at 2.16 waves = three waves (93 us for [2048, 512] x 25 088 rows; one wave of four splits: 65 us).
// Split boundaries on multiples of 64 rows, at least 8 pipeline stages of work per split.
static void dw_plan(int M, int n_out, int k_in, int& bn, int& tiles, int& splits, int& rows_per_split) {
  bn = k_in > 128 ? 256 : 128;
  const int tiles_m = (n_out + DW_BM - 1) / DW_BM, tiles_n = (k_in + bn - 1) / bn;
  tiles = tiles_m * tiles_n;
  const int kb_total = (M + DW_BK - 1) / DW_BK;
  const int sms = num_sms();
  const int max_splits = std::max(1, std::min(kb_total / 8, 64));
  const double t_kb = bn == 256 ? 0.6 : 0.4, t_cta = 3.0;
  const double t_split = (double)n_out * k_in * 8.0 / 5.0e6;         // us: write + read of one fp32 partial set at 5 TB/s
  double best = 1e30;
  int best_s = 1;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const int kb_per = (kb_total + sp - 1) / sp;
    const int real = (kb_total + kb_per - 1) / kb_per;               // splits that actually get rows
    const long long waves = ((long long)tiles * real + sms - 1) / sms;
    const double cost = waves * (kb_per * t_kb + t_cta) + (real > 1 ? real * t_split : 0.0);
    if (cost < best) { best = cost; best_s = real; }
  }
  splits = best_s;
  const int kb_per = (kb_total + splits - 1) / splits;
  rows_per_split = kb_per * DW_BK;
  splits = (kb_total + kb_per - 1) / kb_per;
}
// floats of the partials workspace mvuld_gemm_dw needs (0: a single split writes dW directly)
extern "C" long long mvuld_gemm_dw_workspace(int M, int n_out, int k_in) {
  int bn, tiles, splits, rps;
  dw_plan(M, n_out, k_in, bn, tiles, splits, rps);
  return splits > 1 ? (long long)splits * n_out * k_in : 0;
}
extern "C" int mvuld_gemm_dw(const void* dY, int ld_dy, const void* X, int ld_x, float* dW, int ld_dw, float* partials,
                             int M, int n_out, int k_in, cudaStream_t stream) {
  MV_CHECK_ARG(M > 0 && n_out > 0 && k_in > 0, "gemm_dw: empty problem M=%d n_out=%d k_in=%d", M, n_out, k_in);
  MV_CHECK_ARG(ld_dy % 8 == 0 && ld_x % 8 == 0, "gemm_dw: operand row strides must be multiples of 8 elements");
  MV_CHECK_ARG(k_in % 4 == 0 && ld_dw % 4 == 0, "gemm_dw: k_in and ld_dw must be multiples of 4");
  int bn, tiles, splits, rps;
  dw_plan(M, n_out, k_in, bn, tiles, splits, rps);
  MV_CHECK_ARG(splits == 1 || partials != nullptr, "gemm_dw: the partials workspace (mvuld_gemm_dw_workspace floats) is null");
  CUtensorMap tmA, tmB;
  uint64_t da[2] = {(uint64_t)n_out, (uint64_t)M}, sa[1] = {(uint64_t)ld_dy * 2};
  uint64_t db[2] = {(uint64_t)k_in, (uint64_t)M}, sb[1] = {(uint64_t)ld_x * 2};
  uint32_t box[2] = {64, DW_BK};
  int rc;
  if ((rc = make_tmap_16b(&tmA, dY, 2, da, sa, box, 128))) return rc;
  if ((rc = make_tmap_16b(&tmB, X, 2, db, sb, box, 128))) return rc;
  const int tiles_n = (k_in + bn - 1) / bn;
  float* out = splits > 1 ? partials : dW;
  const int ld_out = splits > 1 ? k_in : ld_dw;
  const long long sstride = (long long)n_out * k_in;
  dim3 grid(tiles, splits);
  if (bn == 256) {
    auto kern = gemm_dw_kernel<256>;
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<256>::SMEM));
    kern<<<grid, DW_THREADS, DwCfg<256>::SMEM, stream>>>(tmA, tmB, out, ld_out, sstride, n_out, k_in, M, rps, tiles_n);
  } else {
    auto kern = gemm_dw_kernel<128>;
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DwCfg<128>::SMEM));
    kern<<<grid, DW_THREADS, DwCfg<128>::SMEM, stream>>>(tmA, tmB, out, ld_out, sstride, n_out, k_in, M, rps, tiles_n);
  }
  MV_LAUNCH_OK();
  if (splits > 1) {
    const long long n4 = (long long)n_out * (k_in / 4);
    dw_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(partials, sstride, splits, dW, ld_dw, n_out, k_in);
    MV_LAUNCH_OK();
  }
  return 0;
}
