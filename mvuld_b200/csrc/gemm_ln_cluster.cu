// tcgen05 GEMM with a full-row LayerNorm (+ residual) epilogue for rows WIDER than one CTA's TMEM can double-buffer:
//
//     x[M,N] = shortcut + LayerNorm(A[M,K] W[N,K]^T + bias) * gamma + beta,     N = CL * 256,  CL in {2, 3, 4}
//
// (SwinV2 stage 2 / 3 res-post-norm, swin_transformer_v2.py:301,304, C = 512 / 1024; PatchMerging norm(reduction(x)),
// :361-362, 2C = 1024.)
//
// gemm_ln.cu gives a CTA the whole row: at N = 512 that is all 512 TMEM columns, a single accumulator, and the
// epilogue (two TMEM passes + 10 bytes of HBM traffic per element) serialises with the mainloop -- measured slower than
// GEMM + a separate LayerNorm pass.  Here a thread-block CLUSTER of CL CTAs owns a 128-row tile: CTA `rank` computes
// columns [256 rank, 256 rank + 256) into one of TWO 256-column TMEM accumulators, so the epilogue of tile i overlaps the
// mainloop of tile i + 1 as in the plain GEMM, and the row statistics are exchanged through distributed shared memory:
// every epilogue thread stores its (sum, sum of squares) over 128 columns into the statistics slab of EVERY CTA of the
// cluster (st.shared::cluster) and arrives (release.cluster) on that CTA's mbarrier; a CTA normalises once its own
// barrier has collected CL x 256 arrivals.  One launch and no bf16 round trip of the GEMM output (12 B / element of
// HBM traffic saved at C = 512 over 36 LayerNorms per SwinV2 forward).
//
//   warp 0      TMA producer : A box [128 x 64], W box [256 x 64] (this CTA's column block) -> 3-stage ring, SW128
//   warp 1      MMA issuer   : tcgen05.mma M128 x N256 x K16
//   warps 2..9  epilogue     : thread = (row, 128-column half); pass 1 statistics, DSMEM exchange, pass 2 normalise in
//                              16-column chunks transposed through shared memory for coalesced global stores
//   warp 10     residual TMA : the fp32 residual is the only long-latency operand of the epilogue; register prefetch
//                              (4 x 16 B per thread, one chunk ahead) kept 16 KB in flight per SM = 2.4 TB/s over the
//                              chip and made the first version of this kernel slower than GEMM + LayerNorm.  One
//                              thread streams [128 rows x 16 columns] fp32 boxes of the residual through two 3-slot
//                              rings (one per column half, 48 KB in flight per SM), running ahead across tiles.
#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int GLC_BM = 128;
constexpr int GLC_NC = 256;          // columns per CTA
constexpr int GLC_BK = 64;
constexpr int GLC_STAGES = 3;
constexpr int GLC_THREADS = 352;
constexpr int GLC_EPI_THREADS = 256;
constexpr int GLC_SC_SLOTS = 3;
constexpr int GLC_SC_BYTES = GLC_BM * 64;      // one residual box: 128 rows x 16 fp32

template <int CL>
struct GlcCfg {
  static constexpr int A_BYTES = GLC_BM * GLC_BK * 2;                         // 16 KB
  static constexpr int B_BYTES = GLC_NC * GLC_BK * 2;                         // 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STATS_BYTES = 2 /*parity*/ * CL * 2 /*half*/ * GLC_BM * 8;
  static constexpr int XPOSE_BYTES = 8 /*epilogue warps*/ * 32 * 64;          // one [32 rows x 16 fp32] chunk per warp
  static constexpr int SC_BYTES = 2 * GLC_SC_SLOTS * GLC_SC_BYTES;            // residual rings
  static constexpr int SMEM_BYTES = GLC_STAGES * STAGE_BYTES + STATS_BYTES + XPOSE_BYTES + SC_BYTES + 256 + 1024;
};

struct GlcParams {
  const float* bias;       // [N] or null
  const float* gamma;      // [N]
  const float* beta;       // [N]
  const float* shortcut;   // fp32 [M, N] or null (may alias x32)
  float* x32;              // fp32 [M, N] or null
  bf16* xb;                // bf16 [M, N] or null
  float eps;
};

// Tried and not kept: the A tile multicast over the cluster (each CTA loads 128 / CL of its rows into every CTA's ring
// slot, the empty barriers count CL multicast commits): correct, but 128.1 vs 126.8 us (N = 512, K = 2 048) and 112 vs
// 107 us (N = 1 024, K = 4 096) -- the CTAs of a cluster then release ring slots in lockstep, which costs what the
// 17 % of L2 -> SM traffic saves.
template <int CL>
__global__ void __launch_bounds__(GLC_THREADS, 1)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmS, int M, int K, GlcParams ep) {
  using Cfg = GlcCfg<CL>;
  constexpr int N = CL * GLC_NC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + GLC_STAGES * Cfg::A_BYTES;
  float2* sStats = reinterpret_cast<float2*>(smem + GLC_STAGES * Cfg::STAGE_BYTES);   // [parity][src rank][half][row]
  uint8_t* sX = smem + GLC_STAGES * Cfg::STAGE_BYTES + Cfg::STATS_BYTES;              // [warp][32 rows][4 x 16 B, swizzled]
  uint8_t* sSc = sX + Cfg::XPOSE_BYTES;                                               // [half][slot][128 rows][64 B]
  uint64_t* full = reinterpret_cast<uint64_t*>(sSc + Cfg::SC_BYTES);
  uint64_t* empty = full + GLC_STAGES;
  uint64_t* tfull = empty + GLC_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* stat = tempty + 2;
  uint64_t* sc_full = stat + 2;                                                       // [half][slot]
  uint64_t* sc_empty = sc_full + 2 * GLC_SC_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_empty + 2 * GLC_SC_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;
  const int num_tiles = (M + GLC_BM - 1) / GLC_BM;
  const int num_kb = (K + GLC_BK - 1) / GLC_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (ep.shortcut) prefetch_tmap(&tmS);
    for (int s = 0; s < GLC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], GLC_EPI_THREADS);
      mbar_init(&stat[a], CL * GLC_EPI_THREADS);
    }
    for (int i = 0; i < 2 * GLC_SC_SLOTS; ++i) {
      mbar_init(&sc_full[i], 1);
      mbar_init(&sc_empty[i], GLC_EPI_THREADS / 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // every CTA's barriers are initialised before a peer may arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * GLC_BK, tile * GLC_BM);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * GLC_BK, rank * GLC_NC);
          if (++stage == GLC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(GLC_BM, GLC_NC, 0, 0);
    const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_smem_desc(smem_u32(sA), 16, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(sB), 16, 1024, 2);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int acc = local & 1;
      mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * GLC_NC;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase, 3);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (Cfg::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < GLC_BK / 16; ++k)
            umma_ss(d_tmem, ((uint64_t)desc_hi << 32) | (a_lo + k * 2), ((uint64_t)desc_hi << 32) | (b_lo + k * 2), idesc,
                    (kb | k) != 0);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == GLC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else if (warp == 10) {
    // residual boxes in the order the epilogue consumes them: (tile, chunk, half); the fp32 matrix is addressed as
    // 16-bit elements (TMA moves bytes), so a 16-column box is 32 elements wide
    if (lane == 0 && ep.shortcut) {
      uint32_t n = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        for (int cc = 0; cc < GLC_NC / 2 / 16; ++cc, ++n) {
          const uint32_t slot = n % GLC_SC_SLOTS, par = (n / GLC_SC_SLOTS) & 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint64_t* fb = &sc_full[h * GLC_SC_SLOTS + slot];
            mbar_wait(&sc_empty[h * GLC_SC_SLOTS + slot], par ^ 1, 7);
            mbar_arrive_expect_tx(fb, GLC_SC_BYTES);
            tma_load_2d(sSc + (h * GLC_SC_SLOTS + slot) * GLC_SC_BYTES, &tmS, fb,
                        (rank * GLC_NC + h * (GLC_NC / 2) + cc * 16) * 2, tile * GLC_BM);
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = GLC_NC / 2;                // 128 columns per epilogue thread
    const int r = quarter * 32 + lane;            // row within the tile
    const float invN = 1.0f / (float)N;
    const int colbase = rank * GLC_NC + half * HC;   // first global column of this thread
    // transposed lane mapping of the global accesses: lane handles float4 column c4 = lane % 4 of rows 8 k + lane / 4
    const int c4 = lane & 3, rsub = lane >> 2;
    uint8_t* myX = sX + (warp - 2) * 2048;
    const uint32_t stats_local = smem_u32(sStats);
    const uint32_t stat_local = smem_u32(stat);
    int local = 0;
    uint32_t nsc = 0;                             // residual boxes consumed so far (ring position)
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int acc = local & 1;
      const uint32_t par = (local >> 1) & 1;
      const int wrow0 = tile * GLC_BM + quarter * 32;               // first global row of this warp
      mbar_wait(&tfull[acc], par, 4);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * GLC_NC + half * HC;
      __syncwarp();                                               // converged before the .sync.aligned TMEM loads
      // pass 1: sum and sum of squares of (acc + bias) over this thread's 128 columns
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
      // exchange: my partial goes into slot [tile parity][my rank][half][row] of every CTA of the cluster
      const uint32_t slot = (uint32_t)((((local & 1) * CL + rank) * 2 + half) * GLC_BM + r) * 8u;
#pragma unroll
      for (int t = 0; t < CL; ++t) {
        st_cluster_f2(mapa_u32(stats_local + slot, (uint32_t)t), s1, s2);
        mbar_arrive_cluster(mapa_u32(stat_local + (uint32_t)(local & 1) * 8u, (uint32_t)t));
      }
      mbar_wait_cluster(&stat[local & 1], par, 6);
      __syncwarp();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int t = 0; t < CL * 2; ++t) {                              // fixed order: identical statistics in every CTA
        const float2 p = sStats[((local & 1) * CL * 2 + t) * GLC_BM + r];
        t1 += p.x;
        t2 += p.y;
      }
      const float mean = t1 * invN;
      const float var = fmaxf(t2 * invN - mean * mean, 0.f);
      const float rstd = rsqrtf(var + ep.eps);
      const float nmr = -mean * rstd;
      // pass 2: normalise in the row-owner layout, transpose the 32 x 16 chunk through shared memory (16-byte units
      // XOR-swizzled: conflict free both ways), then affine + residual + coalesced stores
      uint32_t v[2][16];
      tmem_ld16p(t0, v[0]);
#pragma unroll
      for (int cc = 0; cc < HC / 16; ++cc, ++nsc) {
        const int c = cc * 16;
        const int col = colbase + c;
        tmem_ld_wait();
        if (cc + 1 < HC / 16) tmem_ld16p(t0 + c + 16, v[(cc + 1) & 1]);
        __syncwarp();                                               // previous chunk's reads of myX are done
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4 * i));
          float4 o;
          o.x = fmaf(__uint_as_float(v[cc & 1][4 * i]) + b.x, rstd, nmr);
          o.y = fmaf(__uint_as_float(v[cc & 1][4 * i + 1]) + b.y, rstd, nmr);
          o.z = fmaf(__uint_as_float(v[cc & 1][4 * i + 2]) + b.z, rstd, nmr);
          o.w = fmaf(__uint_as_float(v[cc & 1][4 * i + 3]) + b.w, rstd, nmr);
          *reinterpret_cast<float4*>(myX + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = o;
        }
        __syncwarp();
        // this chunk's residual box (rows of the whole tile; this warp reads its own 32)
        float4 sc[4];
        if (ep.shortcut) {
          const uint32_t slot = nsc % GLC_SC_SLOTS;
          mbar_wait(&sc_full[half * GLC_SC_SLOTS + slot], (nsc / GLC_SC_SLOTS) & 1, 8);
          __syncwarp();                                           // lanes leave the spin loop at different times
          const uint8_t* box = sSc + (half * GLC_SC_SLOTS + slot) * GLC_SC_BYTES + quarter * 32 * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) sc[k] = *reinterpret_cast<const float4*>(box + (8 * k + rsub) * 64 + c4 * 16);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) sc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + col + 4 * c4));
        const float4 be = __ldg(reinterpret_cast<const float4*>(ep.beta + col + 4 * c4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rl = 8 * k + rsub;
          const int grow = wrow0 + rl;
          const float4 a = *reinterpret_cast<const float4*>(myX + rl * 64 + ((c4 ^ ((rl >> 1) & 3)) << 4));
          const float4 s = sc[k];
          const float4 o = make_float4(fmaf(a.x, g.x, be.x) + s.x, fmaf(a.y, g.y, be.y) + s.y, fmaf(a.z, g.z, be.z) + s.z,
                                       fmaf(a.w, g.w, be.w) + s.w);
          if (grow < M) {
            const size_t off = (size_t)grow * N + col + 4 * c4;
            if (ep.x32) *reinterpret_cast<float4*>(ep.x32 + off) = o;
            if (ep.xb) *reinterpret_cast<uint2*>(ep.xb + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          }
        }
        // The ring slot is released only HERE, after the values read from it have been consumed by the stores above.
        // Arriving right after issuing the ld.shared let the TMA refill the slot before the loads had read it (an
        // mbarrier arrive does not wait for the thread's outstanding shared-memory loads): under memory pressure from
        // another stream a few rows then received the residual of a later chunk.
        if (ep.shortcut) mbar_arrive(&sc_empty[half * GLC_SC_SLOTS + nsc % GLC_SC_SLOTS]);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // no CTA leaves while a peer may still store into its statistics slab
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int CL>
static int launch_gemm_ln_cluster(const void* A, int lda, const void* W, int ldw, int M, int K, const GlcParams& ep,
                                  cudaStream_t stream) {
  using Cfg = GlcCfg<CL>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "shared memory budget");
  constexpr int N = CL * GLC_NC;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)GLC_BK, GLC_BM};
    int rc = make_tmap_16b(&tmA, A, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)GLC_BK, (uint32_t)GLC_NC};
    int rc = make_tmap_16b(&tmB, W, 2, dims, str, box, 128);
    if (rc) return rc;
  }
  CUtensorMap tmS = tmA;               // placeholder when there is no residual (never dereferenced)
  if (ep.shortcut) {
    uint64_t dims[2] = {(uint64_t)N * 2, (uint64_t)M};             // fp32 [M, N] addressed as 16-bit [M, 2N]
    uint64_t str[1] = {(uint64_t)N * 4};
    uint32_t box[2] = {32, GLC_BM};
    int rc = make_tmap_16b(&tmS, ep.shortcut, 2, dims, str, box, 0);
    if (rc) return rc;
  }
  auto kern = gemm_ln_cluster_kernel<CL>;
  static int max_clusters_dev[64] = {0};   // per template instantiation and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int& max_clusters = max_clusters_dev[dev];
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(GLC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters == 0) {
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    cfg.gridDim = dim3(num_sms() / CL * CL);
    int n = 0;
    MV_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    MV_CHECK_ARG(n > 0, "gemm_ln_cluster: no cluster of %d CTAs fits the device", CL);
    max_clusters = n;
  }
  const int tiles = (M + GLC_BM - 1) / GLC_BM;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cfg.gridDim = dim3(clusters * CL);
  MV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmS, M, K, ep));
  return 0;
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_gemm_ln_wide_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                       const float* bias, const float* gamma, const float* beta, float eps,
                                       const float* shortcut_f32, float* x32, void* xb, cudaStream_t stream) {
  MV_CHECK_ARG(M > 0 && K > 0, "gemm_ln_wide: empty problem M=%d K=%d", M, K);
  MV_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0, "gemm_ln_wide: lda/ldw must be multiples of 8 elements: %d %d", lda, ldw);
  MV_CHECK_ARG(gamma && beta && (x32 || xb), "gemm_ln_wide: gamma, beta and an output are required");
  GlcParams ep;
  ep.bias = bias; ep.gamma = gamma; ep.beta = beta; ep.shortcut = shortcut_f32; ep.x32 = x32;
  ep.xb = reinterpret_cast<bf16*>(xb); ep.eps = eps;
  switch (N) {
    case 512: return launch_gemm_ln_cluster<2>(A, lda, W, ldw, M, K, ep, stream);
    case 768: return launch_gemm_ln_cluster<3>(A, lda, W, ldw, M, K, ep, stream);
    case 1024: return launch_gemm_ln_cluster<4>(A, lda, W, ldw, M, K, ep, stream);
    default: return mv::fail(-1, "gemm_ln_wide: N = %d not instantiated (512, 768, 1024)", N);
  }
}
