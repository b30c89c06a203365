// Single-tile UMMA probe: one TMA box for A, one for B, `nk` tcgen05.mma K-steps with caller-chosen descriptor
// parameters, accumulator dumped as fp32 [128, N].  Used by tests/ to pin every shared-memory layout the GEMM
// and attention kernels rely on (K-major SW128 / SW64, MN-major SW64 / SW128, fp16 and bf16 operands)
// independently of the big kernels.  Test fixture: built into its own library (libmvuld_probe.so), not into the product
// libmvuld_b200.so.
#include "../common.cuh"
#include "../host_util.h"

namespace mv {

struct ProbeParams {
  int a_bytes, b_bytes;        // TMA box bytes (expect_tx)
  int nk;                      // MMA K-steps (K = 16 each)
  int a_step, b_step;          // descriptor start-address advance per K-step (bytes)
  int a_lbo, a_sbo, a_layout;  // descriptor fields for A
  int b_lbo, b_sbo, b_layout;  // descriptor fields for B
  uint32_t idesc;
  int N;
  float* out;                  // [128, N]
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], p.a_bytes + p.b_bytes);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0, 90);
    tc_fence_after();
    for (int k = 0; k < p.nk; ++k) {
      const uint64_t ad = make_smem_desc(smem_u32(sA) + k * p.a_step, p.a_lbo, p.a_sbo, p.a_layout);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + k * p.b_step, p.b_lbo, p.b_sbo, p.b_layout);
      umma_ss(tmem, ad, bd, p.idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, 91);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int q = 0; q < 16; ++q)
      if (c0 + q < p.N) p.out[(size_t)row * p.N + c0 + q] = __uint_as_float(r[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace mv

using namespace mv;

// A: 16-bit [a_rows, a_inner] row-major, B: 16-bit [b_rows, b_inner] row-major; each loaded as ONE box with the given
// swizzle.  fmt: 0 = fp16, 1 = bf16 (both operands: a bf16 A with an fp16 B, tried here on a B200, raises an illegal
// instruction -- kind::f16 wants one format for A and B).  a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
extern "C" int mvuld_probe_umma(const void* A, int a_inner, int a_rows, int a_swizzle, const void* B, int b_inner,
                                int b_rows, int b_swizzle, int N, int nk, int a_step, int b_step, int a_lbo, int a_sbo,
                                int a_layout, int b_lbo, int b_sbo, int b_layout, int a_mn, int b_mn, int fmt,
                                float* out, cudaStream_t stream) {
  MV_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256, "probe: N");
  MV_CHECK_ARG(a_inner * a_rows * 2 <= 65536 && b_inner * b_rows * 2 <= 65536, "probe: box too large");
  CUtensorMap tmA, tmB;
  uint64_t da[2] = {(uint64_t)a_inner, (uint64_t)a_rows};
  uint64_t sa[1] = {(uint64_t)a_inner * 2};
  uint32_t ba[2] = {(uint32_t)a_inner, (uint32_t)a_rows};
  int rc = make_tmap_16b(&tmA, A, 2, da, sa, ba, a_swizzle);
  if (rc) return rc;
  uint64_t db[2] = {(uint64_t)b_inner, (uint64_t)b_rows};
  uint64_t sb[1] = {(uint64_t)b_inner * 2};
  uint32_t bb[2] = {(uint32_t)b_inner, (uint32_t)b_rows};
  rc = make_tmap_16b(&tmB, B, 2, db, sb, bb, b_swizzle);
  if (rc) return rc;
  ProbeParams p;
  p.a_bytes = a_inner * a_rows * 2;
  p.b_bytes = b_inner * b_rows * 2;
  p.nk = nk; p.a_step = a_step; p.b_step = b_step;
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_layout = a_layout;
  p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_layout = b_layout;
  uint32_t id = make_idesc_bf16(128, N, a_mn, b_mn);
  if (fmt == 0) id &= ~((1u << 7) | (1u << 10));
  p.idesc = id;
  p.N = N;
  p.out = out;
  const int smem = 131072 + 256 + 1024;
  MV_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, stream>>>(tmA, tmB, p);
  MV_LAUNCH_OK();
  return 0;
}
