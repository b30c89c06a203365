// Fusion-branch kernels of Multi_DefectModel_new_GCN (GraphModel.py:189-209):
//   * Rs_GCN affinity:  R = theta phi^T / N ;  y = R g   per graph (Rs_GCN.py:57-66; no softmax)
//   * fusion head: l2norm over the node-slot axis + mean over slots + concat(image, graph, text) + BatchNorm1d(1536)
//     (folded) + Linear(1536, num_classes) in ONE kernel (GraphModel.py:200-209).
#include <cstdlib>

#include "affinity_rows.cuh"
#include "common.cuh"
#include "host_util.h"

namespace mv {

// tpg: bf16 [B*n, 3C] rows = (theta | phi | g) of one node slot; y: bf16 [B*n, C].  One CTA (256 threads, 200 active
// in the FMA phases) per graph; n <= 100 slots, C % 64 == 0.
//
// F32 variant (the one the models use): tpg is fp32 and y is written as the bf16x3 split operand [B*n, 3C] =
// (hi | lo | hi), hi = bf16(v), lo = bf16(v - hi), so that the following 1x1-conv GEMM against (W_hi | W_hi | W_lo)
// reproduces the fp32 product to ~2^-16.  Rs_GCN has no softmax and is followed by a BatchNorm whose batch deviation
// is small against the magnitude of W y: a plain bf16 y costs 1.6 % per block there (measured), 9 % after 8 blocks.
constexpr int RS_MAXN = 100;
__device__ __forceinline__ void rs_load8(const bf16* p, float* f) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ void rs_load8(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// 8 fp32 -> (hi, lo) bf16 words
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    l[i] = pack_bf16x2(v[2 * i] - bf16_lo(h[i]), v[2 * i + 1] - bf16_hi(h[i]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
template <typename TIn, bool SPLIT>
__global__ void __launch_bounds__(256)
rs_gcn_affinity_kernel(const TIn* __restrict__ tpg, bf16* __restrict__ y, float* __restrict__ r_out, int n, int C) {
  extern __shared__ float sm[];
  float* R = sm;                                  // [RS_MAXN][RS_MAXN + 1]
  float* bufA = R + RS_MAXN * (RS_MAXN + 1);      // phase 1: theta chunk [RS_MAXN][33]; phase 2: g chunk [RS_MAXN][64]
  float* bufB = bufA + RS_MAXN * 64;              // phase 1: phi chunk [RS_MAXN][33]
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const TIn* base = tpg + (size_t)b * n * 3 * C;

  // ---- phase 1: R = theta phi^T / n ; thread (ti, tj) owns rows ti*5..+5, cols tj*10..+10 ----
  const int ti = tid / 10, tj = tid % 10;
  float acc[5][10];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int c = 0; c < 10; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < C; k0 += 32) {
    // stage theta[:, k0:k0+32] and phi[:, k0:k0+32] (4 uint4 per row each)
    for (int i = tid; i < RS_MAXN * 8; i += 256) {
      const int row = i >> 3, part = i & 7;       // part 0..3 theta, 4..7 phi
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < n) {
        const int col = (part < 4 ? 0 : C) + k0 + (part & 3) * 8;
        rs_load8(base + (size_t)row * 3 * C + col, f);
      }
      float* dstp = (part < 4 ? bufA : bufB) + row * 33 + (part & 3) * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) dstp[q] = f[q];
    }
    __syncthreads();
    if (tid < 200) {
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        float th[5], ph[10];
#pragma unroll
        for (int a = 0; a < 5; ++a) th[a] = bufA[(ti * 5 + a) * 33 + k];
#pragma unroll
        for (int c = 0; c < 10; ++c) ph[c] = bufB[(tj * 10 + c) * 33 + k];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
          for (int c = 0; c < 10; ++c) acc[a][c] += th[a] * ph[c];
      }
    }
    __syncthreads();
  }
  if (tid < 200) {
    const float inv = 1.0f / (float)n;            // R.size(-1) == number of slots
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int c = 0; c < 10; ++c) {
        const float v = acc[a][c] * inv;
        R[(ti * 5 + a) * (RS_MAXN + 1) + tj * 10 + c] = v;
        if (r_out && blockIdx.y == 0 && ti * 5 + a < n && tj * 10 + c < n)
          r_out[((size_t)b * n + ti * 5 + a) * n + tj * 10 + c] = v;
      }
  }
  __syncthreads();

  // ---- phase 2: y = R g ; thread (yi, yj) owns rows yi*4..+4, cols yj*8..+8 of each 64-column chunk ----
  // gridDim.y CTAs share a graph: each recomputes R (phase 1) and produces its own slice of y's columns, so that
  // 64 graphs fill the 148 SMs
  const int yi = tid / 8, yj = tid % 8;
  const int c_per = ((C / 64 + gridDim.y - 1) / gridDim.y) * 64;
  const int c_beg = blockIdx.y * c_per, c_end = min(C, c_beg + c_per);
  for (int c0 = c_beg; c0 < c_end; c0 += 64) {
    for (int i = tid; i < RS_MAXN * 8; i += 256) {
      const int row = i >> 3, part = i & 7;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < n) rs_load8(base + (size_t)row * 3 * C + 2 * C + c0 + part * 8, f);
#pragma unroll
      for (int q = 0; q < 8; ++q) bufA[row * 64 + part * 8 + q] = f[q];
    }
    __syncthreads();
    if (tid < 200) {
      float o[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 8; ++q) o[a][q] = 0.f;
      for (int j = 0; j < n; ++j) {
        float rv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) rv[a] = R[(yi * 4 + a) * (RS_MAXN + 1) + j];
        const float4 g0 = *reinterpret_cast<const float4*>(bufA + j * 64 + yj * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(bufA + j * 64 + yj * 8 + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int q = 0; q < 8; ++q) o[a][q] += rv[a] * gv[q];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int row = yi * 4 + a;
        if (row < n) {
          if (SPLIT) {
            uint4 hi, lo;
            split8(o[a], hi, lo);
            bf16* yr = y + ((size_t)b * n + row) * 3 * C + c0 + yj * 8;
            *reinterpret_cast<uint4*>(yr) = hi;
            *reinterpret_cast<uint4*>(yr + C) = lo;
            *reinterpret_cast<uint4*>(yr + 2 * C) = hi;
          } else {
            uint4 w;
            w.x = pack_bf16x2(o[a][0], o[a][1]); w.y = pack_bf16x2(o[a][2], o[a][3]);
            w.z = pack_bf16x2(o[a][4], o[a][5]); w.w = pack_bf16x2(o[a][6], o[a][7]);
            *reinterpret_cast<uint4*>(y + ((size_t)b * n + row) * C + c0 + yj * 8) = w;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Row-split variant of the fp32 -> split affinity (what mvuld_rs_gcn_affinity_f32 launches): CTA (b, s) owns the 25 rows
// [25 s, 25 s + 25) of R and of y (affinity_rows.cuh), so nothing is computed twice -- the column-split kernel above
// recomputes the whole R in each of its CTAs: 6.4 M FMA per CTA against 2.6 M here (138.7 -> 73.0 us at 64 graphs).
__global__ void __launch_bounds__(256)
rs_gcn_affinity_rows_kernel(const float* __restrict__ tpg, bf16* __restrict__ y, float* __restrict__ r_out, int n,
                            int C) {
  extern __shared__ __align__(16) float smr[];
  float* Rs = smr;                                // [AR_ROWS][AR_MAXN]
  float* U = Rs + AR_ROWS * AR_MAXN;
  const int b = blockIdx.x;
  const int r0 = blockIdx.y * AR_ROWS;
  const float* base = tpg + (size_t)b * n * 3 * C;
  ar_nt<float>(base, 3 * C, r0, base + C, 3 * C, n, C, 1.0f / (float)n, Rs, U);      // R.size(-1) == number of slots
  if (r_out) {
    for (int i = threadIdx.x; i < AR_ROWS * AR_MAXN; i += 256) {
      const int row = r0 + i / AR_MAXN, col = i % AR_MAXN;
      if (row < n && col < n) r_out[((size_t)b * n + row) * n + col] = Rs[i];
    }
  }
  bf16* yb = y + (size_t)b * n * 3 * C;
  ar_sy<float>(Rs, base + 2 * C, 3 * C, r0, n, C, U, [&](int row, int col, const float* o) {
    const uint32_t h0 = pack_bf16x2(o[0], o[1]), h1 = pack_bf16x2(o[2], o[3]);
    const uint32_t l0 = pack_bf16x2(o[0] - bf16_lo(h0), o[1] - bf16_hi(h0));
    const uint32_t l1 = pack_bf16x2(o[2] - bf16_lo(h1), o[3] - bf16_hi(h1));
    bf16* yr = yb + (size_t)row * 3 * C + col;
    *reinterpret_cast<uint2*>(yr) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(yr + C) = make_uint2(l0, l1);
    *reinterpret_cast<uint2*>(yr + 2 * C) = make_uint2(h0, h1);
  });
}

// bf16x3 split of an fp32 matrix x [R, C] (row stride ldx) into out bf16 [R, 3C]:
//   w_side == 0 (activations, the A operand): (hi | lo | hi);   w_side != 0 (weights, the W operand): (hi | hi | lo)
// so that  A3 W3^T = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T  (the lo*lo term, ~2^-18, is dropped).
__global__ void split3_kernel(const float* __restrict__ x, int ldx, bf16* __restrict__ out, int R, int C, int w_side) {
  const int units = C >> 3;
  const long long total = (long long)R * units;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    const long long r = i / units;
    float f[8];
    rs_load8(x + r * ldx + u * 8, f);
    uint4 hi, lo;
    split8(f, hi, lo);
    bf16* o = out + r * 3 * C + u * 8;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + C) = w_side ? hi : lo;
    *reinterpret_cast<uint4*>(o + 2 * C) = w_side ? lo : hi;
  }
}

// One CTA per function.  z: fp32 [B, n, D] (token-major Rs_GCN output);  img / txt: fp32 [B, D] (already ELU(fc(bn))).
// feat = cat(img, l2norm_dim1(z).mean(1), txt);  logits = Wf feat + bf with BatchNorm1d(3D) folded into (Wf, bf).
// mode 1 (new_model.py:317, Multi_DefectModel_noFunc): feat = cat(img, graph) [2D];
// mode 2 (new_model.py:196, Multi_DefectModel_noGlobalImage): feat = txt * graph [D].
__global__ void __launch_bounds__(512)
fusion_head_kernel(const float* __restrict__ z, const float* __restrict__ img, const float* __restrict__ txt,
                   const float* __restrict__ wf, const float* __restrict__ bf, float* __restrict__ logits,
                   float* __restrict__ feat_out, int n, int D, int num_classes, int mode) {
  extern __shared__ float feat[];                 // [3D]
  const int b = blockIdx.x;
  const int F = mode == 0 ? 3 * D : (mode == 1 ? 2 * D : D);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f, sq = 0.f;
    const float* zp = z + (size_t)b * n * D + c;
    for (int r = 0; r < n; ++r) {
      const float v = __ldg(zp + (size_t)r * D);
      s += v;
      sq += v * v;
    }
    const float gf = (s / sqrtf(sq)) / (float)n;  // l2norm has no eps (GraphModel.py:74-79)
    if (mode == 2) {
      feat[c] = __ldg(txt + (size_t)b * D + c) * gf;
    } else {
      feat[c] = __ldg(img + (size_t)b * D + c);
      feat[D + c] = gf;
      if (mode == 0) feat[2 * D + c] = __ldg(txt + (size_t)b * D + c);
    }
  }
  __syncthreads();
  if (feat_out)
    for (int c = threadIdx.x; c < F; c += blockDim.x) feat_out[(size_t)b * F + c] = feat[c];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < num_classes; k += blockDim.x >> 5) {
    float acc = 0.f;
    for (int c = lane; c < F; c += 32) acc += __ldg(wf + (size_t)k * F + c) * feat[c];
    acc = warp_sum(acc);
    if (lane == 0) logits[(size_t)b * num_classes + k] = acc + __ldg(bf + k);
  }
}

// Small-N fp32 linear (classification heads): out[m, n] = act(<x[m,:], w[n,:]> + b[n]); one warp per row.
// act: 0 none, 3 sigmoid.  swin_transformer_v2.py:642 (head), baselines/models/reveal/ggnn/model.py:29-30.
__global__ void __launch_bounds__(256)
linear_small_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ out, float* __restrict__ out_act, int M, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  for (int n = 0; n < N; ++n) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += __ldg(x + (size_t)row * K + k) * __ldg(w + (size_t)n * K + k);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += b ? __ldg(b + n) : 0.f;
      out[(size_t)row * N + n] = acc;
      if (out_act) out_act[(size_t)row * N + n] = 1.f / (1.f + expf(-acc));
    }
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_linear_small(const float* x, const float* w, const float* b, float* out, float* out_sigmoid,
                                  int M, int N, int K, cudaStream_t stream) {
  MV_CHECK_ARG(N >= 1 && N <= 64, "linear_small: N must be in [1, 64]");
  if (M <= 0) return 0;
  linear_small_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, w, b, out, out_sigmoid, M, N, K);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_rs_gcn_affinity_f32(const float* tpg, void* y3, float* r_out, int B, int n, int C,
                                         cudaStream_t stream) {
  MV_CHECK_ARG(n >= 1 && n <= RS_MAXN, "rs_gcn_affinity_f32: n must be in [1, %d]", RS_MAXN);
  MV_CHECK_ARG(C % 64 == 0, "rs_gcn_affinity_f32: C %% 64");
  if (B <= 0) return 0;
  if (C % AR_GCH == 0) {
    auto kern = rs_gcn_affinity_rows_kernel;
    MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM_BYTES));
    kern<<<dim3(B, (n + AR_ROWS - 1) / AR_ROWS), 256, AR_SMEM_BYTES, stream>>>(tpg, reinterpret_cast<bf16*>(y3), r_out, n, C);
    MV_LAUNCH_OK();
    return 0;
  }
  const int smem = (RS_MAXN * (RS_MAXN + 1) + RS_MAXN * 64 + RS_MAXN * 33) * sizeof(float);
  auto kern = rs_gcn_affinity_kernel<float, true>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<dim3(B, B >= 148 ? 2 : 4), 256, smem, stream>>>(tpg, reinterpret_cast<bf16*>(y3), r_out, n, C);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_split3_bf16(const float* x, int ldx, void* out, int R, int C, int w_side, cudaStream_t stream) {
  MV_CHECK_ARG(C % 8 == 0 && ldx % 4 == 0 && ldx >= C, "split3: C %% 8, ldx %% 4, ldx >= C (C=%d ldx=%d)", C, ldx);
  if (R <= 0) return 0;
  long long blocks = ((long long)R * (C / 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split3_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, ldx, reinterpret_cast<bf16*>(out), R, C, w_side);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_fusion_head_mode(const float* z, const float* img, const float* txt, const float* wf,
                                      const float* bf, float* logits, float* feat_out, int B, int n, int D,
                                      int num_classes, int mode, cudaStream_t stream) {
  MV_CHECK_ARG(mode >= 0 && mode <= 2, "fusion_head: mode must be 0 (img|graph|txt), 1 (img|graph) or 2 (txt*graph)");
  if (B <= 0) return 0;
  fusion_head_kernel<<<B, 512, 3 * D * sizeof(float), stream>>>(z, img, txt, wf, bf, logits, feat_out, n, D,
                                                               num_classes, mode);
  MV_LAUNCH_OK();
  return 0;
}
