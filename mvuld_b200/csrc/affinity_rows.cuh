// Row-split building blocks of the Rs_GCN affinity (Rs_GCN.py:57-66) and of its backward pass: a CTA owns AR_ROWS = 25
// consecutive rows of one graph's n <= 100 slot matrices, so the CTAs of a graph share no work (the column-split
// kernels recomputed the n x n matrices in every CTA).  256 threads; both functions end with a __syncthreads().
//   ar_nt   Rs[i][j] = scale * sum_c X[r0 + i][c] * Y[j][c]        (25 x n block of an "NT" product over C features)
//           200 threads = 4 k-groups x (5 x 10) thread grid; a thread accumulates 5 x 10 outputs over its quarter of
//           every 32-wide k chunk, the 4 partial tiles are combined through shared memory in a fixed order
//   ar_sy   out[r0 + i][c] = sum_j Rs[i][j] * Y[j][c]              (25 x C block), 128 columns of Y staged per pass,
//           160 threads = 5 row groups x 32 float4 columns
#pragma once
#include "common.cuh"

namespace mv {

constexpr int AR_MAXN = 100;                      // slots per graph (max_node)
constexpr int AR_ROWS = 25;                       // rows per CTA
constexpr int AR_GCH = 128;                       // columns of Y staged per pass of ar_sy
constexpr int AR_SMEM_BYTES = (AR_ROWS * AR_MAXN + AR_MAXN * AR_GCH) * 4;     // Rs + the larger of the phase buffers

__device__ __forceinline__ float4 ar_load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ar_load4(const bf16* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}

// Rs: [AR_ROWS][AR_MAXN]; U: scratch of AR_MAXN * AR_GCH floats (16-byte aligned)
template <typename T>
__device__ void ar_nt(const T* __restrict__ X, int ldx, int r0, const T* __restrict__ Y, int ldy, int n, int C,
                      float scale, float* __restrict__ Rs, float* __restrict__ U) {
  float* bufX = U;                                // [AR_ROWS][33]
  float* bufY = U + AR_ROWS * 33;                 // [AR_MAXN][33]
  const int tid = threadIdx.x;
  const int kg = tid / 50, t50 = tid % 50;
  const int ti = t50 / 10, tj = t50 % 10;
  float acc[5][10];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int c = 0; c < 10; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < C; k0 += 32) {
    for (int i = tid; i < (AR_ROWS + AR_MAXN) * 8; i += 256) {
      const int row = i >> 3, part = i & 7;       // rows [0, 25): X row r0 + row; rows [25, 125): Y row row - 25
      const bool is_x = row < AR_ROWS;
      const int grow = is_x ? r0 + row : row - AR_ROWS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < n) v = ar_load4((is_x ? X + (size_t)grow * ldx : Y + (size_t)grow * ldy) + k0 + part * 4);
      float* dstp = (is_x ? bufX + row * 33 : bufY + (row - AR_ROWS) * 33) + part * 4;
      dstp[0] = v.x; dstp[1] = v.y; dstp[2] = v.z; dstp[3] = v.w;
    }
    __syncthreads();
    if (tid < 200) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = kg * 8 + kk;
        float xa[5], yb[10];
#pragma unroll
        for (int a = 0; a < 5; ++a) xa[a] = bufX[(ti * 5 + a) * 33 + k];
#pragma unroll
        for (int c = 0; c < 10; ++c) yb[c] = bufY[(tj * 10 + c) * 33 + k];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
          for (int c = 0; c < 10; ++c) acc[a][c] += xa[a] * yb[c];
      }
    }
    __syncthreads();
  }
  if (tid < 200) {
    float* part = U + kg * AR_ROWS * AR_MAXN;     // the chunk buffers are free: 4 partial tiles
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int c = 0; c < 10; ++c) part[(ti * 5 + a) * AR_MAXN + tj * 10 + c] = acc[a][c];
  }
  __syncthreads();
  constexpr int T1 = AR_ROWS * AR_MAXN;
  for (int i = tid; i < T1; i += 256) Rs[i] = ((U[i] + U[T1 + i]) + (U[2 * T1 + i] + U[3 * T1 + i])) * scale;
  __syncthreads();
}

// store(global_row, column, v) receives 4 consecutive output columns of one row
template <typename T, typename Store>
__device__ void ar_sy(const float* __restrict__ Rs, const T* __restrict__ Y, int ldy, int r0, int n, int C,
                      float* __restrict__ U, Store store) {
  const int tid = threadIdx.x;
  const int ri = tid >> 5, cj = tid & 31;
  for (int c0 = 0; c0 < C; c0 += AR_GCH) {
    for (int i = tid; i < AR_MAXN * (AR_GCH / 4); i += 256) {
      const int row = i / (AR_GCH / 4), part = i % (AR_GCH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n) v = ar_load4(Y + (size_t)row * ldy + c0 + part * 4);
      *reinterpret_cast<float4*>(U + row * AR_GCH + part * 4) = v;
    }
    __syncthreads();
    if (ri < 5) {
      float o[5][4];
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) o[a][q] = 0.f;
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const float4 g4 = *reinterpret_cast<const float4*>(U + j * AR_GCH + cj * 4);
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const float rv = Rs[(ri * 5 + a) * AR_MAXN + j];
          o[a][0] += rv * g4.x; o[a][1] += rv * g4.y; o[a][2] += rv * g4.z; o[a][3] += rv * g4.w;
        }
      }
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        const int row = r0 + ri * 5 + a;
        if (row < n) store(row, c0 + cj * 4, o[a]);
      }
    }
    __syncthreads();
  }
}

}  // namespace mv
