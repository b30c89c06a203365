// Backward / optimiser kernels of the fusion model's training step (reference loop: mvuld/main_bigvul.py:294-342;
// model: mvuld/models/GraphModel.py:150-211, mvuld/models/Rs_GCN.py:52-73, DGL GATConv).
//
// Every dense backward (dX = dY W, dW = dY^T X) runs on the tcgen05 GEMM of gemm.cu: the kernels here provide the
// operand transposes, the bias / BatchNorm / activation / dropout backward passes, the sparse GATConv backward
// (edge-softmax backward per destination, gather over out-edges per source), the Rs_GCN affinity backward, the
// l2norm / mean / cross-entropy head, and the clipped AdamW update (optimizer.py:11-33, config.py:157).
// Gradients of activations travel in bf16 (GEMM operands) or fp32 (residual stream, BatchNorm), parameter
// gradients and optimiser state in fp32.
#include <algorithm>
#include <cstdlib>

#include "affinity_rows.cuh"
#include "common.cuh"
#include "host_util.h"

namespace mv {

// ------------------------------------------------ small utilities ------------------------------------------------
// counter-based uniform in [0,1): one 32-bit mix of (seed, index); the backward pass regenerates the same mask
__device__ __forceinline__ float u01(unsigned long long seed, unsigned long long idx) {
  unsigned long long x = idx * 0x9E3779B97F4A7C15ull + seed;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  return (float)(unsigned int)(x >> 40) * (1.0f / 16777216.0f);
}

// out[c, r] = in[r, c] (in row stride ldi); out row stride ldo >= R, columns [R, ldo) zero filled (TMA strides need
// 16-byte multiples)
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, int ldi, bf16* __restrict__ out, int R, int C,
                                      int ldo) {
  __shared__ bf16 tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ldi + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ldo) out[(size_t)c * ldo + r] = tile[threadIdx.x][i];
  }
}
// The same transpose for a TABLE of matrices in one launch (the trainers refresh ~100 transposed weight copies per step:
// 100 launches of a few microseconds each were launch latency, not work).  desc[k] = {in, out, R, C, ldi, ldo} as six
// 64-bit words, tile_end[k] = exclusive end of matrix k's 32 x 32 tiles in the flattened grid.
__global__ void transpose_bf16_batched_kernel(const long long* __restrict__ desc, const int* __restrict__ tile_end,
                                              int nmat) {
  __shared__ bf16 tile[32][33];
  int lo = 0, hi = nmat - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x < tile_end[mid]) hi = mid; else lo = mid + 1;
  }
  const long long* d = desc + (size_t)lo * 6;
  const bf16* in = reinterpret_cast<const bf16*>(d[0]);
  bf16* out = reinterpret_cast<bf16*>(d[1]);
  const int R = (int)d[2], C = (int)d[3], ldi = (int)d[4], ldo = (int)d[5];
  const int t = (int)blockIdx.x - (lo ? tile_end[lo - 1] : 0);
  const int tx = (C + 31) / 32;
  const int r0 = (t / tx) * 32, c0 = (t % tx) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ldi + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ldo) out[(size_t)c * ldo + r] = tile[threadIdx.x][i];
  }
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p, size_t i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p, size_t i) { return __bfloat162float(p[i]); }

// Fixed-order block reduction of NV per-thread values (256 threads): xor-shuffle tree inside each warp, then the 8 warp
// results in warp order.  Thread t < NV returns the total of value t; every run adds in the same order, so the column
// sums below (bias / attention-vector / LayerNorm gradients) are bit-reproducible -- a float atomicAdd across blocks is
// not, and its noise, amplified by the batch-statistics BatchNorms, moved the training loss by 4e-3 after two steps.
template <int NV>
__device__ __forceinline__ float block_reduce_fixed(float (&acc)[NV], float (*part)[NV]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) part[warp][k] = acc[k];
  }
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < NV) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
  }
  return t;
}

// out[c] += sum_r x[r, c].  Block (column segment of up to 256 eight-column units, row slab): thread = (row lane, unit),
// so a warp reads 32 consecutive 16 / 32-byte units of a row (the first version gave a thread one unit of 256 different
// rows: 32 half-used sectors per warp load, and a [25 088, 512] gradient took 21 us against 4 us of HBM time); the row
// lanes of a unit are summed in lane order through shared memory; with one slab the block adds straight into out,
// otherwise it writes its row of the partials workspace and colsum_final_kernel sums the slabs in a fixed order.
// No atomics: bit-reproducible.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int ldx, float* __restrict__ out, float* __restrict__ partials, int R, int C,
              int rows_per_slab) {
  __shared__ float red[256 * 8];
  const int units_total = (C + 7) >> 3;
  const int u0 = blockIdx.x * 256;
  const int upr = min(256, units_total - u0);          // units of a row this block covers
  const int rpp = 256 / upr;                           // rows per pass
  const int tr = threadIdx.x / upr, tu = threadIdx.x - tr * upr;
  const bool active = tr < rpp;
  const int c0 = (u0 + tu) * 8;
  const int nc = min(8, C - c0);
  const int rbeg = blockIdx.y * rows_per_slab, rend = min(R, rbeg + rows_per_slab);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec = nc == 8 && (ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 31) == 0;
  if (active) {
    if (vec) {
      // four rows in flight per thread: the loop carries only the adds
      int r = rbeg + tr;
      for (; r + 3 * rpp < rend; r += 4 * rpp) {
        if (sizeof(T) == 2) {
          uint4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(r + j * rpp) * ldx + c0));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0] += bf16_lo(v[j].x); acc[1] += bf16_hi(v[j].x); acc[2] += bf16_lo(v[j].y); acc[3] += bf16_hi(v[j].y);
            acc[4] += bf16_lo(v[j].z); acc[5] += bf16_hi(v[j].z); acc[6] += bf16_lo(v[j].w); acc[7] += bf16_hi(v[j].w);
          }
        } else {
          float4 a4[4], b4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4* p = reinterpret_cast<const float4*>(x + (size_t)(r + j * rpp) * ldx + c0);
            a4[j] = __ldg(p);
            b4[j] = __ldg(p + 1);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0] += a4[j].x; acc[1] += a4[j].y; acc[2] += a4[j].z; acc[3] += a4[j].w;
            acc[4] += b4[j].x; acc[5] += b4[j].y; acc[6] += b4[j].z; acc[7] += b4[j].w;
          }
        }
      }
      for (; r < rend; r += rpp) {
        const T* p = x + (size_t)r * ldx + c0;
        if (sizeof(T) == 2) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
          acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
          acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        } else {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(p)), b4 = __ldg(reinterpret_cast<const float4*>(p) + 1);
          acc[0] += a4.x; acc[1] += a4.y; acc[2] += a4.z; acc[3] += a4.w;
          acc[4] += b4.x; acc[5] += b4.y; acc[6] += b4.z; acc[7] += b4.w;
        }
      }
    } else {
      for (int r = rbeg + tr; r < rend; r += rpp) {
        const T* p = x + (size_t)r * ldx + c0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nc) acc[k] += ldf<T>(p, k);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  if (active && tr == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
      for (int q = 0; q < rpp; ++q) t += red[(q * upr + tu) * 8 + k];
      if (k < nc) {
        if (gridDim.y == 1) out[c0 + k] += t;
        else partials[(size_t)blockIdx.y * C + c0 + k] = t;
      }
    }
  }
}
// slabs summed in a fixed order, split over 32 thread rows (see ln_rows_bwd_final_kernel)
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ partials, float* __restrict__ out, int slabs, int C) {
  __shared__ float sa[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, slice = threadIdx.y;
  const int per = (slabs + 31) / 32;
  const int k0 = slice * per, k1 = min(slabs, k0 + per);
  float a = 0.f;
  if (c < C)
    for (int k = k0; k < k1; ++k) a += partials[(size_t)k * C + c];
  sa[slice][threadIdx.x] = a;
  __syncthreads();
  if (slice == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) t += sa[s2][threadIdx.x];
    out[c] += t;
  }
}

// activation backward through y = dropout(elu(pre)):  kept elements carry elu(pre) / (1 - p)
__global__ void elu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, bf16* __restrict__ dx,
                               long long n, unsigned long long seed, float p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = __bfloat162float(dy[i]);
  float e = __bfloat162float(y[i]);
  if (p > 0.f) {
    const bool keep = u01(seed, (unsigned long long)i) >= p;
    const float inv = 1.0f / (1.0f - p);
    g = keep ? g * inv : 0.f;
    e = e * (1.0f - p);                     // elu(pre) of a kept element (dropped ones have g == 0 anyway)
  }
  dx[i] = __float2bfloat16(g * (e > 0.f ? 1.0f : e + 1.0f));
}
__global__ void elu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                   long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = y[i];
  dx[i] = dy[i] * (e > 0.f ? 1.0f : e + 1.0f);
}
// out = mask * x / (1 - p)   (forward dropout, and the backward of a dropout applied to an input)
__global__ void dropout_bf16_kernel(const bf16* x, bf16* out, long long n,
                                    unsigned long long seed, float p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool keep = u01(seed, (unsigned long long)i) >= p;
  out[i] = keep ? __float2bfloat16(__bfloat162float(x[i]) / (1.0f - p)) : __float2bfloat16(0.f);
}

// ------------------------------------------- BatchNorm over columns -------------------------------------------
// x fp32 [R, C], statistics per column over the R rows (biased variance, F.batch_norm training semantics).
// One block per BN_CL = 8 columns (a 32-byte sector per row), BN_RL = 32 row lanes; two passes (mean, then centred
// second moment: the Rs_GCN outputs have |mean| >> deviation, a one-pass E[x^2] - E[x]^2 would cancel).  8-column
// blocks give C / 8 CTAs (64 for C = 512) where 32-column blocks left 16 CTAs on 148 SMs (105 us per launch).
constexpr int BN_CL = 8, BN_RL = 32;
// Column sums run in fp64: the Rs_GCN BatchNorms normalise outputs whose batch deviation is small against their
// magnitude, and their backward pass subtracts nearly equal sums (dy - mean(dy) - xhat mean(dy xhat)); the fp32
// rounding of those sums was visible (percent level) in the gradients of the deepest block after 8 such stages.
__device__ __forceinline__ double bn_col_total(double v, double (*part)[BN_CL + 1], int ry, int cl) {
  part[ry][cl] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < BN_RL; ++k) t += part[k][cl];
  __syncthreads();
  return t;                                  // every thread of column cl gets the same fixed-order total
}
__global__ void __launch_bounds__(BN_CL * BN_RL)
bn_cols_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, const float* res, int ldr, float* y32, int ldy, bf16* __restrict__ yb,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ run_mean,
                   float* __restrict__ run_var, float momentum, int R, int C) {
  __shared__ double part[BN_RL][BN_CL + 1];
  const int cl = threadIdx.x % BN_CL, ry = threadIdx.x / BN_CL;
  const int c = blockIdx.x * BN_CL + cl;
  const bool ok = c < C;
  double s = 0.0;
  if (ok)
    for (int r = ry; r < R; r += BN_RL) s += (double)x[(size_t)r * C + c];
  const double mean_d = bn_col_total(s, part, ry, cl) / (double)R;
  const float mean = (float)mean_d;
  double q = 0.0;
  if (ok)
    for (int r = ry; r < R; r += BN_RL) {
      const double d = (double)x[(size_t)r * C + c] - mean_d;
      q += d * d;
    }
  const float var = (float)(bn_col_total(q, part, ry, cl) / (double)R);
  const float rstd = rsqrtf(var + eps);
  if (!ok) return;
  if (ry == 0) {
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    if (run_mean) {                       // nn.BatchNorm1d running statistics (unbiased variance, momentum 0.1)
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * ((float)R / (float)max(R - 1, 1));
    }
  }
  const float g = gamma[c], b = beta[c];
  for (int r = ry; r < R; r += BN_RL) {
    float v = (x[(size_t)r * C + c] - mean) * rstd * g + b;
    if (res) v += res[(size_t)r * ldr + c];            // Rs_GCN.py:70: W_y + v (res may alias y32)
    if (y32) y32[(size_t)r * ldy + c] = v;
    if (yb) yb[(size_t)r * C + c] = __float2bfloat16(v);
  }
}

// dx = gamma rstd / R * (R dy - sum(dy) - xhat sum(dy xhat));  dgamma += sum(dy xhat);  dbeta += sum(dy)
__global__ void __launch_bounds__(BN_CL * BN_RL)
bn_cols_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int ldy, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx32,
                   bf16* __restrict__ dxb, float* __restrict__ dgamma, float* __restrict__ dbeta, int R, int C) {
  __shared__ double part[BN_RL][BN_CL + 1];
  const int cl = threadIdx.x % BN_CL, ry = threadIdx.x / BN_CL;
  const int c = blockIdx.x * BN_CL + cl;
  const bool ok = c < C;
  const float m = ok ? mean[c] : 0.f, rs = ok ? rstd[c] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (ok)
    for (int r = ry; r < R; r += BN_RL) {
      const float g = dy[(size_t)r * ldy + c];
      s1 += (double)g;
      s2 += (double)g * (double)((x[(size_t)r * C + c] - m) * rs);
    }
  const float a = (float)bn_col_total(s1, part, ry, cl);
  const float b = (float)bn_col_total(s2, part, ry, cl);
  if (!ok) return;
  if (ry == 0) {
    dbeta[c] += a;
    dgamma[c] += b;
  }
  if (!dx32 && !dxb) return;
  const float k = gamma[c] * rs / (float)R;
  for (int r = ry; r < R; r += BN_RL) {
    const float xh = (x[(size_t)r * C + c] - m) * rs;
    const float v = k * ((float)R * dy[(size_t)r * ldy + c] - a - xh * b);
    if (dx32) dx32[(size_t)r * C + c] = v;
    if (dxb) dxb[(size_t)r * C + c] = __float2bfloat16(v);
  }
}

// --------------------------------- BatchNorm1d(max_node) over the node-slot axis ---------------------------------
// x bf16 [B, n, F]: channel = slot r; statistics over the B * F values of the slot (GraphModel.py:135,186).
// One block per slot.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}
__global__ void __launch_bounds__(256)
bn_slot_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                   float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, int B, int n, int F) {
  __shared__ float sh[8];
  const int r = blockIdx.x;
  const int cnt = B * F;
  float s = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x)
    s += __bfloat162float(x[((size_t)(i / F) * n + r) * F + i % F]);
  const float mean = block_sum(s, sh) / (float)cnt;
  float q = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float d = __bfloat162float(x[((size_t)(i / F) * n + r) * F + i % F]) - mean;
    q += d * d;
  }
  const float var = block_sum(q, sh) / (float)cnt;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
    if (run_mean) {
      run_mean[r] = (1.f - momentum) * run_mean[r] + momentum * mean;
      run_var[r] = (1.f - momentum) * run_var[r] + momentum * var * ((float)cnt / (float)max(cnt - 1, 1));
    }
  }
  const float g = gamma[r], b = beta[r];
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const size_t o = ((size_t)(i / F) * n + r) * F + i % F;
    y[o] = __float2bfloat16((__bfloat162float(x[o]) - mean) * rstd * g + b);
  }
}
__global__ void __launch_bounds__(256)
bn_slot_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, bf16* __restrict__ dx,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int n, int F) {
  __shared__ float sh[8];
  const int r = blockIdx.x;
  const int cnt = B * F;
  const float m = mean[r], rs = rstd[r];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const size_t o = ((size_t)(i / F) * n + r) * F + i % F;
    const float g = __bfloat162float(dy[o]);
    s1 += g;
    s2 += g * (__bfloat162float(x[o]) - m) * rs;
  }
  const float a = block_sum(s1, sh);
  const float b = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    dbeta[r] += a;
    dgamma[r] += b;
  }
  if (!dx) return;
  const float k = gamma[r] * rs / (float)cnt;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const size_t o = ((size_t)(i / F) * n + r) * F + i % F;
    const float xh = (__bfloat162float(x[o]) - m) * rs;
    dx[o] = __float2bfloat16(k * ((float)cnt * __bfloat162float(dy[o]) - a - xh * b));
  }
}


// strided fp32 ELU backward with a bf16 result: dx[r, c] = dy[r, c] * ELU'(pre) from y = ELU(pre)  (image / text
// projections, GraphModel.py:153-159: dy and y are column slices of the [B, 1536] feature row)
__global__ void elu_bwd_rows_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ y, int ldyy,
                                    bf16* __restrict__ dx, int ldx, int R, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * C) return;
  const int r = (int)(i / C), c = (int)(i % C);
  const float e = y[(size_t)r * ldyy + c];
  dx[(size_t)r * ldx + c] = __float2bfloat16(dy[(size_t)r * ldy + c] * (e > 0.f ? 1.0f : e + 1.0f));
}

// ------------------------------ bounding-box branch: BatchNorm1d(max_node) statistics ------------------------------
// pos fp32 [N, 4] -> padded [B, n, 4] (zeros past a graph's node count, GraphModel.py:30-54); slot r's statistics run
// over its B * 4 values (GraphModel.py:137,187).  Emits the affine (scale, shift) that mvuld_pos_branch applies.
__global__ void __launch_bounds__(128)
pos_slot_stats_kernel(const float* __restrict__ pos, const long long* __restrict__ off, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float* __restrict__ scale, float* __restrict__ shift,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ run_mean,
                      float* __restrict__ run_var, float momentum, int B, int n) {
  __shared__ float sh[8];
  const int r = blockIdx.x;
  const int cnt = B * 4;
  float s = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int b = i >> 2;
    const long long beg = off[b], nn = off[b + 1] - beg;
    s += r < nn ? pos[(beg + r) * 4 + (i & 3)] : 0.f;
  }
  const float mean = block_sum(s, sh) / (float)cnt;
  float q = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int b = i >> 2;
    const long long beg = off[b], nn = off[b + 1] - beg;
    const float d = (r < nn ? pos[(beg + r) * 4 + (i & 3)] : 0.f) - mean;
    q += d * d;
  }
  const float var = block_sum(q, sh) / (float)cnt;
  if (threadIdx.x == 0) {
    const float rstd = rsqrtf(var + eps);
    mean_out[r] = mean;
    rstd_out[r] = rstd;
    scale[r] = gamma[r] * rstd;
    shift[r] = beta[r] - mean * gamma[r] * rstd;
    if (run_mean) {
      run_mean[r] = (1.f - momentum) * run_mean[r] + momentum * mean;
      run_var[r] = (1.f - momentum) * run_var[r] + momentum * var * ((float)cnt / (float)max(cnt - 1, 1));
    }
  }
}
// backward of ELU(fc_bbox(bn_bbox(pad(pos)))): dpre bf16 = columns [col0, col0 + 32) of the [B*n, ld] gradient (ELU'
// already applied); accumulates dW [32, 4], db [32], dgamma [n], dbeta [n].  One block per slot, warp = 32 outputs.
__global__ void __launch_bounds__(128)
pos_branch_bwd_kernel(const float* __restrict__ pos, const long long* __restrict__ off, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ w, const bf16* __restrict__ dpre,
                      float* __restrict__ partials, float* __restrict__ dgamma, float* __restrict__ dbeta, int B,
                      int n, int ld, int col0) {
  const int r = blockIdx.x;
  const int o = threadIdx.x & 31, q = threadIdx.x >> 5;
  const float m = mean[r], rs = rstd[r], ga = gamma[r], be = beta[r];
  float wk[4], aw[4] = {0.f, 0.f, 0.f, 0.f}, ab = 0.f, ag = 0.f, abt = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) wk[k] = w[o * 4 + k];
  for (int b = q; b < B; b += 4) {
    const long long beg = off[b], nn = off[b + 1] - beg;
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < nn) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pos + (beg + r) * 4));
      p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    }
    const float d = __bfloat162float(dpre[((size_t)b * n + r) * ld + col0 + o]);
    ab += d;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (p[k] - m) * rs;
      aw[k] += d * (xh * ga + be);
      const float dxn = warp_sum(d * wk[k]);          // d L / d bn_out[b, r, k]
      ag += dxn * xh;
      abt += dxn;
    }
  }
  // this slot's contribution: the 4 warps in warp order (fixed), then one row of the partials buffer; the sum over slots
  // is taken in slot order by pos_branch_bwd_final_kernel -- bit-reproducible, no atomics
  __shared__ float sh[4][162];
#pragma unroll
  for (int k = 0; k < 4; ++k) sh[q][o * 4 + k] = aw[k];
  sh[q][128 + o] = ab;
  if (o == 0) {
    sh[q][160] = ag;
    sh[q][161] = abt;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 162; i += 128) {
    const float t = ((sh[0][i] + sh[1][i]) + sh[2][i]) + sh[3][i];
    if (i < 160) partials[(size_t)r * 160 + i] = t;
    else if (i == 160) dgamma[r] += t;
    else dbeta[r] += t;
  }
}
__global__ void pos_branch_bwd_final_kernel(const float* __restrict__ partials, float* __restrict__ dw,
                                            float* __restrict__ db, int n) {
  const int i = threadIdx.x;                                 // 160 threads: dW [32, 4] then db [32]
  float t = 0.f;
  for (int r = 0; r < n; ++r) t += partials[(size_t)r * 160 + i];
  if (i < 128) dw[i] += t;
  else db[i - 128] += t;
}

// backward of unbatch_features pad / truncate (GraphModel.py:30-54): dh[node] = dhp[b, r] for r < max_node, else 0
__global__ void unbatch_pad_bwd_kernel(const bf16* __restrict__ dhp, const long long* __restrict__ off,
                                       bf16* __restrict__ dh, int B, int max_node, int F) {
  const int b = blockIdx.y;
  const long long beg = off[b], cnt = off[b + 1] - beg;
  const int units = F >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt * units; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / units;
    const int u = (int)(i % units);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < max_node) v = reinterpret_cast<const uint4*>(dhp + ((size_t)b * max_node + r) * F)[u];
    reinterpret_cast<uint4*>(dh + (size_t)(beg + r) * F)[u] = v;
  }
}

// ------------------------------------------------ GATConv backward ------------------------------------------------
// Pass 1, one warp per destination v: recompute the edge softmax, a_e = <dout[v,h,:], z[u,h,:]>,
// ds_e = alpha_e (a_e - sum alpha a) * leaky'(el[u] + er[v]); writes alpha and ds per in-CSR position and
// der[v,h] = sum_e ds_e.  z, dout bf16 [N, H*F]; F % 256 == 0; H <= 4.
template <int NH>
__global__ void __launch_bounds__(128)
gat_bwd_dst_kernel(const bf16* __restrict__ z, const bf16* __restrict__ dout, const float* __restrict__ el,
                   const float* __restrict__ er, const int* __restrict__ indptr, const int* __restrict__ idx_src,
                   float* __restrict__ alpha_e, float* __restrict__ ds_e, float* __restrict__ der, int N, int F,
                   float slope) {
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int lane = threadIdx.x & 31;
  const int beg = __ldg(indptr + node), end = __ldg(indptr + node + 1);
  float hmax[NH], hsum[NH], erd[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    hmax[h] = -INFINITY;
    hsum[h] = 0.f;
    erd[h] = __ldg(er + (size_t)node * NH + h);
  }
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const int s = e < end ? __ldg(idx_src + e) : 0;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float sc = -INFINITY;
      if (e < end) {
        sc = __ldg(el + (size_t)s * NH + h) + erd[h];
        sc = sc > 0.f ? sc : sc * slope;
      }
      hmax[h] = fmaxf(hmax[h], warp_max(sc));
    }
  }
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const int s = e < end ? __ldg(idx_src + e) : 0;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float pe = 0.f;
      if (e < end) {
        float sc = __ldg(el + (size_t)s * NH + h) + erd[h];
        sc = sc > 0.f ? sc : sc * slope;
        pe = __expf(sc - hmax[h]);
      }
      hsum[h] += warp_sum(pe);
    }
  }
  // a_e per edge (warp-cooperative dot over the H*F row), accumulated t_h = sum alpha a
  const int upr = F >> 3;                         // uint4 per head
  float th[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) th[h] = 0.f;
  const uint4* drow = reinterpret_cast<const uint4*>(dout + (size_t)node * NH * F);
  for (int e = beg; e < end; ++e) {
    const int s = __ldg(idx_src + e);
    const uint4* zrow = reinterpret_cast<const uint4*>(z + (size_t)s * NH * F);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float a = 0.f;
      for (int u = lane; u < upr; u += 32) {
        const uint4 zv = __ldg(zrow + h * upr + u), dv = __ldg(drow + h * upr + u);
        a += bf16_lo(zv.x) * bf16_lo(dv.x) + bf16_hi(zv.x) * bf16_hi(dv.x) + bf16_lo(zv.y) * bf16_lo(dv.y) +
             bf16_hi(zv.y) * bf16_hi(dv.y) + bf16_lo(zv.z) * bf16_lo(dv.z) + bf16_hi(zv.z) * bf16_hi(dv.z) +
             bf16_lo(zv.w) * bf16_lo(dv.w) + bf16_hi(zv.w) * bf16_hi(dv.w);
      }
      a = warp_sum(a);
      float sc = __ldg(el + (size_t)s * NH + h) + erd[h];
      sc = sc > 0.f ? sc : sc * slope;
      const float al = __expf(sc - hmax[h]) / hsum[h];
      th[h] += al * a;
      if (lane == 0) {
        alpha_e[(size_t)e * NH + h] = al;
        ds_e[(size_t)e * NH + h] = a;             // a_e for now; turned into ds_e below
      }
    }
  }
  __syncwarp();
  float dsum[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) dsum[h] = 0.f;
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    if (e < end) {
      const int s = __ldg(idx_src + e);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float pre = __ldg(el + (size_t)s * NH + h) + erd[h];
        const float d = alpha_e[(size_t)e * NH + h] * (ds_e[(size_t)e * NH + h] - th[h]) * (pre > 0.f ? 1.0f : slope);
        ds_e[(size_t)e * NH + h] = d;
        dsum[h] += d;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float t = warp_sum(dsum[h]);
    if (lane == 0) der[(size_t)node * NH + h] = t;
  }
}

// Pass 2, one warp per source u over its out-edges (out-CSR sorted by src; pos_in[k] = in-CSR position of out-edge k):
//   dz[u,h,:] = sum_e alpha_e dout[dst_e,h,:] + del[u,h] attn_l[h,:] + der[u,h] attn_r[h,:],  del[u,h] = sum_e ds_e
template <int NH>
__global__ void __launch_bounds__(128)
gat_bwd_src_kernel(const bf16* __restrict__ dout, const float* __restrict__ alpha_e, const float* __restrict__ ds_e,
                   const float* __restrict__ der, const int* __restrict__ out_indptr, const int* __restrict__ out_dst,
                   const int* __restrict__ pos_in, const float* __restrict__ attn_l, const float* __restrict__ attn_r,
                   bf16* __restrict__ dz, float* __restrict__ del, int N, int F) {
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int lane = threadIdx.x & 31;
  const int beg = __ldg(out_indptr + node), end = __ldg(out_indptr + node + 1);
  float dl[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) dl[h] = 0.f;
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    if (k < end) {
      const int pe = __ldg(pos_in + k);
#pragma unroll
      for (int h = 0; h < NH; ++h) dl[h] += __ldg(ds_e + (size_t)pe * NH + h);
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    dl[h] = warp_sum(dl[h]);
    if (lane == 0) del[(size_t)node * NH + h] = dl[h];
  }
  const int upr = F >> 3;
  for (int h = 0; h < NH; ++h) {
    const float drh = __ldg(der + (size_t)node * NH + h);
    for (int u = lane; u < upr; u += 32) {
      float acc[8];
      const float4* lp = reinterpret_cast<const float4*>(attn_l + h * F + u * 8);
      const float4* rp = reinterpret_cast<const float4*>(attn_r + h * F + u * 8);
      const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), r0 = __ldg(rp), r1 = __ldg(rp + 1);
      acc[0] = dl[h] * l0.x + drh * r0.x; acc[1] = dl[h] * l0.y + drh * r0.y;
      acc[2] = dl[h] * l0.z + drh * r0.z; acc[3] = dl[h] * l0.w + drh * r0.w;
      acc[4] = dl[h] * l1.x + drh * r1.x; acc[5] = dl[h] * l1.y + drh * r1.y;
      acc[6] = dl[h] * l1.z + drh * r1.z; acc[7] = dl[h] * l1.w + drh * r1.w;
      for (int k = beg; k < end; ++k) {
        const int d = __ldg(out_dst + k);
        const float al = __ldg(alpha_e + (size_t)__ldg(pos_in + k) * NH + h);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)d * NH * F) + h * upr + u);
        acc[0] += al * bf16_lo(v.x); acc[1] += al * bf16_hi(v.x); acc[2] += al * bf16_lo(v.y); acc[3] += al * bf16_hi(v.y);
        acc[4] += al * bf16_lo(v.z); acc[5] += al * bf16_hi(v.z); acc[6] += al * bf16_lo(v.w); acc[7] += al * bf16_hi(v.w);
      }
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
      reinterpret_cast<uint4*>(dz + (size_t)node * NH * F)[h * upr + u] = o;
    }
  }
}

// dattn_l[h,f] += sum_n del[n,h] z[n,h,f];  dattn_r likewise with der.  One block owns 8 columns of the H*F row over
// ALL nodes (thread = node lane, one 16-byte load per node); fixed-order block reduction, no atomics.
__global__ void __launch_bounds__(256)
gat_attn_grad_kernel(const bf16* __restrict__ z, const float* __restrict__ del, const float* __restrict__ der,
                     float* __restrict__ dal, float* __restrict__ dar, int N, int H, int F) {
  __shared__ float part[8][16];
  const int col0 = blockIdx.x * 8;                          // F % 8 == 0: the 8 columns share a head
  const int h = col0 / F;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(z + (size_t)n * H * F + col0));
    const float zl[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                         bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
    const float a = __ldg(del + (size_t)n * H + h), b = __ldg(der + (size_t)n * H + h);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] += a * zl[k];
      acc[8 + k] += b * zl[k];
    }
  }
  const float t = block_reduce_fixed<16>(acc, part);
  if (threadIdx.x < 8) dal[col0 + threadIdx.x] += t;
  else if (threadIdx.x < 16) dar[col0 + threadIdx.x - 8] += t;
}

// ------------------------------------- encoder backward row kernels (first pieces) -------------------------------------
// LayerNorm backward of the three forward forms of mvuld_ln_rows (mode 0: x = LN(y); 1: x = shortcut + LN(y), SwinV2
// res-post-norm, swin_transformer_v2.py:301,304; 2: x = LN(y + shortcut), RoBERTa).  The LN input v (y, or y + shortcut
// in mode 2) is recomputed from what the forward already keeps (y bf16, shortcut fp32): no saved statistics.
//   xhat = (v - mean) rstd;  g = dout gamma;  dv = rstd (g - mean(g) - xhat mean(g xhat));
//   dgamma += sum_rows dout xhat;  dbeta += sum_rows dout
// One warp walks rows (grid-stride), lane owns 8-element units; the per-lane column sums stay in registers over all
// rows of the warp, are combined over the 8 warps of the block in shared memory and leave as one row of the partials
// workspace per block (grid capped at 2 blocks per SM), which ln_rows_bwd_final_kernel sums in block order: no atomics,
// bit-reproducible gradients.  dv is written as bf16 (the operand of the dense backward) and / or fp32
// (mode 2: it is also the shortcut's gradient); the residual gradient of mode 1 is dout itself and needs no kernel.
// Latency: a warp keeps RPW = NR * (32 / LPR) rows in flight (LPR = 16 lanes per row when a row has <= 16 units, so a
// C = 128 row does not idle half the warp), and y, the shortcut and dout of all of them are requested before the first
// reduction (one row per warp with dout loaded after two shuffle reductions ran the C = 128 launches at 1.8 TB/s).
// DB: also the column sums of dv (= the bias gradient of the dense layer in front of the LayerNorm) as a third row of the
// block's partials: saves a separate pass over dv.
template <int UNITS, int LPR, int NR, bool DB>
__global__ void __launch_bounds__(256, 2)          // two resident blocks: the C = 768 / 1024 forms would take 151+ registers
ln_rows_bwd_kernel(const bf16* __restrict__ y, const float* __restrict__ shortcut, const float* __restrict__ gamma,
                   const float* __restrict__ dout, bf16* __restrict__ dvb, float* __restrict__ dv32,
                   float* __restrict__ partials, int M, int C, float eps, int mode) {
  extern __shared__ float red[];                  // [8 warps][32 / LPR][NP][C]
  constexpr int NP = DB ? 3 : 2;                  // rows of a block's partials: dgamma, dbeta (, dbias)
  constexpr int RPP = 32 / LPR;                   // rows per pass of a warp
  constexpr int RPW = RPP * NR;                   // rows a warp has in flight
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int units = C >> 3;
  const float invC = 1.0f / (float)C;
  auto seg_sum = [](float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  float dg[UNITS][8], db[UNITS][8], dbi[DB ? UNITS : 1][8];
#pragma unroll
  for (int k = 0; k < UNITS; ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q) dg[k][q] = db[k][q] = 0.f;
#pragma unroll
  for (int k = 0; k < (DB ? UNITS : 1); ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q) dbi[k][q] = 0.f;
  for (int row0 = (blockIdx.x * 8 + warp) * RPW; row0 < M; row0 += gridDim.x * 8 * RPW) {
    float v[NR][UNITS][8], d[NR][UNITS][8];
#pragma unroll
    for (int n = 0; n < NR; ++n) {
      const int row = row0 + n * RPP + sub;
#pragma unroll
      for (int k = 0; k < UNITS; ++k) {
        const int u = l + k * LPR;
        if (row < M && u < units) {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(y + (size_t)row * C) + u);
          v[n][k][0] = bf16_lo(raw.x); v[n][k][1] = bf16_hi(raw.x); v[n][k][2] = bf16_lo(raw.y); v[n][k][3] = bf16_hi(raw.y);
          v[n][k][4] = bf16_lo(raw.z); v[n][k][5] = bf16_hi(raw.z); v[n][k][6] = bf16_lo(raw.w); v[n][k][7] = bf16_hi(raw.w);
          if (mode == 2) {
            const float4* sp = reinterpret_cast<const float4*>(shortcut + (size_t)row * C + u * 8);
            const float4 a = __ldg(sp), b = __ldg(sp + 1);
            v[n][k][0] += a.x; v[n][k][1] += a.y; v[n][k][2] += a.z; v[n][k][3] += a.w;
            v[n][k][4] += b.x; v[n][k][5] += b.y; v[n][k][6] += b.z; v[n][k][7] += b.w;
          }
          const float4* dp = reinterpret_cast<const float4*>(dout + (size_t)row * C + u * 8);
          const float4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
          d[n][k][0] = d0.x; d[n][k][1] = d0.y; d[n][k][2] = d0.z; d[n][k][3] = d0.w;
          d[n][k][4] = d1.x; d[n][k][5] = d1.y; d[n][k][6] = d1.z; d[n][k][7] = d1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[n][k][q] = d[n][k][q] = 0.f;
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NR; ++n) {
      const int row = row0 + n * RPP + sub;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < UNITS; ++k)
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += v[n][k][q];
      const float mean = seg_sum(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < UNITS; ++k)
        if (l + k * LPR < units) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float dd = v[n][k][q] - mean;
            sq += dd * dd;
          }
        }
      const float rstd = rsqrtf(seg_sum(sq) * invC + eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < UNITS; ++k) {
        const int u = l + k * LPR;
        if (u < units) {
          const float4* gp = reinterpret_cast<const float4*>(gamma + u * 8);
          const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float xh = (v[n][k][q] - mean) * rstd;
            const float dq = d[n][k][q];
            v[n][k][q] = xh;
            dg[k][q] += dq * xh;                     // rows past M contribute dq = 0
            db[k][q] += dq;
            d[n][k][q] = dq * gm[q];
            s1 += d[n][k][q];
            s2 += d[n][k][q] * xh;
          }
        }
      }
      const float m1 = seg_sum(s1) * invC, m2 = seg_sum(s2) * invC;
#pragma unroll
      for (int k = 0; k < UNITS; ++k) {
        const int u = l + k * LPR;
        if (row < M && u < units) {
          float o[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = rstd * (d[n][k][q] - m1 - v[n][k][q] * m2);
          if (DB) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dbi[k][q] += o[q];
          }
          if (dv32) {
            float4* op = reinterpret_cast<float4*>(dv32 + (size_t)row * C + u * 8);
            op[0] = make_float4(o[0], o[1], o[2], o[3]);
            op[1] = make_float4(o[4], o[5], o[6], o[7]);
          }
          if (dvb) {
            uint4 w;
            w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
            w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
            reinterpret_cast<uint4*>(dvb + (size_t)row * C)[u] = w;
          }
        }
      }
    }
  }
  // column sums: 8 warps x RPP row lanes -> shared memory -> one row of the partials workspace per block (fixed order)
#pragma unroll
  for (int k = 0; k < UNITS; ++k) {
    const int u = l + k * LPR;
    if (u < units) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        red[((warp * RPP + sub) * NP + 0) * C + u * 8 + q] = dg[k][q];
        red[((warp * RPP + sub) * NP + 1) * C + u * 8 + q] = db[k][q];
        if (DB) red[((warp * RPP + sub) * NP + 2) * C + u * 8 + q] = dbi[k][q];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) a[j] = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8 * RPP; ++w8)
#pragma unroll
      for (int j = 0; j < NP; ++j) a[j] += red[(w8 * NP + j) * C + c];
#pragma unroll
    for (int j = 0; j < NP; ++j)
      partials[((size_t)blockIdx.x * NP + j) * C + c] = a[j];   // block order is fixed by ln_rows_bwd_final_kernel
  }
}
// Sum of the per-block partials in a FIXED order with the blocks split over 32 thread rows: thread (slice, column) adds
// its contiguous run of blocks in ascending order, the 32 slice sums are added in slice order.  (One thread per column
// walking all ~900 blocks was a serial chain of dependent L2 round trips on two SMs: ~100 us of the 133 us a C = 512
// LayerNorm backward took.)
__global__ void __launch_bounds__(1024)
ln_rows_bwd_final_kernel(const float* __restrict__ partials, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dbias, int nblocks, int C) {
  __shared__ float sa[3][32][33];
  const int NP = dbias ? 3 : 2;
  const int c = blockIdx.x * 32 + threadIdx.x, slice = threadIdx.y;
  const int per = (nblocks + 31) / 32;
  const int k0 = slice * per, k1 = min(nblocks, k0 + per);
  float a[3] = {0.f, 0.f, 0.f};
  if (c < C) {
    for (int k = k0; k < k1; ++k)
      for (int j = 0; j < NP; ++j) a[j] += partials[((size_t)k * NP + j) * C + c];
  }
  for (int j = 0; j < NP; ++j) sa[j][slice][threadIdx.x] = a[j];
  __syncthreads();
  if (slice < NP && c < C) {                       // thread row j finishes output j
    float t = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) t += sa[slice][s2][threadIdx.x];
    float* out = slice == 0 ? dgamma : (slice == 1 ? dbeta : dbias);
    out[c] += t;
  }
}

// exact (erf) GELU backward, nn.GELU default (swin_transformer_v2.py:26-32 Mlp, HF "gelu"): dpre = dh (Phi(x) + x phi(x))
__global__ void gelu_bwd_kernel(const bf16* __restrict__ pre, const bf16* __restrict__ dh, bf16* __restrict__ dpre,
                                long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 xr = __ldg(reinterpret_cast<const uint4*>(pre) + i), dr = __ldg(reinterpret_cast<const uint4*>(dh) + i);
    const uint32_t xs[4] = {xr.x, xr.y, xr.z, xr.w}, ds[4] = {dr.x, dr.y, dr.z, dr.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x0 = bf16_lo(xs[q]), x1 = bf16_hi(xs[q]);
      const float c0 = gelu_erf_grad(x0), c1 = gelu_erf_grad(x1);
      o[q] = pack_bf16x2(bf16_lo(ds[q]) * c0, bf16_hi(ds[q]) * c1);
    }
    reinterpret_cast<uint4*>(dpre)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// --------------------------------------------- Rs_GCN affinity backward ---------------------------------------------
// Per graph (n <= 100 slots, C features): theta, phi, g = column blocks of tpg bf16 [n, 3C]; dy bf16 [n, C].
//   R = theta phi^T / n;  dg = R^T dy;  dS = dy g^T / n;  dtheta = dS phi;  dphi = dS^T theta   -> dtpg bf16 [n, 3C]
constexpr int AB_N = 100;
// S[i][j] = scale * sum_c X[i, c] Y[j, c]   (X, Y: bf16 rows with stride ld, C columns); 256 threads, 200 active
__device__ void aff_nt(const bf16* __restrict__ X, int ldx, const bf16* __restrict__ Y, int ldy, int n, int C,
                       float scale, float* __restrict__ S, float* bufA, float* bufB) {
  const int tid = threadIdx.x;
  const int ti = tid / 10, tj = tid % 10;
  float acc[5][10];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int c = 0; c < 10; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < C; k0 += 32) {
    for (int i = tid; i < AB_N * 8; i += 256) {
      const int row = i >> 3, part = i & 7;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < n) {
        const bf16* src = (part < 4 ? X + (size_t)row * ldx : Y + (size_t)row * ldy) + k0 + (part & 3) * 8;
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
      }
      float* dstp = (part < 4 ? bufA : bufB) + row * 33 + (part & 3) * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) dstp[q] = f[q];
    }
    __syncthreads();
    if (tid < 200) {
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        float xa[5], yb[10];
#pragma unroll
        for (int a = 0; a < 5; ++a) xa[a] = bufA[(ti * 5 + a) * 33 + k];
#pragma unroll
        for (int c = 0; c < 10; ++c) yb[c] = bufB[(tj * 10 + c) * 33 + k];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
          for (int c = 0; c < 10; ++c) acc[a][c] += xa[a] * yb[c];
      }
    }
    __syncthreads();
  }
  if (tid < 200) {
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int c = 0; c < 10; ++c) S[(ti * 5 + a) * (AB_N + 1) + tj * 10 + c] = acc[a][c] * scale;
  }
  __syncthreads();
}
// out[i, c] = sum_j (TRANS ? S[j][i] : S[i][j]) Y[j, c]   -> bf16 rows (stride ldo)
template <bool TRANS>
__device__ void aff_sy(const float* __restrict__ S, const bf16* __restrict__ Y, int ldy, int n, int c_beg, int c_end,
                       bf16* __restrict__ out, int ldo, float* bufA) {
  const int tid = threadIdx.x;
  const int yi = tid / 8, yj = tid % 8;
  for (int c0 = c_beg; c0 < c_end; c0 += 64) {
    for (int i = tid; i < AB_N * 8; i += 256) {
      const int row = i >> 3, part = i & 7;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(Y + (size_t)row * ldy + c0 + part * 8);
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
      }
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
        for (int a = 0; a < 4; ++a)
          rv[a] = TRANS ? S[j * (AB_N + 1) + yi * 4 + a] : S[(yi * 4 + a) * (AB_N + 1) + j];
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
          uint4 w;
          w.x = pack_bf16x2(o[a][0], o[a][1]); w.y = pack_bf16x2(o[a][2], o[a][3]);
          w.z = pack_bf16x2(o[a][4], o[a][5]); w.w = pack_bf16x2(o[a][6], o[a][7]);
          *reinterpret_cast<uint4*>(out + (size_t)row * ldo + c0 + yj * 8) = w;
        }
      }
    }
    __syncthreads();
  }
}
// Row-split backward (affinity_rows.cuh): CTA (b, s) owns rows [25 s, 25 s + 25) of dg, dtheta and dphi and forms the
// three 25 x n blocks it needs itself -- (R^T) rows = phi rows . theta^T / n, dS rows = dy rows . g^T / n, (dS^T) rows =
// g rows . dy^T / n -- one product pair per CTA (blockIdx.z), 2.6 M FMA each, where a CTA of the column-split kernel
// below spends 14.1 M (both n x n matrices in every CTA).
__global__ void __launch_bounds__(256)
rs_gcn_affinity_bwd_rows_kernel(const bf16* __restrict__ tpg, const bf16* __restrict__ dy, bf16* __restrict__ dtpg,
                                int n, int C) {
  extern __shared__ __align__(16) float smr[];
  float* Rs = smr;
  float* U = Rs + AR_ROWS * AR_MAXN;
  const int b = blockIdx.x;
  const int r0 = blockIdx.y * AR_ROWS;
  const bf16* th = tpg + (size_t)b * n * 3 * C;
  const bf16* ph = th + C;
  const bf16* gg = th + 2 * C;
  const bf16* dyb = dy + (size_t)b * n * C;
  bf16* dth = dtpg + (size_t)b * n * 3 * C;
  const float inv = 1.0f / (float)n;
  const int ldo = 3 * C;
  auto store_to = [&](bf16* outp) {
    return [=](int row, int col, const float* o) {
      *reinterpret_cast<uint2*>(outp + (size_t)row * ldo + col) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    };
  };
  // blockIdx.z picks one of the three independent products, so that 32 graphs give 384 CTAs instead of 128
  if (blockIdx.z == 0) {
    ar_nt<bf16>(ph, 3 * C, r0, th, 3 * C, n, C, inv, Rs, U);           // (R^T)[j][i] = phi_j . theta_i / n
    ar_sy<bf16>(Rs, dyb, C, r0, n, C, U, store_to(dth + 2 * C));       // dg     = R^T dy
  } else if (blockIdx.z == 1) {
    ar_nt<bf16>(dyb, C, r0, gg, 3 * C, n, C, inv, Rs, U);              // dS[i][j] = dy_i . g_j / n
    ar_sy<bf16>(Rs, ph, 3 * C, r0, n, C, U, store_to(dth));            // dtheta = dS phi
  } else {
    ar_nt<bf16>(gg, 3 * C, r0, dyb, C, n, C, inv, Rs, U);              // (dS^T)[j][i] = g_j . dy_i / n
    ar_sy<bf16>(Rs, th, 3 * C, r0, n, C, U, store_to(dth + C));        // dphi   = dS^T theta
  }
}
__global__ void __launch_bounds__(256)
rs_gcn_affinity_bwd_kernel(const bf16* __restrict__ tpg, const bf16* __restrict__ dy, bf16* __restrict__ dtpg, int n,
                           int C) {
  extern __shared__ float sm[];
  float* R = sm;                                  // [AB_N][AB_N + 1]
  float* dS = R + AB_N * (AB_N + 1);
  float* bufA = dS + AB_N * (AB_N + 1);           // [AB_N][64]
  float* bufB = bufA + AB_N * 64;                 // [AB_N][33]
  const int b = blockIdx.x;
  const bf16* th = tpg + (size_t)b * n * 3 * C;
  const bf16* ph = th + C;
  const bf16* gg = th + 2 * C;
  const bf16* dyb = dy + (size_t)b * n * C;
  bf16* dth = dtpg + (size_t)b * n * 3 * C;
  const float inv = 1.0f / (float)n;
  // gridDim.y CTAs share a graph: each recomputes the two n x n matrices and writes its own slice of the columns
  const int c_per = ((C / 64 + gridDim.y - 1) / gridDim.y) * 64;
  const int c_beg = blockIdx.y * c_per, c_end = min(C, c_beg + c_per);
  aff_nt(th, 3 * C, ph, 3 * C, n, C, inv, R, bufA, bufB);     // R  = theta phi^T / n
  aff_nt(dyb, C, gg, 3 * C, n, C, inv, dS, bufA, bufB);       // dS = dy g^T / n
  aff_sy<true>(R, dyb, C, n, c_beg, c_end, dth + 2 * C, 3 * C, bufA);    // dg     = R^T dy
  aff_sy<false>(dS, ph, 3 * C, n, c_beg, c_end, dth, 3 * C, bufA);       // dtheta = dS phi
  aff_sy<true>(dS, th, 3 * C, n, c_beg, c_end, dth + C, 3 * C, bufA);    // dphi   = dS^T theta
}

// --------------------------------------- l2norm over the node axis + node mean ---------------------------------------
// z fp32 [B, n, D] -> out[b, d] = mean_r z[b,r,d] / sqrt(sum_r z[b,r,d]^2)   (GraphModel.py:74-79,200-204; no eps);
// out has row stride ldo (it is the middle third of the [B, 3D] feature row); inv_s[b,d] = 1 / sqrt(sum z^2)
__global__ void l2norm_mean_fwd_kernel(const float* __restrict__ z, float* __restrict__ out, int ldo,
                                       float* __restrict__ inv_s, int n, int D) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float* zp = z + (size_t)b * n * D + d;
  float s = 0.f, sq = 0.f;
  for (int r = 0; r < n; ++r) {
    const float v = zp[(size_t)r * D];
    s += v;
    sq += v * v;
  }
  const float is = rsqrtf(sq);
  out[(size_t)b * ldo + d] = s * is / (float)n;
  inv_s[(size_t)b * D + d] = is;
}
// y_r = z_r inv_s;  dy_r = dm / n;  dz_r = (dy_r - y_r sum_m dy_m y_m) inv_s = dm inv_s / n (1 - y_r sum_m y_m)
__global__ void l2norm_mean_bwd_kernel(const float* __restrict__ z, const float* __restrict__ inv_s,
                                       const float* __restrict__ dm, int ldm, float* __restrict__ dz32,
                                       bf16* __restrict__ dzb, int n, int D) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float* zp = z + (size_t)b * n * D + d;
  const float is = inv_s[(size_t)b * D + d];
  float sy = 0.f;
  for (int r = 0; r < n; ++r) sy += zp[(size_t)r * D] * is;
  const float k = dm[(size_t)b * ldm + d] * is / (float)n;
  for (int r = 0; r < n; ++r) {
    const float v = k * (1.0f - zp[(size_t)r * D] * is * sy);
    const size_t o = ((size_t)b * n + r) * D + d;
    dz32[o] = v;
    if (dzb) dzb[o] = __float2bfloat16(v);
  }
}

// ------------------------------------------- cross entropy (mean) fwd + bwd -------------------------------------------
// logits fp32 [B, C<=8], labels int64; loss_sum += sum_b CE_b * scale; dlogits = (softmax - onehot) * scale
// One block walks the batch: the per-row losses are summed in a fixed order (shared-memory tree), so the reported loss
// is bit-reproducible (a float atomicAdd per row was not).
__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ loss_sum,
               float* __restrict__ dlogits, int B, int C, float scale) {
  __shared__ float part[8][1];
  float acc[1] = {0.f};
  for (int b = threadIdx.x; b < B; b += 256) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, logits[(size_t)b * C + c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(logits[(size_t)b * C + c] - m);
    const int y = (int)labels[b];
    const float lse = m + logf(s);
    acc[0] += (lse - logits[(size_t)b * C + y]) * scale;
    for (int c = 0; c < C; ++c)
      dlogits[(size_t)b * C + c] = (expf(logits[(size_t)b * C + c] - lse) - (c == y ? 1.f : 0.f)) * scale;
  }
  const float t = block_reduce_fixed<1>(acc, part);
  if (threadIdx.x == 0) loss_sum[0] += t;
}

// small fp32 linear backward (heads): dX[M,K] = dY[M,N] W[N,K];  dW[N,K] += dY^T X;  db[N] += sum dY.   N <= 8
__global__ void __launch_bounds__(256)
linear_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy,
                        float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int M, int N, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) {
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int m = 0; m < M; ++m) acc += dy[(size_t)m * N + n] * x[(size_t)m * K + k];
      dw[(size_t)n * K + k] += acc;
    }
    if (dx)
      for (int m = 0; m < M; ++m) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc += dy[(size_t)m * N + n] * w[(size_t)n * K + k];
        dx[(size_t)m * K + k] = acc;
      }
  }
  if (blockIdx.x == 0 && threadIdx.x < N) {
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc += dy[(size_t)m * N + threadIdx.x];
    db[threadIdx.x] += acc;
  }
}

// ------------------------------------------------ optimiser ------------------------------------------------
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += v * v;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;       // per-block partial: the final sum runs in a fixed order
}
// out[0] += sum of the partials, one block, fixed order: identical gradients give a bit-identical norm on every rank
// (a float atomicAdd tree would let data-parallel replicas drift apart through the clip factor)
__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] += s;
}
// AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments) on a flat fp32 parameter
// buffer; gradients are first scaled by min(1, max_norm / (||g|| + 1e-6)) (clip_grad_norm_, utils_multi.py:233).
// seg_end[k] / seg_wd[k]: exclusive end offset and weight decay of parameter segment k (no-decay group = 0).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, const long long* __restrict__ seg_end, const float* __restrict__ seg_wd, int nseg,
             const float* __restrict__ gnorm_sq, float max_norm, float lr, float beta1, float beta2, float eps,
             float bc1, float bc2) {
  // four elements per thread (128-bit accesses; the flat buffers are 256-byte aligned), ONE segment search per thread:
  // one element per thread with a nine-step dependent search each ran the 231 M-parameter update at half the HBM rate
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  float clip = 1.0f;
  if (max_norm > 0.f) {
    const float gn = sqrtf(*gnorm_sq);
    clip = fminf(1.0f, max_norm / (gn + 1e-6f));
  }
  auto seg_of = [&](long long i) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (i < seg_end[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  int sg = seg_of(i0);
  long long end = seg_end[sg];
  float wd = seg_wd[sg];
  auto update = [&](long long i, float& pi, float gi, float& mi, float& vi) {
    if (i >= end && sg < nseg - 1) {                  // crossed into the next segment (rare)
      sg = seg_of(i);
      end = seg_end[sg];
      wd = seg_wd[sg];
    }
    // (the statements of the one-element-per-thread form, kept verbatim: the trainer tests pin the trajectory)
    gi *= clip;
    mi = beta1 * mi + (1.f - beta1) * gi;
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    pi = pi * (1.f - lr * wd);
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  };
  const bool vec = i0 + 3 < n && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                  reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    float4 p4 = *reinterpret_cast<float4*>(p + i0), m4 = *reinterpret_cast<float4*>(m + i0);
    float4 v4 = *reinterpret_cast<float4*>(v + i0);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + i0));
    update(i0, p4.x, g4.x, m4.x, v4.x);
    update(i0 + 1, p4.y, g4.y, m4.y, v4.y);
    update(i0 + 2, p4.z, g4.z, m4.z, v4.z);
    update(i0 + 3, p4.w, g4.w, m4.w, v4.w);
    *reinterpret_cast<float4*>(p + i0) = p4;
    *reinterpret_cast<float4*>(m + i0) = m4;
    *reinterpret_cast<float4*>(v + i0) = v4;
  } else {
    for (long long i = i0; i < n && i < i0 + 4; ++i) {
      float pi = p[i], mi = m[i], vi = v[i];
      update(i, pi, g[i], mi, vi);
      p[i] = pi; m[i] = mi; v[i] = vi;
    }
  }
}

}  // namespace mv

using namespace mv;

#define GRID1(n, t) (unsigned)(((n) + (t) - 1) / (t))

extern "C" int mvuld_transpose_bf16_batched(const long long* desc, const int* tile_end, int nmat, int total_tiles,
                                            cudaStream_t stream) {
  MV_CHECK_ARG(desc && tile_end && nmat >= 1, "transpose_batched: null table");
  if (total_tiles <= 0) return 0;
  transpose_bf16_batched_kernel<<<total_tiles, dim3(32, 8), 0, stream>>>(desc, tile_end, nmat);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_transpose_bf16(const void* in, int ldi, void* out, int R, int C, int ldo, cudaStream_t stream) {
  MV_CHECK_ARG(ldo >= R && ldi >= C, "transpose: ldo < R or ldi < C");
  if (R <= 0 || C <= 0) return 0;
  dim3 grid((C + 31) / 32, (ldo + 31) / 32), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const bf16*>(in), ldi, reinterpret_cast<bf16*>(out), R, C, ldo);
  MV_LAUNCH_OK();
  return 0;
}
// row slabs (= rows of the fp32 [slabs, C] partials workspace) mvuld_colsum uses for R rows: 8 blocks per SM (the
// kernel is a pure stream: it needs the SM's full thread count to keep enough loads in flight), at least four passes of
// the block's row lanes per slab
extern "C" int mvuld_colsum_slabs(int R, int C) {
  const int units = (C + 7) / 8;
  const int segs = (units + 255) / 256;
  const int rpp = 256 / (units < 256 ? units : 256);
  int slabs = (8 * num_sms() + segs - 1) / segs;
  const int max_by_rows = (R + 4 * rpp - 1) / (4 * rpp);
  if (slabs > max_by_rows) slabs = max_by_rows;
  return slabs < 1 ? 1 : slabs;
}
extern "C" int mvuld_colsum(const void* x, int is_bf16, int ldx, float* out, float* partials, int R, int C,
                            cudaStream_t stream) {
  MV_CHECK_ARG(ldx >= C, "colsum: ldx < C");
  if (R <= 0 || C <= 0) return 0;
  const int slabs = mvuld_colsum_slabs(R, C);
  MV_CHECK_ARG(slabs == 1 || partials != nullptr, "colsum: the [mvuld_colsum_slabs(R, C), C] partials workspace is null");
  const int rps = (R + slabs - 1) / slabs;
  dim3 grid(((C + 7) / 8 + 255) / 256, slabs);
  if (is_bf16) colsum_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), ldx, out, partials, R, C, rps);
  else colsum_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), ldx, out, partials, R, C, rps);
  MV_LAUNCH_OK();
  if (slabs > 1) {
    colsum_final_kernel<<<(C + 31) / 32, dim3(32, 32), 0, stream>>>(partials, out, slabs, C);
    MV_LAUNCH_OK();
  }
  return 0;
}
extern "C" int mvuld_elu_bwd(const void* dy, const void* y, void* dx, long long n, int is_f32, unsigned long long seed,
                             float p, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (is_f32) elu_bwd_f32_kernel<<<GRID1(n, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(y), reinterpret_cast<float*>(dx), n);
  else elu_bwd_kernel<<<GRID1(n, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(y), reinterpret_cast<bf16*>(dx), n, seed, p);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_dropout_bf16(const void* x, void* out, long long n, unsigned long long seed, float p,
                                  cudaStream_t stream) {
  MV_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  if (n <= 0) return 0;
  dropout_bf16_kernel<<<GRID1(n, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out), n, seed, p);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_bn_cols_fwd(const float* x, const float* gamma, const float* beta, float eps, const float* res,
                                 int ldr, float* y32, int ldy, void* yb, float* mean, float* rstd, float* run_mean,
                                 float* run_var, float momentum, int R, int C, cudaStream_t stream) {
  MV_CHECK_ARG(R >= 1 && C >= 1, "bn_cols_fwd: empty");
  bn_cols_fwd_kernel<<<(C + BN_CL - 1) / BN_CL, BN_CL * BN_RL, 0, stream>>>(x, gamma, beta, eps, res, ldr, y32, ldy, reinterpret_cast<bf16*>(yb), mean, rstd, run_mean, run_var, momentum, R, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_bn_cols_bwd(const float* x, const float* dy, int ldy, const float* gamma, const float* mean,
                                 const float* rstd, float* dx32, void* dxb, float* dgamma, float* dbeta, int R, int C,
                                 cudaStream_t stream) {
  bn_cols_bwd_kernel<<<(C + BN_CL - 1) / BN_CL, BN_CL * BN_RL, 0, stream>>>(x, dy, ldy, gamma, mean, rstd, dx32, reinterpret_cast<bf16*>(dxb), dgamma, dbeta, R, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_bn_slot_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                                 float* rstd, float* run_mean, float* run_var, float momentum, int B, int n, int F,
                                 cudaStream_t stream) {
  bn_slot_fwd_kernel<<<n, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), gamma, beta, eps, reinterpret_cast<bf16*>(y), mean, rstd, run_mean, run_var, momentum, B, n, F);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_bn_slot_bwd(const void* x, const void* dy, const float* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, int B, int n, int F,
                                 cudaStream_t stream) {
  bn_slot_bwd_kernel<<<n, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy), gamma, mean, rstd, reinterpret_cast<bf16*>(dx), dgamma, dbeta, B, n, F);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_ln_rows_bwd_blocks(int M);
extern "C" int mvuld_ln_rows_bwd(const void* y, const float* shortcut, const float* gamma, const float* dout, void* dv_bf16,
                                 float* dv_f32, float* dgamma, float* dbeta, float* dbias, float* partials, int M, int C,
                                 float eps, int mode, cudaStream_t stream) {
  MV_CHECK_ARG(partials != nullptr, "ln_rows_bwd: the [mvuld_ln_rows_bwd_blocks(M), 3, C] partials workspace is null");
  MV_CHECK_ARG(C % 8 == 0 && C <= 1024, "ln_rows_bwd: C=%d must be a multiple of 8 and <= 1024", C);
  MV_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || shortcut), "ln_rows_bwd: mode %d (mode 2 needs the shortcut)", mode);
  MV_CHECK_ARG(dgamma && dbeta && (dv_bf16 || dv_f32), "ln_rows_bwd: dgamma / dbeta and one of the dv outputs are required");
  if (M <= 0) return 0;
  const bf16* yp = reinterpret_cast<const bf16*>(y);
  bf16* dvp = reinterpret_cast<bf16*>(dv_bf16);
  int grid = 0;
#define MV_LN_BWD(U, LPR, NR, DB)                                                                                      \
  do {                                                                                                                 \
    constexpr int RPW = (32 / (LPR)) * (NR);                                                                           \
    const size_t smem = (size_t)8 * ((DB) ? 3 : 2) * (32 / (LPR)) * C * sizeof(float);                                 \
    auto kern = ln_rows_bwd_kernel<U, LPR, NR, DB>;                                                                    \
    static unsigned long long attr_set = 0;                                                                            \
    if (first_use_on_current_device(&attr_set))                                                                        \
      MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * (32 / (LPR)) * 1024 * 4)); \
    /* exactly the blocks the GPU holds at once: a second wave would serialise one more load -> reduce -> store chain */ \
    int per_sm = 1;                                                                                                    \
    MV_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));                               \
    per_sm = std::max(1, std::min(per_sm, 6));                                                                         \
    grid = std::min((M + 8 * RPW - 1) / (8 * RPW), std::min(per_sm * num_sms(), mvuld_ln_rows_bwd_blocks(M)));         \
    kern<<<grid, 256, smem, stream>>>(yp, shortcut, gamma, dout, dvp, dv_f32, partials, M, C, eps, mode);              \
  } while (0)
  if (dbias) {        // one row per warp pass where the third accumulator would cost a resident block
    if (C <= 128) MV_LN_BWD(1, 16, 2, true);
    else if (C <= 256) MV_LN_BWD(1, 32, 2, true);
    else if (C <= 512) MV_LN_BWD(2, 32, 1, true);
    else if (C <= 768) MV_LN_BWD(3, 32, 1, true);
    else MV_LN_BWD(4, 32, 1, true);
  } else {
    if (C <= 128) MV_LN_BWD(1, 16, 2, false);
    else if (C <= 256) MV_LN_BWD(1, 32, 2, false);
    else if (C <= 512) MV_LN_BWD(2, 32, 2, false);
    else if (C <= 768) MV_LN_BWD(3, 32, 1, false);
    else MV_LN_BWD(4, 32, 1, false);
  }
#undef MV_LN_BWD
  MV_LAUNCH_OK();
  ln_rows_bwd_final_kernel<<<(C + 31) / 32, dim3(32, 32), 0, stream>>>(partials, dgamma, dbeta, dbias, grid, C);
  MV_LAUNCH_OK();
  return 0;
}
// rows of the partials workspace mvuld_ln_rows_bwd needs for M rows ([blocks, 3, C] floats)
extern "C" int mvuld_ln_rows_bwd_blocks(int M) { return std::min((M + 7) / 8, 6 * num_sms()); }   // upper bound: the launch uses the resident blocks (<= 6 per SM)
// GELU backward + the column sums of its result in one pass (dpre = dh GELU'(pre) is the gradient of fc1's output: its
// column sums are fc1's bias gradient).  Thread = (row lane, 8-column unit) as in colsum_kernel; four rows in flight.
namespace mv {
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const bf16* __restrict__ pre, const bf16* __restrict__ dh, bf16* __restrict__ dpre,
                       float* __restrict__ out, float* __restrict__ partials, int R, int C, int rows_per_slab) {
  __shared__ float red[256 * 8];
  const int units_total = C >> 3;
  const int u0 = blockIdx.x * 256;
  const int upr = min(256, units_total - u0);
  const int rpp = 256 / upr;
  const int tr = threadIdx.x / upr, tu = threadIdx.x - tr * upr;
  const bool active = tr < rpp;
  const int c0 = (u0 + tu) * 8;
  const int rbeg = blockIdx.y * rows_per_slab, rend = min(R, rbeg + rows_per_slab);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto one = [&](const uint4& xr, const uint4& dr, size_t off) {
    const uint32_t xs[4] = {xr.x, xr.y, xr.z, xr.w}, ds[4] = {dr.x, dr.y, dr.z, dr.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g0 = bf16_lo(ds[q]) * gelu_erf_grad(bf16_lo(xs[q])), g1 = bf16_hi(ds[q]) * gelu_erf_grad(bf16_hi(xs[q]));
      acc[2 * q] += g0;
      acc[2 * q + 1] += g1;
      o[q] = pack_bf16x2(g0, g1);
    }
    *reinterpret_cast<uint4*>(dpre + off) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  if (active) {
    int r = rbeg + tr;
    for (; r + 3 * rpp < rend; r += 4 * rpp) {         // four rows (eight 16-byte loads) in flight per thread
      uint4 x[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t o = (size_t)(r + j * rpp) * C + c0;
        x[j] = __ldg(reinterpret_cast<const uint4*>(pre + o));
        d[j] = __ldg(reinterpret_cast<const uint4*>(dh + o));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) one(x[j], d[j], (size_t)(r + j * rpp) * C + c0);
    }
    for (; r < rend; r += rpp) {
      const size_t o0 = (size_t)r * C + c0;
      one(__ldg(reinterpret_cast<const uint4*>(pre + o0)), __ldg(reinterpret_cast<const uint4*>(dh + o0)), o0);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  if (active && tr == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
      for (int q = 0; q < rpp; ++q) t += red[(q * upr + tu) * 8 + k];
      if (gridDim.y == 1) out[c0 + k] += t;
      else partials[(size_t)blockIdx.y * C + c0 + k] = t;
    }
  }
}
}  // namespace mv
extern "C" int mvuld_gelu_bwd_colsum(const void* pre, const void* dh, void* dpre, float* dbias, float* partials, int M,
                                     int C, cudaStream_t stream) {
  MV_CHECK_ARG(C % 8 == 0 && pre && dh && dpre && dbias, "gelu_bwd_colsum: C %% 8 and non-null pointers");
  if (M <= 0) return 0;
  int slabs = mvuld_colsum_slabs(M, C);
  MV_CHECK_ARG(slabs == 1 || partials != nullptr, "gelu_bwd_colsum: the [mvuld_colsum_slabs(M, C), C] partials workspace is null");
  // no more blocks than are resident at once (40 registers: 6 per SM, not the 8 the slab count assumes -- a second,
  // one-third-full wave cost a third of the kernel's time); the workspace bound stays mvuld_colsum_slabs
  {
    int per_sm = 1;
    MV_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gelu_bwd_colsum_kernel, 256, 0));
    const int segs = (C / 8 + 255) / 256;
    const int cap = std::max(1, per_sm * num_sms() / segs);
    if (slabs > cap) slabs = cap;
  }
  const int rps = (M + slabs - 1) / slabs;
  dim3 grid((C / 8 + 255) / 256, slabs);
  gelu_bwd_colsum_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(pre), reinterpret_cast<const bf16*>(dh),
                                                   reinterpret_cast<bf16*>(dpre), dbias, partials, M, C, rps);
  MV_LAUNCH_OK();
  if (slabs > 1) {
    colsum_final_kernel<<<(C + 31) / 32, dim3(32, 32), 0, stream>>>(partials, dbias, slabs, C);
    MV_LAUNCH_OK();
  }
  return 0;
}
extern "C" int mvuld_gelu_bwd(const void* pre, const void* dh, void* dpre, long long n, cudaStream_t stream) {
  MV_CHECK_ARG(n % 8 == 0, "gelu_bwd: n %% 8");
  if (n <= 0) return 0;
  gelu_bwd_kernel<<<GRID1(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(pre), reinterpret_cast<const bf16*>(dh), reinterpret_cast<bf16*>(dpre), n / 8);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_elu_bwd_rows(const float* dy, int ldy, const float* y, int ldyy, void* dx, int ldx, int R, int C,
                                  cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  elu_bwd_rows_kernel<<<GRID1((long long)R * C, 256), 256, 0, stream>>>(dy, ldy, y, ldyy, reinterpret_cast<bf16*>(dx), ldx, R, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_pos_slot_stats(const float* pos, const long long* offsets, const float* gamma, const float* beta,
                                    float eps, float* scale, float* shift, float* mean, float* rstd, float* run_mean,
                                    float* run_var, float momentum, int B, int n, cudaStream_t stream) {
  MV_CHECK_ARG(B >= 1 && n >= 1, "pos_slot_stats: empty");
  pos_slot_stats_kernel<<<n, 128, 0, stream>>>(pos, offsets, gamma, beta, eps, scale, shift, mean, rstd, run_mean, run_var, momentum, B, n);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_pos_branch_bwd(const float* pos, const long long* offsets, const float* mean, const float* rstd,
                                    const float* gamma, const float* beta, const float* w, const void* dpre, float* dw,
                                    float* db, float* dgamma, float* dbeta, float* partials, int B, int n, int OUT,
                                    int ld, int col0, cudaStream_t stream) {
  MV_CHECK_ARG(OUT == 32, "pos_branch_bwd: fc_bbox has 32 outputs (got %d)", OUT);
  MV_CHECK_ARG(partials != nullptr, "pos_branch_bwd: the [n, 160] partials workspace is null");
  if (B <= 0) return 0;
  pos_branch_bwd_kernel<<<n, 128, 0, stream>>>(pos, offsets, mean, rstd, gamma, beta, w, reinterpret_cast<const bf16*>(dpre), partials, dgamma, dbeta, B, n, ld, col0);
  MV_LAUNCH_OK();
  pos_branch_bwd_final_kernel<<<1, 160, 0, stream>>>(partials, dw, db, n);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_unbatch_pad_bwd(const void* dhp, const long long* offsets, void* dh, int B, int max_node, int F,
                                     cudaStream_t stream) {
  MV_CHECK_ARG(F % 8 == 0, "unbatch_pad_bwd: F %% 8");
  if (B <= 0) return 0;
  dim3 grid(64, B);
  unbatch_pad_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dhp), offsets, reinterpret_cast<bf16*>(dh), B, max_node, F);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_gat_bwd(const void* z, const void* dout, const float* el, const float* er, const int* indptr,
                             const int* idx_src, const int* out_indptr, const int* out_dst, const int* pos_in,
                             const float* attn_l, const float* attn_r, float* alpha_e, float* ds_e, float* del,
                             float* der, void* dz, float* dattn_l, float* dattn_r, int N, int H, int F, float slope,
                             cudaStream_t stream) {
  MV_CHECK_ARG(H == 4 && F % 256 == 0, "gat_bwd: H == 4 and F %% 256 == 0 (H=%d F=%d)", H, F);
  if (N <= 0) return 0;
  gat_bwd_dst_kernel<4><<<(N + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const bf16*>(z), reinterpret_cast<const bf16*>(dout), el, er, indptr, idx_src, alpha_e, ds_e, der, N, F, slope);
  MV_LAUNCH_OK();
  gat_bwd_src_kernel<4><<<(N + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const bf16*>(dout), alpha_e, ds_e, der, out_indptr, out_dst, pos_in, attn_l, attn_r, reinterpret_cast<bf16*>(dz), del, N, F);
  MV_LAUNCH_OK();
  gat_attn_grad_kernel<<<H * F / 8, 256, 0, stream>>>(reinterpret_cast<const bf16*>(z), del, der, dattn_l, dattn_r, N, H, F);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_rs_gcn_affinity_bwd(const void* tpg, const void* dy, void* dtpg, int B, int n, int C,
                                         cudaStream_t stream) {
  MV_CHECK_ARG(n >= 1 && n <= AB_N && C % 64 == 0, "rs_gcn_affinity_bwd: n in [1, %d], C %% 64", AB_N);
  if (B <= 0) return 0;
  if (C % AR_GCH == 0) {
    MV_CUDA_OK(cudaFuncSetAttribute(rs_gcn_affinity_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM_BYTES));
    rs_gcn_affinity_bwd_rows_kernel<<<dim3(B, (n + AR_ROWS - 1) / AR_ROWS, 3), 256, AR_SMEM_BYTES, stream>>>(reinterpret_cast<const bf16*>(tpg), reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dtpg), n, C);
    MV_LAUNCH_OK();
    return 0;
  }
  const int smem = (2 * AB_N * (AB_N + 1) + AB_N * 64 + AB_N * 33) * sizeof(float);
  MV_CUDA_OK(cudaFuncSetAttribute(rs_gcn_affinity_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rs_gcn_affinity_bwd_kernel<<<dim3(B, B >= 148 ? 2 : 4), 256, smem, stream>>>(reinterpret_cast<const bf16*>(tpg), reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dtpg), n, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_l2norm_mean_fwd(const float* z, float* out, int ldo, float* inv_s, int B, int n, int D,
                                     cudaStream_t stream) {
  if (B <= 0) return 0;
  dim3 grid((D + 127) / 128, B);
  l2norm_mean_fwd_kernel<<<grid, 128, 0, stream>>>(z, out, ldo, inv_s, n, D);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_l2norm_mean_bwd(const float* z, const float* inv_s, const float* dm, int ldm, float* dz32,
                                     void* dzb, int B, int n, int D, cudaStream_t stream) {
  if (B <= 0) return 0;
  dim3 grid((D + 127) / 128, B);
  l2norm_mean_bwd_kernel<<<grid, 128, 0, stream>>>(z, inv_s, dm, ldm, dz32, reinterpret_cast<bf16*>(dzb), n, D);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_ce_loss(const float* logits, const long long* labels, float* loss_sum, float* dlogits, int B,
                             int C, float scale, cudaStream_t stream) {
  MV_CHECK_ARG(C >= 1 && C <= 64, "ce_loss: C in [1, 64]");
  if (B <= 0) return 0;
  ce_loss_kernel<<<1, 256, 0, stream>>>(logits, labels, loss_sum, dlogits, B, C, scale);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_linear_small_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                                      int M, int N, int K, cudaStream_t stream) {
  MV_CHECK_ARG(N >= 1 && N <= 8, "linear_small_bwd: N in [1, 8]");
  if (M <= 0) return 0;
  linear_small_bwd_kernel<<<GRID1(K, 256), 256, 0, stream>>>(x, w, dy, dx, dw, db, M, N, K);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_sumsq_f32(const float* x, long long n, float* partials, float* out, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, partials);
  MV_LAUNCH_OK();
  sumsq_final_kernel<<<1, 256, 0, stream>>>(partials, (int)blocks, out);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_adamw(float* p, const float* g, float* m, float* v, long long n, const long long* seg_end,
                           const float* seg_wd, int nseg, const float* gnorm_sq, float max_norm, float lr, float beta1,
                           float beta2, float eps, int step, cudaStream_t stream) {
  MV_CHECK_ARG(nseg >= 1 && step >= 1, "adamw: nseg >= 1 and step >= 1");
  if (n <= 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<GRID1((n + 3) / 4, 256), 256, 0, stream>>>(p, g, m, v, n, seg_end, seg_wd, nseg, gnorm_sq, max_norm, lr, beta1, beta2, eps, bc1, bc2);
  MV_LAUNCH_OK();
  return 0;
}
