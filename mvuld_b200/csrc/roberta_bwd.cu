// Row kernels of the RoBERTa (UniXcoder) encoder backward (mvuld/models/unixcoder.py:33-38 under autograd): the masked
// mean over the valid tokens and the three embedding tables.  The dense products run on gemm.cu / gemm_dw.cu, the
// attention on attention_bwd.cu, LayerNorm / GELU on train.cu.  Fixed summation orders: bit-reproducible.
#include "common.cuh"
#include "host_util.h"

namespace mv {

// sentence[b] = sum_{t < len_b} tok[b, t] / len_b  ->  dtok[b, t] = dsent[b] / len_b for t < len_b, 0 otherwise
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dsent, const int* __restrict__ len,
                                       float* __restrict__ dtok, int L, int C) {
  const int row = blockIdx.x;                       // b * L + t
  const int b = row / L, t = row - b * L;
  const int n = len[b];
  const float inv = (t < n && n > 0) ? 1.0f / (float)n : 0.f;
  const float4* src = reinterpret_cast<const float4*>(dsent + (size_t)b * C);
  float4* dst = reinterpret_cast<float4*>(dtok + (size_t)row * C);
  for (int c = threadIdx.x; c < C / 4; c += blockDim.x) {
    const float4 v = __ldg(src + c);
    dst[c] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// Embedding-table gradient (nn.Embedding backward = index_add of the incoming rows): the token rows are grouped by
// table index beforehand (mvuld_csr_from_coo with dst = index: stable, so the rows of one index keep their token order);
// one warp sums the rows of one index in that order.  Rows of `skip` (padding_idx) get no gradient.
__global__ void __launch_bounds__(256)
embed_grad_rows_kernel(const float* __restrict__ d, const int* __restrict__ indptr, const int* __restrict__ rows,
                       float* __restrict__ dtab, int n_index, int C, int skip) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= n_index || v == skip) return;
  const int beg = indptr[v], end = indptr[v + 1];
  if (beg == end) return;
  const int lane = threadIdx.x & 31;
  for (int c = lane * 4; c < C; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = beg; k < end; ++k) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(d + (size_t)rows[k] * C + c));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    float4* dst = reinterpret_cast<float4*>(dtab + (size_t)v * C + c);
    const float4 old = *dst;
    *dst = make_float4(old.x + acc.x, old.y + acc.y, old.z + acc.z, old.w + acc.w);
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_masked_mean_bwd(const float* dsent, const int* len, float* dtok, int B, int L, int C,
                                     cudaStream_t stream) {
  MV_CHECK_ARG(C % 4 == 0, "masked_mean_bwd: C %% 4");
  if (B <= 0) return 0;
  masked_mean_bwd_kernel<<<B * L, 192, 0, stream>>>(dsent, len, dtok, L, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_embed_grad_rows(const float* d, const int* indptr, const int* rows, float* dtab, int n_index, int C,
                                     int skip_index, cudaStream_t stream) {
  MV_CHECK_ARG(C % 4 == 0, "embed_grad_rows: C %% 4");
  if (n_index <= 0) return 0;
  embed_grad_rows_kernel<<<(n_index + 7) / 8, 256, 0, stream>>>(d, indptr, rows, dtab, n_index, C, skip_index);
  MV_LAUNCH_OK();
  return 0;
}
