// Kernels only the fusion ablation classes need:
//   * node_linear4        per-node ELU(Linear(4, OUT)) of the box feature `pos_emb`
//                         (GraphModel.py:791 _GATPOS, :1132 _NOGAT3, :1242 _NOGAT4)
//   * unbatch_pad_bn_elu  ELU(bn_gat(pad(h))) with no projection behind it (GraphModel.py:928 _011)
//   * gru_sequence        nn.GRU(512, 512, 1, batch_first=True) over the padded node axis, last hidden state
//                         (myModels.py:324,385-387: projection_layer == 'gru')
//   * gate_fusion         softmax(tanh(x * h), dim=1) * h                      (myModels.py:407-413: fusion == 'attention')
// HBM / latency bound helper kernels: fp32 arithmetic, fixed summation order.
#include <cooperative_groups.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

__global__ void node_linear4_kernel(const float* __restrict__ pos, const float* __restrict__ w,
                                    const float* __restrict__ bias, bf16* __restrict__ out, int N, int OUT, int ld,
                                    int col0) {
  const long long total = (long long)N * OUT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % OUT);
    const long long r = i / OUT;
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + r);
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + o);
    float acc = __ldg(bias + o);
    acc = fmaf(p.x, ww.x, acc);
    acc = fmaf(p.y, ww.y, acc);
    acc = fmaf(p.z, ww.z, acc);
    acc = fmaf(p.w, ww.w, acc);
    out[r * ld + col0 + o] = __float2bfloat16(elu1(acc));
  }
}

__global__ void unbatch_pad_bn_elu_kernel(const bf16* __restrict__ feat, const long long* __restrict__ off,
                                          const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                          float* __restrict__ z32, bf16* __restrict__ zb, int B, int max_node, int F) {
  const int units = F >> 3;
  const long long total = (long long)B * max_node * units;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    const long long br = i / units;
    const int r = (int)(br % max_node);
    const int b = (int)(br / max_node);
    const long long beg = off[b], n = off[b + 1] - beg;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(feat + (beg + r) * F) + u);
      f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
      f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
    const float sc = __ldg(bn_scale + r), sh = __ldg(bn_shift + r);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = elu1(f[k] * sc + sh);
    float4* o32 = reinterpret_cast<float4*>(z32) + i * 2;
    o32[0] = make_float4(f[0], f[1], f[2], f[3]);
    o32[1] = make_float4(f[4], f[5], f[6], f[7]);
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    reinterpret_cast<uint4*>(zb)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GRU over a sequence.  gi fp32 [B, T, 3H] = x_t W_ih^T + b_ih for every step (one GEMM, gate order r | z | n as in
// torch.nn.GRU), w_hh fp32 [3H, H], b_hh fp32 [3H]; h_0 = 0; h_out fp32 [B, H] = h_T.
//   r = sigmoid(gi_r + W_hr h + b_hr), z = sigmoid(gi_z + W_hz h + b_hz), n = tanh(gi_n + r (W_hn h + b_hn)),
//   h' = (1 - z) n + z h
// Cooperative grid of H / GRU_HC CTAs; CTA c owns hidden units [c GRU_HC, (c + 1) GRU_HC) and keeps the 3 GRU_HC rows
// of W_hh it needs in shared memory for the whole sequence; the state ping-pongs between two fp32 [B, H] buffers in
// global memory (read through L2: ld.global.cg) with one grid barrier per step.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GRU_HC = 8;
constexpr int GRU_THREADS = 256;

__device__ __forceinline__ void gru_grid_barrier(unsigned int* counter, unsigned int target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
      if (clock64() - t0 > 4000000000ll) __trap();           // ~2 s watchdog: a protocol bug becomes a CUDA error
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(GRU_THREADS)
gru_sequence_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                    float* __restrict__ hbuf /* [2, B, H], zero on entry */, float* __restrict__ h_out,
                    unsigned int* __restrict__ counter /* zero on entry */, int B, int T, int H) {
  extern __shared__ float gru_w[];                               // [3][GRU_HC][H + 4]: the 4 rows a warp reads at once
  const int j0 = blockIdx.x * GRU_HC;                            // (one per unit pair) fall into different banks
  const int HP = H + 4;
  for (int i = threadIdx.x; i < 3 * GRU_HC * H; i += GRU_THREADS) {
    const int g = i / (GRU_HC * H), rem = i % (GRU_HC * H);
    gru_w[(g * GRU_HC + rem / H) * HP + rem % H] = __ldg(w_hh + ((size_t)g * H + j0 + rem / H) * H + rem % H);
  }
  __syncthreads();
  // thread -> (batch row b, pair of hidden units jp): 4 threads per batch row, 2 units each, 3 gates per unit
  const int jp = threadIdx.x & 3;
  const int H4 = H >> 2;
  float bh[3][2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int u = 0; u < 2; ++u) bh[g][u] = __ldg(b_hh + g * H + j0 + jp * 2 + u);
  for (int t = 0; t < T; ++t) {
    const float* hcur = hbuf + (size_t)(t & 1) * B * H;
    float* hnext = hbuf + (size_t)((t + 1) & 1) * B * H;
    for (int b = threadIdx.x >> 2; b < B; b += GRU_THREADS / 4) {
      float acc[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      const float4* hrow = reinterpret_cast<const float4*>(hcur + (size_t)b * H);
      for (int k = 0; k < H4; ++k) {
        const float4 hv = __ldcg(hrow + k);
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 wv = *reinterpret_cast<const float4*>(gru_w + ((g * GRU_HC + jp * 2 + u) * HP + k * 4));
            acc[g][u] = fmaf(hv.x, wv.x, acc[g][u]);
            acc[g][u] = fmaf(hv.y, wv.y, acc[g][u]);
            acc[g][u] = fmaf(hv.z, wv.z, acc[g][u]);
            acc[g][u] = fmaf(hv.w, wv.w, acc[g][u]);
          }
      }
      const float* gir = gi + ((size_t)b * T + t) * 3 * H + j0 + jp * 2;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float r = 1.f / (1.f + __expf(-(__ldg(gir + u) + acc[0][u] + bh[0][u])));
        const float z = 1.f / (1.f + __expf(-(__ldg(gir + H + u) + acc[1][u] + bh[1][u])));
        const float n = tanhf(__ldg(gir + 2 * H + u) + r * (acc[2][u] + bh[2][u]));
        const float hp = __ldcg(hcur + (size_t)b * H + j0 + jp * 2 + u);
        const float hn = (1.f - z) * n + z * hp;
        hnext[(size_t)b * H + j0 + jp * 2 + u] = hn;
        if (t == T - 1) h_out[(size_t)b * H + j0 + jp * 2 + u] = hn;
      }
    }
    if (t + 1 < T) gru_grid_barrier(counter, (unsigned int)(t + 1) * gridDim.x);
  }
}

// mode 0: out[b, col0 + c] = softmax_c(tanh(x[b, c] h[b, c])) h[b, c]  (myModels.py:409-413, fusion 'attention');
// mode 1: out[b, col0 + c] = x[b, c] h[b, c]                               (myModels.py:419, fusion 'dot').
// One block per row, C <= 1024.
__global__ void gate_fusion_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ out,
                                   int C, int ld, int col0, int mode) {
  __shared__ float red[32];
  const int b = blockIdx.x, c = threadIdx.x;
  const float hv = c < C ? h[(size_t)b * C + c] : 0.f;
  if (mode == 1) {
    if (c < C) out[(size_t)b * ld + col0 + c] = x[(size_t)b * C + c] * hv;
    return;
  }
  const float v = c < C ? tanhf(x[(size_t)b * C + c] * hv) : -INFINITY;
  float m = v;
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((c & 31) == 0) red[c >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float e = c < C ? expf(v - m) : 0.f;
  float s = e;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((c & 31) == 0) red[c >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];       // fixed order
  if (c < C) out[(size_t)b * ld + col0 + c] = e / s * hv;
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_node_linear4(const float* pos, const float* w, const float* bias, void* out_bf16, int N, int OUT,
                                  int ld, int col0, cudaStream_t stream) {
  MV_CHECK_ARG(OUT > 0 && ld >= col0 + OUT, "node_linear4: columns [%d, %d) exceed the row stride %d", col0, col0 + OUT, ld);
  if (N <= 0) return 0;
  node_linear4_kernel<<<grid_for((long long)N * OUT, 256), 256, 0, stream>>>(pos, w, bias,
                                                                             reinterpret_cast<bf16*>(out_bf16), N, OUT,
                                                                             ld, col0);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_unbatch_pad_bn_elu(const void* feat, const long long* offsets, const float* bn_scale,
                                        const float* bn_shift, float* z32, void* zb, int B, int max_node, int F,
                                        cudaStream_t stream) {
  MV_CHECK_ARG(F % 8 == 0, "unbatch_pad_bn_elu: F %% 8");
  if (B <= 0) return 0;
  unbatch_pad_bn_elu_kernel<<<grid_for((long long)B * max_node * (F / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(feat), offsets, bn_scale, bn_shift, z32, reinterpret_cast<bf16*>(zb), B, max_node, F);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" long long mvuld_gru_sequence_workspace(int B, int H) { return 2ll * B * H * 4 + 256; }

extern "C" int mvuld_gru_sequence(const float* gi, const float* w_hh, const float* b_hh, float* h_out, void* workspace,
                                  int B, int T, int H, cudaStream_t stream) {
  MV_CHECK_ARG(H % (4 * GRU_HC) == 0, "gru_sequence: hidden size %d must be a multiple of %d", H, 4 * GRU_HC);
  MV_CHECK_ARG(T >= 1 && B >= 1, "gru_sequence: empty sequence or batch");
  const int grid = H / GRU_HC;
  const size_t smem = (size_t)3 * GRU_HC * (H + 4) * sizeof(float);
  MV_CHECK_ARG(grid <= num_sms(), "gru_sequence: %d CTAs cannot be co-resident on %d SMs", grid, num_sms());
  MV_CHECK_ARG(smem <= 200 * 1024, "gru_sequence: hidden size %d needs %zu bytes of shared memory", H, smem);
  MV_CUDA_OK(cudaFuncSetAttribute(gru_sequence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ws = mvuld_gru_sequence_workspace(B, H);
  MV_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)ws, stream));
  float* hbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  void* args[] = {(void*)&gi, (void*)&w_hh, (void*)&b_hh, (void*)&hbuf, (void*)&h_out, (void*)&counter,
                  (void*)&B, (void*)&T, (void*)&H};
  // cooperative launch: every CTA resident at once (the per-step grid barrier needs it) or the launch fails
  MV_CUDA_OK(cudaLaunchCooperativeKernel((const void*)gru_sequence_kernel, dim3(grid), dim3(GRU_THREADS), args, smem,
                                         stream));
  return 0;
}

extern "C" int mvuld_gate_fusion(const float* x, const float* h, float* out, int B, int C, int ld, int col0, int mode,
                                 cudaStream_t stream) {
  MV_CHECK_ARG(C >= 1 && C <= 1024 && ld >= col0 + C, "gate_fusion: C %d (<= 1024), row stride %d", C, ld);
  if (B <= 0) return 0;
  gate_fusion_kernel<<<B, (C + 31) / 32 * 32, 0, stream>>>(x, h, out, C, ld, col0, mode);
  MV_LAUNCH_OK();
  return 0;
}
