// Row / reduction kernels of the SwinV2 encoder backward (autograd through swin_transformer_v2.py as trained by
// mvuld/main.py:251-300) around the tcgen05 attention backward (attention_bwd.cu) and the dense products (gemm.cu):
//   swin_bias_grad   G^T of every (window, head) -> gradient of the continuous-position-bias table (:159-164)
//   cpb_mlp_bwd      table gradient -> cpb_mlp weights (16 sigmoid(W2 relu(W1 c + b1)), :98-111,159,163)
//   swin_qkv_bwd     (dq^, dk^, dv) -> d qkv, token major: backward of F.normalize and of the logit scale (:155-158),
//                    window_reverse + inverse cyclic shift folded in (:276-299)
//   gelu_fwd, patch_merge_scatter, patch_im2col: training-mode forward / backward plumbing of Mlp, PatchMerging, PatchEmbed
// Every reduction runs in a fixed order (partials + ordered final sum): gradients are bit-reproducible.
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

// ---------------------------------------------------------------------------------------------------------
// Bias-table gradient.  gt: bf16 [n_win * nH, N, npad] = G^T (key major) of every window instance; the table entry of
// (query (qy, qx), key (ky, kx)) is [(qy - ky + ws - 1), (qx - kx + ws - 1)].  Block (head, ky) reads the WHOLE rows of
// its ws keys (contiguous) for every window instance, sums over the instances in registers, then reduces over kx along
// the diagonals qx - kx = const through shared memory: partial[split][head][ky][qy][dx].  The window instances are split
// over blockIdx.z so that nH * ws blocks of a 4-head stage do not leave SMs idle (stage 0 of a 32-image batch: 112 blocks
// walking 512 windows each ran at 2 TB/s).  The final kernel sums the ws partials of a table row in ky order within a split,
// then the splits in split order: a fixed order.
// ---------------------------------------------------------------------------------------------------------
template <int WS, bool SPLIT>
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const bf16* __restrict__ gt, int n_win, int nH, int npad, float* __restrict__ partial) {
  constexpr int N = WS * WS, SIDE = 2 * WS - 1;
  constexpr int VPR = (N + 7) / 8;                       // 16-byte vectors per key row
  constexpr int ITEMS = WS * VPR;                        // (kx, vector) items of this block's ws keys
  constexpr int IPT = (ITEMS + 255) / 256;
  extern __shared__ float tile[];                        // [WS kx][VPR * 8] sums over the window instances
  const int head = blockIdx.x, ky = blockIdx.y;
  float acc[IPT][8];
#pragma unroll
  for (int t = 0; t < IPT; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
  // Code generation of this loop is fragile (measured on the 16-head stage / the 4-head stage): the plain 0 .. n_win loop
  // 179 us; the same kernel with computed bounds [w_beg, w_end) 237 us / 743 us in five splits; a 0 .. count loop from an
  // offset base pointer 2 323 us.  So the unsplit instantiation keeps the plain loop and only the stages with fewer
  // blocks than SMs (four heads: 1 236 us unsplit) take the split form.
  // ncu (16-head stage, profiles/r2_ncu_biasgrad.md): 174 us, 3.6 TB/s of DRAM reads, 21 % warps active (two blocks per SM:
  // 128 registers and 88 KB of shared memory each), every stall a long scoreboard, 1.51 waves.  Tried and not kept:
  // the sums in thread-owned shared-memory slots instead of 88 registers (54 registers, all loads of a window in flight:
  // 230 us, the LDS / STS round trip per load costs more than the occupancy returns); a block's key columns split over
  // three blocks (32 registers of sums, four resident blocks, 1 344 items instead of 448: 254 us on the 16-head stage,
  // 591 vs 347 us on the 8-head stage) -- more, smaller streams per SM lower the DRAM efficiency here, occupancy is not
  // what limits it.
  int w_beg = 0, w_end = n_win;
  if (SPLIT) {
    const int per = (n_win + (int)gridDim.z - 1) / (int)gridDim.z;
    w_beg = blockIdx.z * per;
    w_end = min(n_win, w_beg + per);
  }
  for (int w = w_beg; w < w_end; ++w) {
    const bf16* base = gt + ((size_t)(w * nH + head) * N + ky * WS) * npad;
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
      const int it = threadIdx.x + t * 256;
      if (it < ITEMS) {
        const int kx = it / VPR, vq = it - kx * VPR;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)kx * npad) + vq);
        acc[t][0] += bf16_lo(v.x); acc[t][1] += bf16_hi(v.x); acc[t][2] += bf16_lo(v.y); acc[t][3] += bf16_hi(v.y);
        acc[t][4] += bf16_lo(v.z); acc[t][5] += bf16_hi(v.z); acc[t][6] += bf16_lo(v.w); acc[t][7] += bf16_hi(v.w);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < IPT; ++t) {
    const int it = threadIdx.x + t * 256;
    if (it < ITEMS) {
      const int kx = it / VPR, vq = it - kx * VPR;
#pragma unroll
      for (int e = 0; e < 8; ++e) tile[kx * (VPR * 8) + vq * 8 + e] = acc[t][e];
    }
  }
  __syncthreads();
  float* out = partial + (((size_t)(SPLIT ? blockIdx.z : 0) * nH + head) * WS + ky) * WS * SIDE;      // [qy][dx]
  for (int o = threadIdx.x; o < WS * SIDE; o += 256) {
    const int qy = o / SIDE, dx = o - qy * SIDE;
    float s = 0.f;
    for (int kx = 0; kx < WS; ++kx) {
      const int qx = dx - (WS - 1) + kx;
      if (qx >= 0 && qx < WS) s += tile[kx * (VPR * 8) + qy * WS + qx];
    }
    out[o] = s;
  }
}
template <int WS>
__global__ void bias_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ dtab, int nH, int splits) {
  constexpr int SIDE = 2 * WS - 1;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;   // (head, dy, dx)
  if (o >= nH * SIDE * SIDE) return;
  const int head = o / (SIDE * SIDE), e = o - head * SIDE * SIDE;
  const int dy = e / SIDE, dx = e - dy * SIDE;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) {
    float t = 0.f;
#pragma unroll
    for (int ky = 0; ky < WS; ++ky) {
      const int qy = dy - (WS - 1) + ky;
      if (qy >= 0 && qy < WS) t += partial[((((size_t)z * nH + head) * WS + ky) * WS + qy) * SIDE + dx];
    }
    s += t;
  }
  dtab[o] += s;
}

// ---------------------------------------------------------------------------------------------------------
// cpb_mlp backward.  table[h][e] = 16 sigmoid(o[e][h]), o = W2 relu(W1 c_e + b1); dtab [nH, T] natural units.
// One block per hidden unit u walks the T table entries; fixed-order block reduction of (nH + 3) sums.
// ---------------------------------------------------------------------------------------------------------
constexpr int CPB_MAXH = 32;
__global__ void __launch_bounds__(256)
cpb_mlp_bwd_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                   const float* __restrict__ tab, const float* __restrict__ dtab, int nH, int ws, int pretrained_ws,
                   float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2) {
  __shared__ float part[8][CPB_MAXH + 3];
  const int u = blockIdx.x;
  const int side = 2 * ws - 1, T = side * side;
  const float denom = (float)((pretrained_ws > 0 ? pretrained_ws : ws) - 1);
  auto coord = [&](int idx) {
    float t = (float)(idx - (ws - 1)) / denom * 8.0f;
    float a = log2f(fabsf(t) + 1.0f) / 3.0f;
    return t > 0.f ? a : (t < 0.f ? -a : 0.f);
  };
  const float wa = w1[2 * u], wb = w1[2 * u + 1], bb = b1[u];
  float acc[CPB_MAXH + 3];
#pragma unroll
  for (int i = 0; i < CPB_MAXH + 3; ++i) acc[i] = 0.f;
  for (int e = threadIdx.x; e < T; e += 256) {
    const float ch = coord(e / side), cw = coord(e % side);
    const float pre = wa * ch + wb * cw + bb;
    const float hid = pre > 0.f ? pre : 0.f;
    float dh = 0.f;
#pragma unroll
    for (int h = 0; h < CPB_MAXH; ++h) {
      if (h < nH) {
        const float tv = tab[(size_t)h * T + e];
        const float d_o = dtab[(size_t)h * T + e] * tv * (1.0f - tv * (1.0f / 16.0f));   // 16 s (1 - s), s = tv / 16
        acc[h] += d_o * hid;
        dh += d_o * w2[h * 512 + u];
      }
    }
    if (pre > 0.f) {
      acc[CPB_MAXH] += dh * ch;
      acc[CPB_MAXH + 1] += dh * cw;
      acc[CPB_MAXH + 2] += dh;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < CPB_MAXH + 3; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) part[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < CPB_MAXH + 3) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i < CPB_MAXH) {
      if (i < nH) dw2[i * 512 + u] += t;
    } else if (i == CPB_MAXH) dw1[2 * u] += t;
    else if (i == CPB_MAXH + 1) dw1[2 * u + 1] += t;
    else db1[u] += t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// d qkv from the attention backward.  One thread per (token, head).  With s = exp(min(logit_scale, ln 100)),
// q~ = q / |q| (unit), q^ = s log2e q~ (stored fp16), k^ = k / |k|:
//   kernel dq = G k^   ->  dL/dq~ = s dq        ->  dq_raw = (g - q~ (q~ . g)) / |q|,  g = s dq
//   kernel dk = G^T q^ ->  dL/dk^ = dk / log2e  ->  dk_raw = (h - k^ (k^ . h)) / |k|,  h = dk / log2e
//   d logit_scale[h] += s (q~ . dq)   (0 where the clamp at ln 100 is active)
// Output bf16 token-major [M, 3C] (columns q | k | v like the qkv weight rows), window_reverse + inverse shift applied.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
swin_qkv_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                    const __half* __restrict__ qh, const __half* __restrict__ kh, const float* __restrict__ rq,
                    const float* __restrict__ rk, const float* __restrict__ qscale, bf16* __restrict__ dqkv,
                    float* __restrict__ ls_partial, int B, int H, int W, int C, int nH, int ws, int shift) {
  __shared__ float sls[256];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * nH;
  float ls = 0.f;
  // (token-major thread order kept here: the window-major order that helps attn_bwd_prep_kernel made this kernel slower,
  // 2.5 -> 3.0 ms per step -- its per-head logit-scale partials then need a serial walk over the block's 256 entries)
  if (idx < total) {
    const int head = (int)(idx % nH);
    const long long row = idx / nH;
    const int HW = H * W;
    const int b = (int)(row / HW);
    const int t = (int)(row - (long long)b * HW);
    int hh = t / W, ww = t - hh * W;
    hh -= shift; if (hh < 0) hh += H;
    ww -= shift; if (ww < 0) ww += W;
    const int nWw = W / ws;
    const int win = (hh / ws) * nWw + (ww / ws);
    const int slot = (hh % ws) * ws + (ww % ws);
    const int nW = (H / ws) * nWw;
    const size_t wrow = (((size_t)b * nW + win) * nH + head) * (size_t)(ws * ws) + slot;
    const float qs = __ldg(qscale + head);                    // s log2e
    const float s_nat = qs * 0.6931471805599453f;
    const float inv_qs = 1.0f / qs;
    bf16* orow = dqkv + (size_t)row * 3 * C + head * 32;
    // ---- q ----
    {
      float g[32], u[32];
      const float4* gp = reinterpret_cast<const float4*>(dq + wrow * 32);
      const uint4* qp = reinterpret_cast<const uint4*>(qh + wrow * 32);     // four 16-byte loads, not 16 half2 loads
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(gp + i);
        g[4 * i] = a.x; g[4 * i + 1] = a.y; g[4 * i + 2] = a.z; g[4 * i + 3] = a.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = __ldg(qp + i);
        const uint32_t h2[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
          u[8 * i + 2 * j] = f.x * inv_qs;
          u[8 * i + 2 * j + 1] = f.y * inv_qs;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) dot += u[i] * g[i];      // q~ . dq
      ls = s_nat * dot;
      const float rn = __ldg(rq + wrow) * s_nat;
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        w[i] = pack_bf16x2(rn * (g[2 * i] - u[2 * i] * dot), rn * (g[2 * i + 1] - u[2 * i + 1] * dot));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(orow)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    // ---- k ----
    {
      float g[32], u[32];
      const float4* gp = reinterpret_cast<const float4*>(dk + wrow * 32);
      const uint4* kp = reinterpret_cast<const uint4*>(kh + wrow * 32);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(gp + i);
        g[4 * i] = a.x; g[4 * i + 1] = a.y; g[4 * i + 2] = a.z; g[4 * i + 3] = a.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = __ldg(kp + i);
        const uint32_t h2[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
          u[8 * i + 2 * j] = f.x;
          u[8 * i + 2 * j + 1] = f.y;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) dot += u[i] * g[i];
      const float rn = __ldg(rk + wrow) * 0.6931471805599453f;
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        w[i] = pack_bf16x2(rn * (g[2 * i] - u[2 * i] * dot), rn * (g[2 * i + 1] - u[2 * i + 1] * dot));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(orow + C)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    // ---- v ----
    {
      const float4* gp = reinterpret_cast<const float4*>(dv + wrow * 32);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(gp + i);
        w[2 * i] = pack_bf16x2(a.x, a.y);
        w[2 * i + 1] = pack_bf16x2(a.z, a.w);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(orow + 2 * C)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
  // logit-scale partial of this block: threads of head h are tid = h, h + nH, ... (256 % nH == 0), summed in order
  sls[threadIdx.x] = ls;
  __syncthreads();
  if (threadIdx.x < nH) {
    float t = 0.f;
    for (int i = threadIdx.x; i < 256; i += nH) t += sls[i];
    ls_partial[(size_t)blockIdx.x * nH + threadIdx.x] = t;
  }
}
__global__ void logit_scale_final_kernel(const float* __restrict__ ls_partial, int nblocks, int nH,
                                         const float* __restrict__ logit_scale, float* __restrict__ dls) {
  __shared__ float part[8][1];
  const int h = blockIdx.x;
  float acc[1] = {0.f};
  for (int b = threadIdx.x; b < nblocks; b += 256) acc[0] += ls_partial[(size_t)b * nH + h];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float v = warp_sum(acc[0]);
  if (lane == 0) part[warp][0] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w][0];
    if (logit_scale[h] < 4.605170185988092f) dls[h] += t;      // torch.clamp(max=ln 100): no gradient past the clamp
  }
}

// exact (erf) GELU, bf16 -> bf16 (training forward keeps the pre-activation; inference fuses GELU into the fc1 epilogue)
__global__ void gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = bf16_lo(in[k]), b = bf16_hi(in[k]);
      gelu_erf2(a, b);
      o[k] = pack_bf16x2(a, b);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// inverse of patch_merge_gather for gradients: dgathered fp32 [B*(H/2)*(W/2), 4C] -> dx fp32 [B*H*W, C]
// (swin_transformer_v2.py:352-359: x0 (0,0) | x1 (1,0) | x2 (0,1) | x3 (1,1) of every 2x2 patch; a permutation)
__global__ void patch_merge_scatter_kernel(const float* __restrict__ dg, float* __restrict__ dx, int B, int H, int W,
                                           int C) {
  const int units = C >> 2;
  const long long total = (long long)B * (H / 2) * (W / 2) * 4 * units;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    long long r = i / units;
    const int s = (int)(r % 4);
    r /= 4;
    const int w2 = (int)(r % (W / 2));
    r /= (W / 2);
    const int h2 = (int)(r % (H / 2));
    const int b = (int)(r / (H / 2));
    const int hh = 2 * h2 + (s & 1), ww = 2 * w2 + (s >> 1);
    reinterpret_cast<float4*>(dx + (((size_t)b * H + hh) * W + ww) * C)[u] = __ldg(reinterpret_cast<const float4*>(dg) + i);
  }
}

// PatchEmbed as a product (training): image fp32 [B, 3, Hi, Wi] -> bf16 [B*(Hi/4)*(Wi/4), 48], tap order (c, kh, kw)
// of proj.weight.view(E, 48) (swin_transformer_v2.py:485-493)
__global__ void patch_im2col_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int Hi, int Wi) {
  const int Hp = Hi / 4, Wp = Wi / 4;
  const long long total = (long long)B * Hp * Wp * 12;           // (token, c, kh): 4 contiguous pixels each
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = (int)(i % 12);
    const long long tok = i / 12;
    const int c = ck >> 2, kh = ck & 3;
    const int wp = (int)(tok % Wp);
    const long long r = tok / Wp;
    const int hp = (int)(r % Hp);
    const int b = (int)(r / Hp);
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * Hi + hp * 4 + kh) * Wi + wp * 4));
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + tok * 48 + ck * 4) = o;
  }
}

}  // namespace mv

using namespace mv;

// Window splits of mvuld_swin_bias_grad.  A block's cost is its window instances plus a fixed epilogue (~11 instances'
// worth, measured: splitting the 224 blocks of an 8-head stage five ways was slower than not splitting); the SM holds `per_sm` blocks (the [ws][ws^2] fp32 tile: 88 KB at ws = 28), so a launch runs in
// ceil(blocks / capacity) waves: take the split count that minimises waves x (instances per split + epilogue).  The
// partial workspace holds splits * nH * ws * ws * (2 ws - 1) floats.
extern "C" int mvuld_swin_bias_grad_splits(int n_win, int nH, int ws) {
  const int smem = ws * ((ws * ws + 7) / 8) * 8 * 4;
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const long long cap = (long long)per_sm * num_sms();
  const long long blocks = (long long)nH * ws;
  int best = 1;
  long long best_cost = -1;
  for (int sp = 1; sp <= 32 && sp <= n_win; ++sp) {
    const long long waves = (blocks * sp + cap - 1) / cap;
    const long long cost = waves * ((n_win + sp - 1) / sp + 11);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
  }
  return best;
}
template <int WS>
static int bias_grad(const void* gt, int n_win, int nH, int npad, float* partial, float* dtab, cudaStream_t stream) {
  constexpr int VPR = (WS * WS + 7) / 8;
  const int smem = WS * VPR * 8 * sizeof(float);
  const int splits = mvuld_swin_bias_grad_splits(n_win, nH, WS);
  auto kern = splits > 1 ? bias_grad_partial_kernel<WS, true> : bias_grad_partial_kernel<WS, false>;
  MV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<dim3(nH, WS, splits), 256, smem, stream>>>(reinterpret_cast<const bf16*>(gt), n_win, nH, npad, partial);
  MV_LAUNCH_OK();
  constexpr int SIDE = 2 * WS - 1;
  bias_grad_final_kernel<WS><<<(nH * SIDE * SIDE + 255) / 256, 256, 0, stream>>>(partial, dtab, nH, splits);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_swin_bias_grad(const void* gt, int n_win, int nH, int ws, int npad, float* partial, float* dtab,
                                    cudaStream_t stream) {
  MV_CHECK_ARG(npad >= ws * ws && npad % 8 == 0, "swin_bias_grad: npad must cover ws^2 in multiples of 8");
  MV_CHECK_ARG(partial && dtab, "swin_bias_grad: null workspace / output");
  switch (ws) {
    case 28: return bias_grad<28>(gt, n_win, nH, npad, partial, dtab, stream);
    case 14: return bias_grad<14>(gt, n_win, nH, npad, partial, dtab, stream);
    case 7: return bias_grad<7>(gt, n_win, nH, npad, partial, dtab, stream);
    default: return mv::fail(-1, "swin_bias_grad: window %d not instantiated (7, 14, 28)", ws);
  }
}
extern "C" int mvuld_cpb_mlp_bwd(const float* w1, const float* b1, const float* w2, const float* tab, const float* dtab,
                                 int nH, int ws, int pretrained_ws, float* dw1, float* db1, float* dw2,
                                 cudaStream_t stream) {
  MV_CHECK_ARG(nH >= 1 && nH <= CPB_MAXH, "cpb_mlp_bwd: nH in [1, %d]", CPB_MAXH);
  cpb_mlp_bwd_kernel<<<512, 256, 0, stream>>>(w1, b1, w2, tab, dtab, nH, ws, pretrained_ws, dw1, db1, dw2);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_swin_qkv_bwd_blocks(int B, int H, int W, int nH) {
  return (int)(((long long)B * H * W * nH + 255) / 256);
}
extern "C" int mvuld_swin_qkv_bwd(const float* dq, const float* dk, const float* dv, const void* qh, const void* kh,
                                  const float* rq, const float* rk, const float* qscale, const float* logit_scale,
                                  void* dqkv, float* dlogit_scale, float* ls_partial, int B, int H, int W, int C, int nH,
                                  int ws, int shift, cudaStream_t stream) {
  MV_CHECK_ARG(C == nH * 32 && 256 % nH == 0, "swin_qkv_bwd: head_dim 32 and nH dividing 256 (C=%d nH=%d)", C, nH);
  MV_CHECK_ARG(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "swin_qkv_bwd: bad window geometry");
  const int blocks = mvuld_swin_qkv_bwd_blocks(B, H, W, nH);
  if (blocks <= 0) return 0;
  swin_qkv_bwd_kernel<<<blocks, 256, 0, stream>>>(dq, dk, dv, reinterpret_cast<const __half*>(qh),
                                                 reinterpret_cast<const __half*>(kh), rq, rk, qscale,
                                                 reinterpret_cast<bf16*>(dqkv), ls_partial, B, H, W, C, nH, ws, shift);
  MV_LAUNCH_OK();
  logit_scale_final_kernel<<<nH, 256, 0, stream>>>(ls_partial, blocks, nH, logit_scale, dlogit_scale);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_gelu_fwd(const void* x, void* y, long long n, cudaStream_t stream) {
  MV_CHECK_ARG(n % 8 == 0, "gelu_fwd: n %% 8");
  if (n <= 0) return 0;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  gelu_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n / 8);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_patch_merge_scatter(const float* dg, float* dx, int B, int H, int W, int C, cudaStream_t stream) {
  MV_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "patch_merge_scatter: even H, W and C %% 4");
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  patch_merge_scatter_kernel<<<(int)blocks, 256, 0, stream>>>(dg, dx, B, H, W, C);
  MV_LAUNCH_OK();
  return 0;
}
extern "C" int mvuld_patch_im2col(const float* img, void* out, int B, int Hi, int Wi, cudaStream_t stream) {
  MV_CHECK_ARG(Hi % 4 == 0 && Wi % 4 == 0, "patch_im2col: image size must be a multiple of the 4x4 patch");
  const long long total = (long long)B * (Hi / 4) * (Wi / 4) * 12;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  patch_im2col_kernel<<<(int)blocks, 256, 0, stream>>>(img, reinterpret_cast<bf16*>(out), B, Hi, Wi);
  MV_LAUNCH_OK();
  return 0;
}
