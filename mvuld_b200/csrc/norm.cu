// HBM-bound row kernels of the image and text branches: LayerNorm (+ residual, both post- and pre-norm forms),
// patch embedding, patch-merge gather, final LN + token mean, RoBERTa embeddings, masked mean pooling.
// One warp owns one row; 128-bit loads/stores; statistics in fp32 (two-pass in registers, like torch).
#include <cstdlib>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int LN_MAX_UNITS = 4;   // 8-element units per lane -> C <= 1024

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// mode 0: x = LN(y)                       (PatchMerging.norm, swin_transformer_v2.py:362)
// mode 1: x = shortcut + LN(y)            (res-post-norm, swin_transformer_v2.py:301,304)
// mode 2: x = LN(y + shortcut)            (RoBERTa post-LN blocks, HF RobertaSelfOutput / RobertaOutput)
// UNITS = 8-element units per lane the instantiation holds in registers (C <= 256 UNITS): C = 512 rows need 16 values per
// lane, not the 32 of the C = 1024 case -- 60 registers and 4 blocks per SM become <= 51 / 5.
template <int UNITS>
__global__ void __launch_bounds__(256, UNITS <= 2 ? 5 : 4)
ln_rows_kernel(const bf16* __restrict__ y, const float* __restrict__ shortcut, const float* __restrict__ gamma,
               const float* __restrict__ beta, float* __restrict__ x32, bf16* __restrict__ xb, int M, int C, float eps,
               int mode) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int units = C >> 3;
  float v[UNITS][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < UNITS; ++k) {
    const int u = lane + k * 32;
    if (u < units) {
      uint4 raw = __ldg(reinterpret_cast<const uint4*>(y + (size_t)row * C) + u);
      unpack8(raw, v[k]);
      if (mode == 2) {
        const float4* sp = reinterpret_cast<const float4*>(shortcut + (size_t)row * C + u * 8);
        float4 a = __ldg(sp), b = __ldg(sp + 1);
        v[k][0] += a.x; v[k][1] += a.y; v[k][2] += a.z; v[k][3] += a.w;
        v[k][4] += b.x; v[k][5] += b.y; v[k][6] += b.z; v[k][7] += b.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += v[k][q];
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < UNITS; ++k) {
    if (lane + k * 32 < units) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float d = v[k][q] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
  for (int k = 0; k < UNITS; ++k) {
    const int u = lane + k * 32;
    if (u < units) {
      const float4* gp = reinterpret_cast<const float4*>(gamma + u * 8);
      const float4* bp = reinterpret_cast<const float4*>(beta + u * 8);
      float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = (v[k][q] - mean) * rstd * g[q] + b[q];
      if (mode == 1) {
        const float4* sp = reinterpret_cast<const float4*>(shortcut + (size_t)row * C + u * 8);
        float4 a = __ldg(sp), c = __ldg(sp + 1);
        o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
        o[4] += c.x; o[5] += c.y; o[6] += c.z; o[7] += c.w;
      }
      if (x32) {
        float4* op = reinterpret_cast<float4*>(x32 + (size_t)row * C + u * 8);
        op[0] = make_float4(o[0], o[1], o[2], o[3]);
        op[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (xb) reinterpret_cast<uint4*>(xb + (size_t)row * C)[u] = pack8(o);
    }
  }
}

// PatchEmbed (swin_transformer_v2.py:485-493): Conv2d(3, E, k=4, s=4) as a 48-tap dot per token + LayerNorm, fp32.
// One WARP per group of PE_TOK = 8 consecutive tokens: lane l owns the E / 32 consecutive output channels starting at
// l * E / 32; the weights sit transposed in shared memory ([tap][channel]) and each 128-bit weight read serves all 8
// tokens (with 2 tokens per read the kernel was bound by shared-memory bandwidth: 24 KB of weights per pair); the
// 8 x 48 taps are staged per warp with 96 128-bit loads and read back as broadcasts; LayerNorm is two warp reductions
// per token and the outputs leave as coalesced 512-byte (fp32) / 256-byte (bf16) rows.
constexpr int PE_TOK = 8;
template <int E>
__global__ void __launch_bounds__(256)
patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x32,
                   bf16* __restrict__ xb, int B, int Himg, int Wimg, float eps) {
  constexpr int CPL = E / 32;                               // channels per lane
  static_assert(E % 32 == 0 && CPL >= 1 && CPL <= 4, "embed dim must be 32, 64, 96 or 128");
  __shared__ __align__(16) float wT[48][E];                 // wT[tap][channel] = w[channel][tap]
  __shared__ __align__(16) float taps[8][PE_TOK][48];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * E; i += blockDim.x) {
    const int c = i / 48, tp = i - c * 48;
    wT[tp][c] = __ldg(w + i);
  }
  const int Hp = Himg / 4, Wp = Wimg / 4;
  const int total = B * Hp * Wp;                            // tokens (host checks < 2^31)
  float bo[CPL], ga[CPL], be[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    bo[q] = __ldg(bias + lane * CPL + q);
    ga[q] = __ldg(gamma + lane * CPL + q);
    be[q] = __ldg(beta + lane * CPL + q);
  }
  __syncthreads();
  const int ngroups = (total + PE_TOK - 1) / PE_TOK;
  for (int gr = blockIdx.x * 8 + warp; gr < ngroups; gr += gridDim.x * 8) {
    // stage PE_TOK tokens x 12 (channel, kernel-row) segments of 4 contiguous floats, one 128-bit load each
#pragma unroll
    for (int it = 0; it < (PE_TOK * 12 + 31) / 32; ++it) {
      const int j = it * 32 + lane;
      if (j < PE_TOK * 12) {
        const int tk = j / 12, seg = j - tk * 12;
        const int tok = gr * PE_TOK + tk;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok < total) {
          const int b = tok / (Hp * Wp);
          const int rem = tok - b * (Hp * Wp);
          const int ph = rem / Wp, pw = rem - ph * Wp;
          const int c = seg >> 2, kh = seg & 3;
          val = __ldg(reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * Himg + ph * 4 + kh) * Wimg + pw * 4));
        }
        *reinterpret_cast<float4*>(&taps[warp][tk][seg * 4]) = val;
      }
    }
    __syncwarp();
    float acc[PE_TOK][CPL];
#pragma unroll
    for (int tk = 0; tk < PE_TOK; ++tk)
#pragma unroll
      for (int q = 0; q < CPL; ++q) acc[tk][q] = bo[q];
#pragma unroll 2
    for (int i = 0; i < 48; i += 4) {
      float wv[4][CPL];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (CPL == 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&wT[i + k][lane * 4]);
          wv[k][0] = w4.x; wv[k][1] = w4.y; wv[k][2] = w4.z; wv[k][3] = w4.w;
        } else {
#pragma unroll
          for (int q = 0; q < CPL; ++q) wv[k][q] = wT[i + k][lane * CPL + q];
        }
      }
#pragma unroll
      for (int tk = 0; tk < PE_TOK; ++tk) {
        const float4 t4 = *reinterpret_cast<const float4*>(&taps[warp][tk][i]);     // broadcast
        const float a[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < CPL; ++q) acc[tk][q] = fmaf(wv[k][q], a[k], acc[tk][q]);
      }
    }
    __syncwarp();                                            // the taps may be overwritten by the next group
#pragma unroll
    for (int tk = 0; tk < PE_TOK; ++tk) {
      const int tok = gr * PE_TOK + tk;
      float sm = 0.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) sm += acc[tk][q];
      const float mean = warp_sum(sm) / (float)E;
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) sq += (acc[tk][q] - mean) * (acc[tk][q] - mean);
      const float rstd = rsqrtf(warp_sum(sq) / (float)E + eps);
      if (tok >= total) continue;
      float o[CPL];
#pragma unroll
      for (int q = 0; q < CPL; ++q) o[q] = (acc[tk][q] - mean) * rstd * ga[q] + be[q];
      float* op = x32 + (size_t)tok * E + lane * CPL;
      bf16* ob = xb + (size_t)tok * E + lane * CPL;
      if constexpr (CPL == 4) {
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint2*>(ob) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      } else {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          op[q] = o[q];
          ob[q] = __float2bfloat16(o[q]);
        }
      }
    }
  }
}

// PatchMerging gather (swin_transformer_v2.py:352-359): out[b, h2, w2, s*C + c] = x[b, 2h2 + (s&1), 2w2 + (s>>1), c],
// i.e. segment order (0,0), (1,0), (0,1), (1,1).  Pure 128-bit copy.
__global__ void patch_merge_gather_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int H, int W,
                                          int C) {
  const int units = C >> 3;                          // uint4 per token
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
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + hh) * W + ww) * C) + u);
    reinterpret_cast<uint4*>(out)[i] = val;
  }
}

// Final LayerNorm + AdaptiveAvgPool1d(1) over tokens (swin_transformer_v2.py:632-634) -> [B, C] fp32.
__global__ void __launch_bounds__(256)
ln_meanpool_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ out, int T, int C, float eps) {
  extern __shared__ float part[];                    // [8][C]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = C / 32;                            // <= 32
  float acc[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) acc[q] = 0.f;
  for (int t = warp; t < T; t += 8) {
    const float* row = x + ((size_t)b * T + t) * C;
    float v[32];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      if (q < per) {
        v[q] = __ldg(row + lane + q * 32);
        s += v[q];
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q)
      if (q < per) sq += (v[q] - mean) * (v[q] - mean);
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int q = 0; q < 32; ++q)
      if (q < per) acc[q] += (v[q] - mean) * rstd;
  }
#pragma unroll
  for (int q = 0; q < 32; ++q)
    if (q < per) part[warp * C + lane + q * 32] = acc[q];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) s += part[w8 * C + c];
    out[(size_t)b * C + c] = s / (float)T * __ldg(gamma + c) + __ldg(beta + c);
  }
}

// RoBERTa position ids (HF create_position_ids_from_input_ids) and valid length per sequence.
__global__ void __launch_bounds__(512)
seq_positions_kernel(const long long* __restrict__ ids, int L, int pad, int* __restrict__ pos, int* __restrict__ len,
                     int* __restrict__ suffix_ok) {
  __shared__ int scan[512];
  const int b = blockIdx.x, t = threadIdx.x;
  const int m = (t < L && ids[(size_t)b * L + t] != pad) ? 1 : 0;
  scan[t] = m;
  __syncthreads();
  for (int o = 1; o < 512; o <<= 1) {
    int v = (t >= o) ? scan[t - o] : 0;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  if (t < L) pos[(size_t)b * L + t] = scan[t] * m + pad;
  if (t == L - 1) len[b] = scan[t];
  // pads must form a suffix for the kv-length attention path: valid token after a pad => flag
  if (t < L && m && scan[t] != t + 1) atomicExch(suffix_ok, 0);
}

// embeddings = word[ids] + position[pos] + token_type[0] -> LayerNorm (HF RobertaEmbeddings)
__global__ void __launch_bounds__(256)
roberta_embed_kernel(const long long* __restrict__ ids, const int* __restrict__ pos, const float* __restrict__ word,
                     const float* __restrict__ posemb, const float* __restrict__ type0, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ x32, bf16* __restrict__ xb,
                     bf16* __restrict__ ysum, int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* wp = word + (size_t)ids[row] * C;
  const float* pp = posemb + (size_t)pos[row] * C;
  const int per = C / 32;   // 24 for 768
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    if (q < per) {
      const int c = lane + q * 32;
      v[q] = __ldg(wp + c) + __ldg(pp + c) + __ldg(type0 + c);
      if (ysum) {                                  // training: the LayerNorm input as the backward pass will see it
        ysum[(size_t)row * C + c] = __float2bfloat16(v[q]);
        v[q] = __bfloat162float(__float2bfloat16(v[q]));
      }
      s += v[q];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int q = 0; q < 32; ++q)
    if (q < per) sq += (v[q] - mean) * (v[q] - mean);
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    if (q < per) {
      const int c = lane + q * 32;
      const float o = (v[q] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      x32[(size_t)row * C + c] = o;
      xb[(size_t)row * C + c] = __float2bfloat16(o);
    }
  }
}

// sentence = sum_t mask_t * tok_t / sum_t mask_t (unixcoder.py:37)
__global__ void masked_mean_kernel(const float* __restrict__ tok, const int* __restrict__ len, float* __restrict__ out,
                                   int L, int C) {
  const int b = blockIdx.x;
  const int n = len[b];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < n; ++t) s += tok[((size_t)b * L + t) * C + c];
    out[(size_t)b * C + c] = s / (float)n;
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mvuld_ln_rows(const void* y, const float* shortcut, const float* gamma, const float* beta, float* x32,
                             void* xb, int M, int C, float eps, int mode, cudaStream_t stream) {
  MV_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * LN_MAX_UNITS, "ln_rows: C=%d must be a multiple of 8 and <= 1024", C);
  MV_CHECK_ARG(mode == 0 || shortcut, "ln_rows: mode %d needs a shortcut", mode);
  if (M <= 0) return 0;
  const bf16* yp = reinterpret_cast<const bf16*>(y);
  bf16* xp = reinterpret_cast<bf16*>(xb);
  const int grid = (M + 7) / 8;
  if (C <= 256) ln_rows_kernel<1><<<grid, 256, 0, stream>>>(yp, shortcut, gamma, beta, x32, xp, M, C, eps, mode);
  else if (C <= 512) ln_rows_kernel<2><<<grid, 256, 0, stream>>>(yp, shortcut, gamma, beta, x32, xp, M, C, eps, mode);
  else if (C <= 768) ln_rows_kernel<3><<<grid, 256, 0, stream>>>(yp, shortcut, gamma, beta, x32, xp, M, C, eps, mode);
  else ln_rows_kernel<4><<<grid, 256, 0, stream>>>(yp, shortcut, gamma, beta, x32, xp, M, C, eps, mode);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_patch_embed(const float* img, const float* w, const float* bias, const float* gamma,
                                 const float* beta, float* x32, void* xb, int B, int Himg, int Wimg, int E, float eps,
                                 cudaStream_t stream) {
  MV_CHECK_ARG(Himg % 4 == 0 && Wimg % 4 == 0, "patch_embed: image size must be a multiple of the 4x4 patch");
  MV_CHECK_ARG((long long)B * (Himg / 4) * (Wimg / 4) < (1ll << 31) && ((uintptr_t)img % 16) == 0,
               "patch_embed: too many tokens for 32-bit indexing or image base not 16-byte aligned");
  const long long total = (long long)B * (Himg / 4) * (Wimg / 4);
  long long blocks = (total / PE_TOK + 7) / 8 + 1;            // 8 warps per block, PE_TOK tokens per warp and iteration
  const long long cap = (long long)num_sms() * 2;           // 100 registers x 256 threads: two resident blocks per SM
  if (blocks > cap) blocks = cap;
  if (E == 128)
    patch_embed_kernel<128><<<(int)blocks, 256, 0, stream>>>(img, w, bias, gamma, beta, x32,
                                                               reinterpret_cast<bf16*>(xb), B, Himg, Wimg, eps);
  else if (E == 96)
    patch_embed_kernel<96><<<(int)blocks, 256, 0, stream>>>(img, w, bias, gamma, beta, x32, reinterpret_cast<bf16*>(xb),
                                                             B, Himg, Wimg, eps);
  else if (E == 32)
    patch_embed_kernel<32><<<(int)blocks, 256, 0, stream>>>(img, w, bias, gamma, beta, x32, reinterpret_cast<bf16*>(xb),
                                                             B, Himg, Wimg, eps);
  else if (E == 64)
    patch_embed_kernel<64><<<(int)blocks, 256, 0, stream>>>(img, w, bias, gamma, beta, x32, reinterpret_cast<bf16*>(xb),
                                                             B, Himg, Wimg, eps);
  else
    return mv::fail(-1, "patch_embed: embed dim %d not instantiated (32, 64, 96, 128)", E);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_patch_merge_gather(const void* xb, void* out, int B, int H, int W, int C, cudaStream_t stream) {
  MV_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "patch_merge: x size (%d*%d) are not even or C %% 8", H, W);
  const long long total = (long long)B * (H / 2) * (W / 2) * 4 * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  patch_merge_gather_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(xb),
                                                             reinterpret_cast<bf16*>(out), B, H, W, C);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_ln_meanpool(const float* x, const float* gamma, const float* beta, float* out, int B, int T, int C,
                                 float eps, cudaStream_t stream) {
  MV_CHECK_ARG(C % 32 == 0 && C <= 1024, "ln_meanpool: C");
  ln_meanpool_kernel<<<B, 256, 8 * C * sizeof(float), stream>>>(x, gamma, beta, out, T, C, eps);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_seq_positions(const long long* ids, int B, int L, int pad, int* pos, int* len, int* suffix_ok,
                                   cudaStream_t stream) {
  MV_CHECK_ARG(L <= 512, "seq_positions: L <= 512");
  seq_positions_kernel<<<B, 512, 0, stream>>>(ids, L, pad, pos, len, suffix_ok);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_roberta_embed(const long long* ids, const int* pos, const float* word, const float* posemb,
                                   const float* type0, const float* gamma, const float* beta, float* x32, void* xb,
                                   int M, int C, float eps, cudaStream_t stream) {
  MV_CHECK_ARG(C % 32 == 0 && C <= 1024, "roberta_embed: C");
  roberta_embed_kernel<<<(M + 7) / 8, 256, 0, stream>>>(ids, pos, word, posemb, type0, gamma, beta, x32,
                                                        reinterpret_cast<bf16*>(xb), nullptr, M, C, eps);
  MV_LAUNCH_OK();
  return 0;
}
// training forward: also keeps ysum = word + position + type embeddings (bf16 [M, C], the LayerNorm input; the
// normalisation then runs on exactly these rounded values so that mvuld_ln_rows_bwd recomputes the same statistics)
extern "C" int mvuld_roberta_embed_train(const long long* ids, const int* pos, const float* word, const float* posemb,
                                         const float* type0, const float* gamma, const float* beta, float* x32, void* xb,
                                         void* ysum, int M, int C, float eps, cudaStream_t stream) {
  MV_CHECK_ARG(C % 32 == 0 && C <= 1024 && ysum != nullptr, "roberta_embed_train: C %% 32, C <= 1024, ysum non-null");
  roberta_embed_kernel<<<(M + 7) / 8, 256, 0, stream>>>(ids, pos, word, posemb, type0, gamma, beta, x32,
                                                        reinterpret_cast<bf16*>(xb), reinterpret_cast<bf16*>(ysum), M, C,
                                                        eps);
  MV_LAUNCH_OK();
  return 0;
}

namespace mv {
// mean over the token rows [start[s], start[s] + len[s]) of tok fp32 [T, C] -> out fp32 [n, C]: the sentence vector
// of each packed line (unixcoder.py:37 per line).  Grid (segment, 64-column chunk); the 256 threads of a block are 64
// columns x 4 token phases (phase p sums tokens p, p + 4, ... four loads in flight), combined in a fixed order through
// shared memory -- 64 function-level segments of ~257 tokens x 768 columns used to run as 64 blocks with one serial
// chain per column (132 us); this shape gives 768 blocks (deterministic: the summation order depends on n only).
__global__ void __launch_bounds__(256)
segment_mean_kernel(const float* __restrict__ tok, const int* __restrict__ start, const int* __restrict__ len,
                    const int* __restrict__ dst, float* __restrict__ out, int C) {
  __shared__ float part[4][64];
  const int sgm = blockIdx.x;
  const long long b = start[sgm];
  const int n = len[sgm];
  const size_t orow = dst ? (size_t)dst[sgm] : (size_t)sgm;      // output row of this segment (packing may reorder)
  const int cl = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    const float* p = tok + b * C + c;
    int t = ph;
    for (; t + 12 < n; t += 16) {
      a0 += __ldg(p + (size_t)t * C);
      a1 += __ldg(p + (size_t)(t + 4) * C);
      a2 += __ldg(p + (size_t)(t + 8) * C);
      a3 += __ldg(p + (size_t)(t + 12) * C);
    }
    for (; t < n; t += 4) a0 += __ldg(p + (size_t)t * C);
  }
  part[ph][cl] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ph == 0 && c < C) out[orow * C + c] = ((part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl])) / (float)n;
}
}  // namespace mv
extern "C" int mvuld_seq_segment_mean(const float* tok, const int* seg_start, const int* seg_len, const int* out_row,
                                      float* out, int n, int C, cudaStream_t stream) {
  if (n <= 0) return 0;
  mv::segment_mean_kernel<<<dim3(n, (C + 63) / 64), 256, 0, stream>>>(tok, seg_start, seg_len, out_row, out, C);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_masked_mean(const float* tok, const int* len, float* out, int B, int L, int C,
                                 cudaStream_t stream) {
  masked_mean_kernel<<<B, 256, 0, stream>>>(tok, len, out, L, C);
  MV_LAUNCH_OK();
  return 0;
}
