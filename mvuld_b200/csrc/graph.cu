// Graph-branch kernels (HBM-bound, warp-level CSR gather / segment reduce, 128-bit loads):
//   * in-edge CSR (CSC) construction sorted by (dst, edge id) -- the grouping DGL's edge_softmax / SpMM reduce over
//   * GatedGraphConv message aggregation  a[dst] = sum_{e -> dst} M[src(e), etype(e)]      
//   * GRU gate update                                                                       (K19)
//   * GATConv attention scores + edge-softmax + weighted aggregation in ONE pass per dst    (K13)
//   * per-graph segment sum readout                                                         (K20)
//   * unbatch -> pad/truncate to max_node rows, with the node-slot BatchNorm folded in      (K15, K16)
// Reference call sites: GraphModel.py:30-54,99-105,167-170,180-187; baselines/models/reveal/ggnn/model.py:15-31;
// baselines/models/devign/model.py:15-16,35.  DGL semantics restated in SURVEY.md section 8(c).
#include <cub/cub.cuh>

#include <cstdlib>
#include "common.cuh"
#include "host_util.h"

namespace mv {

// ---------------------------------------------- CSR build ----------------------------------------------
__global__ void csr_prepare_kernel(const long long* __restrict__ dst, int E, int N, int* __restrict__ keys,
                                   int* __restrict__ vals, int* __restrict__ counts, int* __restrict__ bad) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const long long d = dst[e];
    if (d < 0 || d >= N) {
      atomicExch(bad, 1);
      keys[e] = 0;
    } else {
      keys[e] = (int)d;
      atomicAdd(&counts[(int)d + 1], 1);
    }
    vals[e] = e;
  }
}
__global__ void csr_finish_kernel(const long long* __restrict__ src, const int* __restrict__ eid_sorted, int E, int N,
                                  int* __restrict__ idx_src, int* __restrict__ bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
    const long long s = src[eid_sorted[i]];
    if (s < 0 || s >= N) atomicExch(bad, 1);
    idx_src[i] = (int)s;
  }
}
__global__ void gather_etype_kernel(const long long* __restrict__ etype, const int* __restrict__ eid_sorted, int E,
                                    int n_etypes, unsigned char* __restrict__ out, int* __restrict__ bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
    const long long t = etype[eid_sorted[i]];
    if (t < 0 || t >= n_etypes) atomicExch(bad, 1);     // DGL: assert 0 <= etypes < n_etypes
    out[i] = (unsigned char)t;
  }
}

// ------------------------------------------- collate (dgl.add_self_loop + dgl.batch) -------------------------------------------
// Raw per-graph edge lists with LOCAL node ids, concatenated graph by graph (src_l / dst_l int32, etype_l int64 or
// null), edge_off / node_off = exclusive prefix sums of the per-graph edge / node counts (int64 [B + 1]).
// Output in DGL order: graph k contributes its E_k edges in input order with ids shifted by
// node_off[k], then -- when add_loops -- its N_k self loops (i, i) with zero-filled edge data
// (mvuld/data/data_list.py:314 add_self_loop before mvuld/data/bigvul_dataset.py:177-205 batch).
__global__ void collate_edges_kernel(const int* __restrict__ src_l, const int* __restrict__ dst_l,
                                     const long long* __restrict__ etype_l, const long long* __restrict__ edge_off,
                                     const long long* __restrict__ node_off, int B, int add_loops,
                                     long long* __restrict__ src, long long* __restrict__ dst,
                                     long long* __restrict__ etype, long long total, int* __restrict__ bad) {
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = B - 1;                       // graph k with out_off[k] <= o < out_off[k + 1]
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const long long start = edge_off[mid] + (add_loops ? node_off[mid] : 0);
      if (start <= o) lo = mid; else hi = mid - 1;
    }
    const long long nb = node_off[lo], eb = edge_off[lo];
    const long long ek = edge_off[lo + 1] - eb, nk = node_off[lo + 1] - nb;
    const long long l = o - (eb + (add_loops ? nb : 0));
    if (l < ek) {
      const int sl = src_l[eb + l], dl = dst_l[eb + l];
      if (sl < 0 || sl >= nk || dl < 0 || dl >= nk) atomicExch(bad, 1);
      src[o] = nb + sl;
      dst[o] = nb + dl;
      if (etype) etype[o] = etype_l ? etype_l[eb + l] : 0;
    } else {
      src[o] = nb + (l - ek);
      dst[o] = nb + (l - ek);
      if (etype) etype[o] = 0;
    }
  }
}

// ------------------------------------------- GGNN aggregation -------------------------------------------
// msgs: bf16 [N, T, D] (per-etype linear already applied, bias included).  One warp owns NPW consecutive destination
// nodes and walks their (contiguous) in-edge range as one flat list, UNROLL gathered rows in flight at a time: a
// warp-per-node version spent two thirds of its time in the indptr -> (src, etype) -> row dependency chain with
// nothing in flight (measured 3.1 TB/s DRAM); here the chain is paid once per NPW nodes and row loads of the next
// node are issued while the previous node is still being summed.  Sums run in edge order per node (fp32).
template <int UNROLL, int NPW, int MINB = 1>
__global__ void __launch_bounds__(256, MINB)
ggnn_gather_sum_kernel(const bf16* __restrict__ msgs, const int* __restrict__ indptr, const int* __restrict__ idx_src,
                       const unsigned char* __restrict__ etype, bf16* __restrict__ out, int ldo, int N, int T, int D) {
  const int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * NPW;
  if (n0 >= N) return;
  const int lane = threadIdx.x & 31;
  const int units = D >> 3;                 // uint4 per row (25 for D = 200)
  const bool active = lane < units;
  const int n_nodes = min(NPW, N - n0);
  const int my_ptr = __ldg(indptr + n0 + min(lane, n_nodes));        // lanes 0..n_nodes hold the boundaries
  const int e_beg = __shfl_sync(0xffffffffu, my_ptr, 0);
  const int e_end = __shfl_sync(0xffffffffu, my_ptr, n_nodes);
  int cur = 0;                                                          // node (relative to n0) being summed
  int boundary = __shfl_sync(0xffffffffu, my_ptr, 1);                   // first edge of the next node
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  auto flush = [&]() {
    if (active) {
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
      reinterpret_cast<uint4*>(out + (size_t)(n0 + cur) * ldo)[lane] = o;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    ++cur;
    boundary = __shfl_sync(0xffffffffu, my_ptr, min(cur + 1, n_nodes));
  };
  for (int base = e_beg; base < e_end; base += 32) {
    const int n_here = min(32, e_end - base);
    int my_row = 0;
    if (lane < n_here) my_row = __ldg(idx_src + base + lane) * T + (int)__ldg(etype + base + lane);
    for (int e = 0; e < n_here; e += UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int row = __shfl_sync(0xffffffffu, my_row, (e + u) & 31);
        v[u] = (active && e + u < n_here) ? __ldg(reinterpret_cast<const uint4*>(msgs + (size_t)row * D) + lane)
                                          : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (e + u < n_here) {
          while (base + e + u >= boundary && cur < n_nodes - 1) flush();   // (a node without in-edges gets zeros)
          acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x); acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
          acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z); acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
        }
      }
    }
  }
  while (cur < n_nodes) flush();            // the last node (and any trailing nodes without in-edges)
}

// h0 = cat(x, zeros[N, D - in]) (GatedGraphConv zero-pad) -> fp32 state + bf16 shadow
__global__ void ggnn_init_kernel(const float* __restrict__ x, float* __restrict__ h32, bf16* __restrict__ hb, int ldb,
                                 long long N, int in_dim, int D) {
  const long long total = N * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / D;
    const int c = (int)(i % D);
    const float v = c < in_dim ? x[n * in_dim + c] : 0.f;
    h32[i] = v;
    hb[n * ldb + c] = __float2bfloat16(v);
  }
}

// --------------------------------------------- segment sum ---------------------------------------------
// out[b, :] = sum of rows [off[b], off[b+1]) of feat (fp32 [N, D]); one block per graph, float4 columns,
// RG row lanes per column reduced through shared memory at the end.
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ feat, const long long* __restrict__ off, float* __restrict__ out, int D) {
  extern __shared__ float4 red[];
  const int b = blockIdx.x;
  const int vec = D >> 2;
  const int rg = blockDim.x / vec;
  const int col = threadIdx.x % vec;
  const int rl = threadIdx.x / vec;
  const long long beg = off[b], end = off[b + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < rg) {
    long long r = beg + rl;
    for (; r + 3LL * rg < end; r += 4LL * rg) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(feat + r * D) + col);
      const float4 c = __ldg(reinterpret_cast<const float4*>(feat + (r + rg) * D) + col);
      const float4 d = __ldg(reinterpret_cast<const float4*>(feat + (r + 2LL * rg) * D) + col);
      const float4 e = __ldg(reinterpret_cast<const float4*>(feat + (r + 3LL * rg) * D) + col);
      acc.x += (a.x + c.x) + (d.x + e.x); acc.y += (a.y + c.y) + (d.y + e.y);
      acc.z += (a.z + c.z) + (d.z + e.z); acc.w += (a.w + c.w) + (d.w + e.w);
    }
    for (; r < end; r += rg) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(feat + r * D) + col);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    red[rl * vec + col] = acc;
  }
  __syncthreads();
  if (threadIdx.x < vec) {
    float4 s = red[threadIdx.x];
    for (int k = 1; k < rg; ++k) {
      const float4 a = red[k * vec + threadIdx.x];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    reinterpret_cast<float4*>(out + (size_t)b * D)[threadIdx.x] = s;
  }
}

// ------------------------------------------------- GAT -------------------------------------------------
// el[n,h] = <z[n,h,:], attn_l[h,:]>, er likewise.  z bf16 [N, H*F]; one warp per node.
__global__ void __launch_bounds__(256)
gat_scores_kernel(const bf16* __restrict__ z, const float* __restrict__ attn_l, const float* __restrict__ attn_r,
                  float* __restrict__ el, float* __restrict__ er, int N, int H, int F) {
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int lane = threadIdx.x & 31;
  const int upr = F >> 3;                    // uint4 per head
  for (int h = 0; h < H; ++h) {
    float sl = 0.f, sr = 0.f;
    for (int u = lane; u < upr; u += 32) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(z + ((size_t)node * H + h) * F) + u);
      const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                          bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
      const float4* lp = reinterpret_cast<const float4*>(attn_l + h * F + u * 8);
      const float4* rp = reinterpret_cast<const float4*>(attn_r + h * F + u * 8);
      const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), r0 = __ldg(rp), r1 = __ldg(rp + 1);
      sl += f[0] * l0.x + f[1] * l0.y + f[2] * l0.z + f[3] * l0.w + f[4] * l1.x + f[5] * l1.y + f[6] * l1.z + f[7] * l1.w;
      sr += f[0] * r0.x + f[1] * r0.y + f[2] * r0.z + f[3] * r0.w + f[4] * r1.x + f[5] * r1.y + f[6] * r1.z + f[7] * r1.w;
    }
    sl = warp_sum(sl);
    sr = warp_sum(sr);
    if (lane == 0) {
      el[(size_t)node * H + h] = sl;
      er[(size_t)node * H + h] = sr;
    }
  }
}

// One warp per destination node: edge-softmax over its in-edges (per head, max-subtracted) and
// out[dst] = sum alpha * z[src] + bias.  Row = H*F bf16 (4 KB for 4 x 512) = CH chunks of 256 elements.
template <int NH, int CPH>
__global__ void __launch_bounds__(128)
gat_aggregate_kernel(const bf16* __restrict__ z, const float* __restrict__ el, const float* __restrict__ er,
                     const int* __restrict__ indptr, const int* __restrict__ idx_src, const float* __restrict__ bias,
                     bf16* __restrict__ out, int N, int H, int F, float slope, int* __restrict__ zero_deg) {
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int lane = threadIdx.x & 31;
  const int beg = __ldg(indptr + node), end = __ldg(indptr + node + 1);
  constexpr int CH = NH * CPH;               // 256-element chunks per row (CPH per head)
  if (end == beg) {                          // DGL raises on 0-in-degree nodes (allow_zero_in_degree=False)
    if (lane == 0) atomicExch(zero_deg, 1);
    return;
  }
  // pass A/B: per-head max and sum of exp over in-edges (lane == edge, chunks of 32 edges)
  float hmax[NH], hsum[NH], erd[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    hmax[h] = -INFINITY;
    hsum[h] = 0.f;
    erd[h] = __ldg(er + (size_t)node * H + h);
  }
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    const int s = e < end ? __ldg(idx_src + e) : 0;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float sc = -INFINITY;
      if (e < end) {
        sc = __ldg(el + (size_t)s * H + h) + erd[h];
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
        float sc = __ldg(el + (size_t)s * H + h) + erd[h];
        sc = sc > 0.f ? sc : sc * slope;
        pe = __expf(sc - hmax[h]);
      }
      hsum[h] += warp_sum(pe);
    }
  }
  // pass C: weighted gather-sum of the 4 KB source rows
  float acc[CH][8];
#pragma unroll
  for (int k = 0; k < CH; ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[k][q] = 0.f;
  for (int e = beg; e < end; ++e) {
    const int s = __ldg(idx_src + e);
    float alpha[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float sc = __ldg(el + (size_t)s * H + h) + erd[h];
      sc = sc > 0.f ? sc : sc * slope;
      alpha[h] = __expf(sc - hmax[h]) / hsum[h];
    }
    const uint4* row = reinterpret_cast<const uint4*>(z + (size_t)s * H * F);
    uint4 v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) v[k] = __ldg(row + k * 32 + lane);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const float a = alpha[k / CPH];
      acc[k][0] += a * bf16_lo(v[k].x); acc[k][1] += a * bf16_hi(v[k].x);
      acc[k][2] += a * bf16_lo(v[k].y); acc[k][3] += a * bf16_hi(v[k].y);
      acc[k][4] += a * bf16_lo(v[k].z); acc[k][5] += a * bf16_hi(v[k].z);
      acc[k][6] += a * bf16_lo(v[k].w); acc[k][7] += a * bf16_hi(v[k].w);
    }
  }
  uint4* orow = reinterpret_cast<uint4*>(out + (size_t)node * H * F);
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const float4* bp = reinterpret_cast<const float4*>(bias + k * 256 + lane * 8);
    const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
    uint4 o;
    o.x = pack_bf16x2(acc[k][0] + b0.x, acc[k][1] + b0.y); o.y = pack_bf16x2(acc[k][2] + b0.z, acc[k][3] + b0.w);
    o.z = pack_bf16x2(acc[k][4] + b1.x, acc[k][5] + b1.y); o.w = pack_bf16x2(acc[k][6] + b1.z, acc[k][7] + b1.w);
    orow[k * 32 + lane] = o;
  }
}

// ------------------------------------------ unbatch + pad + slot BN ------------------------------------------
// out[b, r, :] = (x_pad[b, r, :] - mean[r]) * rsqrt(var[r] + eps) * w[r] + beta[r], x_pad = node row off[b]+r if
// r < min(N_b, max_node) else 0  (GraphModel.py:30-54 then BatchNorm1d(max_node) over the slot axis, :135,186).
// Also emits the integer gather map (-1 = zero row) so tests can compare it bit-exactly.
__global__ void unbatch_pad_bn_kernel(const bf16* __restrict__ feat, const long long* __restrict__ off,
                                      const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                      bf16* __restrict__ out, long long* __restrict__ gather_map, int B, int max_node,
                                      int F) {
  const int units = F >> 3;
  const long long total = (long long)B * max_node * units;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    const long long br = i / units;
    const int r = (int)(br % max_node);
    const int b = (int)(br / max_node);
    const long long beg = off[b], n = off[b + 1] - beg;
    const bool have = r < n;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (have) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(feat + (beg + r) * F) + u);
      f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
      f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
    const float sc = __ldg(bn_scale + r), sh = __ldg(bn_shift + r);
    uint4 o;
    o.x = pack_bf16x2(f[0] * sc + sh, f[1] * sc + sh); o.y = pack_bf16x2(f[2] * sc + sh, f[3] * sc + sh);
    o.z = pack_bf16x2(f[4] * sc + sh, f[5] * sc + sh); o.w = pack_bf16x2(f[6] * sc + sh, f[7] * sc + sh);
    reinterpret_cast<uint4*>(out)[i] = o;
    if (u == 0 && gather_map) gather_map[br] = have ? beg + r : -1;
  }
}

// pos branch: ELU(fc_bbox(bn_bbox(pad(pos)))) written into columns [col0, col0+OUT) of the [B*max_node, ld] concat
// buffers (GraphModel.py:137-138,187,189).  pos fp32 [N, 4].
__global__ void pos_branch_kernel(const float* __restrict__ pos, const long long* __restrict__ off,
                                  const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                  const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ z32,
                                  bf16* __restrict__ zb, int B, int max_node, int OUT, int ld, int col0) {
  const long long total = (long long)B * max_node * OUT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % OUT);
    const long long br = i / OUT;
    const int r = (int)(br % max_node);
    const int b = (int)(br / max_node);
    const long long beg = off[b], n = off[b + 1] - beg;
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pos + (beg + r) * 4));
      p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    }
    const float sc = __ldg(bn_scale + r), sh = __ldg(bn_shift + r);
    float acc = __ldg(bias + o);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += (p[k] * sc + sh) * __ldg(w + o * 4 + k);
    acc = elu1(acc);
    z32[br * ld + col0 + o] = acc;
    zb[br * ld + col0 + o] = __float2bfloat16(acc);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mv

using namespace mv;

// In-edge CSR sorted by (dst, edge id).  Call with workspace == NULL to get the required workspace bytes in
// *workspace_bytes.  Outputs: indptr int32 [N+1], idx_src int32 [E], eids int32 [E]; status int32 [1] is set to 1 when
// an endpoint is outside [0, N).
extern "C" int mvuld_csr_from_coo(const long long* src, const long long* dst, int E, int N, void* workspace,
                                  size_t* workspace_bytes, int* indptr, int* idx_src, int* eids, int* status,
                                  cudaStream_t stream) {
  MV_CHECK_ARG(E >= 0 && N >= 0, "csr: negative sizes");
  int end_bit = 1;
  while ((1LL << end_bit) < (long long)N + 1 && end_bit < 31) ++end_bit;
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, E, 0, end_bit, stream);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, N + 1, stream);
  const size_t tmp = ((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 255) / 256 * 256;
  const size_t arr = ((size_t)E * sizeof(int) + 255) / 256 * 256;
  const size_t need = tmp + 3 * arr + 256;
  if (!workspace) {
    MV_CHECK_ARG(workspace_bytes, "csr: workspace_bytes is null");
    *workspace_bytes = need;
    return 0;
  }
  MV_CHECK_ARG(!workspace_bytes || *workspace_bytes >= need, "csr: workspace too small");
  char* wsp = reinterpret_cast<char*>(workspace);
  void* d_tmp = wsp;
  int* keys_in = reinterpret_cast<int*>(wsp + tmp);
  int* vals_in = reinterpret_cast<int*>(wsp + tmp + arr);
  int* keys_out = reinterpret_cast<int*>(wsp + tmp + 2 * arr);
  MV_CUDA_OK(cudaMemsetAsync(indptr, 0, (size_t)(N + 1) * sizeof(int), stream));
  MV_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), stream));
  if (E > 0) {
    csr_prepare_kernel<<<grid_for(E, 256), 256, 0, stream>>>(dst, E, N, keys_in, vals_in, indptr, status);
    MV_LAUNCH_OK();
    size_t sb = sort_bytes;
    MV_CUDA_OK(cub::DeviceRadixSort::SortPairs(d_tmp, sb, keys_in, keys_out, vals_in, eids, E, 0, end_bit, stream));
    csr_finish_kernel<<<grid_for(E, 256), 256, 0, stream>>>(src, eids, E, N, idx_src, status);
    MV_LAUNCH_OK();
  }
  size_t cb = scan_bytes;
  MV_CUDA_OK(cub::DeviceScan::InclusiveSum(d_tmp, cb, indptr, indptr, N + 1, stream));
  return 0;
}

extern "C" int mvuld_collate_edges(const int* src_local, const int* dst_local, const long long* etype_local,
                                   const long long* edge_off, const long long* node_off, int B, int add_self_loops,
                                   long long* src, long long* dst, long long* etype, long long total_out, int* status,
                                   cudaStream_t stream) {
  MV_CHECK_ARG(B >= 1 && total_out >= 0, "collate: empty batch");
  if (total_out == 0) return 0;
  collate_edges_kernel<<<grid_for(total_out, 256), 256, 0, stream>>>(src_local, dst_local, etype_local, edge_off,
                                                                     node_off, B, add_self_loops, src, dst, etype,
                                                                     total_out, status);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_gather_etype(const long long* etype, const int* eids, int E, int n_etypes, unsigned char* out,
                                  int* status, cudaStream_t stream) {
  if (E <= 0) return 0;
  gather_etype_kernel<<<grid_for(E, 256), 256, 0, stream>>>(etype, eids, E, n_etypes, out, status);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_ggnn_gather_sum(const void* msgs, const int* indptr, const int* idx_src,
                                     const unsigned char* etype, void* out, int ldo, int N, int T, int D,
                                     cudaStream_t stream) {
  MV_CHECK_ARG(D % 8 == 0 && D <= 256, "ggnn_gather_sum: D must be a multiple of 8 and <= 256");
  MV_CHECK_ARG(ldo >= D && ldo % 8 == 0, "ggnn_gather_sum: ldo must be >= D and a multiple of 8");
  if (N <= 0) return 0;
  const bf16* mp = reinterpret_cast<const bf16*>(msgs);
  bf16* op = reinterpret_cast<bf16*>(out);
  // 4 row loads in flight per lane at 48 registers / 5 blocks per SM: measured 334 us on configs[2] (6.0 TB/s of
  // algorithmic bytes) against 430 us for 8 in flight at 78 registers / 3 blocks -- occupancy hides the gather latency
  // better than per-warp depth.
  ggnn_gather_sum_kernel<4, 8, 5><<<(N + 63) / 64, 256, 0, stream>>>(mp, indptr, idx_src, etype, op, ldo, N, T, D);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_ggnn_init(const float* x, float* h32, void* hb, int ldb, long long N, int in_dim, int D,
                               cudaStream_t stream) {
  MV_CHECK_ARG(in_dim <= D, "ggnn_init: in_feats must be <= out_feats");
  if (N <= 0) return 0;
  ggnn_init_kernel<<<grid_for(N * D, 256), 256, 0, stream>>>(x, h32, reinterpret_cast<bf16*>(hb), ldb, N, in_dim, D);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_segment_sum(const float* feat, const long long* offsets, float* out, int B, int D,
                                 cudaStream_t stream) {
  MV_CHECK_ARG(D % 4 == 0 && D / 4 <= 256, "segment_sum: D must be a multiple of 4 and <= 1024");
  if (B <= 0) return 0;
  const int vec = D / 4;
  const int rg = 256 / vec;
  segment_sum_kernel<<<B, 256, (size_t)rg * vec * sizeof(float4), stream>>>(feat, offsets, out, D);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_gat_scores(const void* z, const float* attn_l, const float* attn_r, float* el, float* er, int N,
                                int H, int F, cudaStream_t stream) {
  MV_CHECK_ARG(F % 8 == 0, "gat_scores: F %% 8");
  if (N <= 0) return 0;
  gat_scores_kernel<<<(N + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const bf16*>(z), attn_l, attn_r, el, er, N, H, F);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_gat_aggregate(const void* z, const float* el, const float* er, const int* indptr,
                                   const int* idx_src, const float* bias, void* out, int N, int H, int F, float slope,
                                   int* zero_deg_flag, cudaStream_t stream) {
  if (N <= 0) return 0;
  const bf16* zp = reinterpret_cast<const bf16*>(z);
  bf16* op = reinterpret_cast<bf16*>(out);
  const int grid = (N + 3) / 4;
  if (H == 4 && F == 512)
    gat_aggregate_kernel<4, 2><<<grid, 128, 0, stream>>>(zp, el, er, indptr, idx_src, bias, op, N, H, F, slope, zero_deg_flag);
  else if (H == 4 && F == 256)
    gat_aggregate_kernel<4, 1><<<grid, 128, 0, stream>>>(zp, el, er, indptr, idx_src, bias, op, N, H, F, slope, zero_deg_flag);
  else if (H == 2 && F == 256)
    gat_aggregate_kernel<2, 1><<<grid, 128, 0, stream>>>(zp, el, er, indptr, idx_src, bias, op, N, H, F, slope, zero_deg_flag);
  else
    return mv::fail(-1, "gat_aggregate: (H=%d, F=%d) not instantiated: (4,512), (4,256), (2,256)", H, F);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_unbatch_pad_bn(const void* feat, const long long* offsets, const float* bn_scale,
                                    const float* bn_shift, void* out, long long* gather_map, int B, int max_node, int F,
                                    cudaStream_t stream) {
  MV_CHECK_ARG(F % 8 == 0, "unbatch_pad_bn: F %% 8");
  if (B <= 0) return 0;
  unbatch_pad_bn_kernel<<<grid_for((long long)B * max_node * (F / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(feat), offsets, bn_scale, bn_shift, reinterpret_cast<bf16*>(out), gather_map, B,
      max_node, F);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_pos_branch(const float* pos, const long long* offsets, const float* bn_scale,
                                const float* bn_shift, const float* w, const float* bias, float* z32, void* zb, int B,
                                int max_node, int OUT, int ld, int col0, cudaStream_t stream) {
  if (B <= 0) return 0;
  pos_branch_kernel<<<grid_for((long long)B * max_node * OUT, 256), 256, 0, stream>>>(
      pos, offsets, bn_scale, bn_shift, w, bias, z32, reinterpret_cast<bf16*>(zb), B, max_node, OUT, ld, col0);
  MV_LAUNCH_OK();
  return 0;
}

extern "C" int mvuld_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, stream>>>(in, reinterpret_cast<bf16*>(out), n);
  MV_LAUNCH_OK();
  return 0;
}
