"""Oracle (test infrastructure): DGL 0.8.1 graph semantics restated in numpy / torch.

**Parity unpinned**: DGL (``dgl-cu102==0.8.1``, /root/reference/environment.yml:159) is a
third-party dependency whose source is not under /root/reference and which cannot be installed
offline.  Each function restates the library's documented behaviour and
cites the reference call site that relies on it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class HostGraph:
    """One homogeneous graph as dgl.graph((src, dst)) would hold it (data_list.py:279)."""
    src: np.ndarray            # int64 [E]
    dst: np.ndarray            # int64 [E]
    num_nodes: int
    ndata: Dict[str, torch.Tensor] = field(default_factory=dict)
    edata: Dict[str, torch.Tensor] = field(default_factory=dict)


def graph(src, dst, num_nodes=None) -> HostGraph:
    """dgl.graph((src,dst)): ids int64, num_nodes = max id + 1, edge id = input position."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    n = int(max(src.max(initial=-1), dst.max(initial=-1)) + 1) if num_nodes is None else int(num_nodes)
    return HostGraph(src, dst, n)


def add_self_loop(g: HostGraph) -> HostGraph:
    """dgl.add_self_loop (data_list.py:314): append (i,i) for i in 0..N-1 AFTER existing edges,
    no dedup; new edata rows are zero-filled (so _ETYPE = 0 for them)."""
    loop = np.arange(g.num_nodes, dtype=np.int64)
    out = HostGraph(np.concatenate([g.src, loop]), np.concatenate([g.dst, loop]), g.num_nodes, dict(g.ndata))
    for k, v in g.edata.items():
        pad = torch.zeros((g.num_nodes,) + tuple(v.shape[1:]), dtype=v.dtype)
        out.edata[k] = torch.cat([v, pad], 0)
    return out


@dataclass
class HostBatch:
    src: np.ndarray
    dst: np.ndarray
    num_nodes: int
    batch_num_nodes: np.ndarray    # int64 [B]
    batch_num_edges: np.ndarray    # int64 [B]
    ndata: Dict[str, torch.Tensor]
    edata: Dict[str, torch.Tensor]


def batch(graphs: Sequence[HostGraph]) -> HostBatch:
    """dgl.batch (GraphDataLoader collate, bigvul_dataset.py:177-205): node ids of graph k are
    shifted by sum_{j<k} N_j; edges concatenated in graph order then original edge order."""
    bnn = np.array([g.num_nodes for g in graphs], dtype=np.int64)
    bne = np.array([len(g.src) for g in graphs], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(bnn)[:-1]]).astype(np.int64) if len(graphs) else np.zeros(0, np.int64)
    src = np.concatenate([g.src + o for g, o in zip(graphs, off)]) if len(graphs) else np.zeros(0, np.int64)
    dst = np.concatenate([g.dst + o for g, o in zip(graphs, off)]) if len(graphs) else np.zeros(0, np.int64)
    nd = {k: torch.cat([g.ndata[k] for g in graphs], 0) for k in (graphs[0].ndata if graphs else {})}
    ed = {k: torch.cat([g.edata[k] for g in graphs], 0) for k in (graphs[0].edata if graphs else {})}
    return HostBatch(src, dst, int(bnn.sum()), bnn, bne, nd, ed)


def node_offsets(batch_num_nodes: np.ndarray) -> np.ndarray:
    """dgl.unbatch segment boundaries: graph k owns rows [off[k], off[k+1]). int64 [B+1]."""
    return np.concatenate([[0], np.cumsum(np.asarray(batch_num_nodes, dtype=np.int64))]).astype(np.int64)


def in_csr(src: np.ndarray, dst: np.ndarray, num_nodes: int):
    """CSC / in-edge CSR sorted by (dst, edge id) -- the grouping DGL's edge_softmax and SpMM
    reduce over.  Returns (indptr int64 [N+1], indices(src) int64 [E], edge_ids int64 [E])."""
    eid = np.argsort(dst, kind="stable").astype(np.int64)
    indptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(indptr, dst + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int64)
    return indptr, src[eid].astype(np.int64), eid


def pad_truncate_map(batch_num_nodes: np.ndarray, max_node: int) -> np.ndarray:
    """unbatch_features row map (GraphModel.py:30-54): slot r of graph k <- node off_k + r if
    r < min(N_k, max_node) else -1 (zero row).  int64 [B, max_node]."""
    off = node_offsets(batch_num_nodes)
    r = np.arange(max_node, dtype=np.int64)[None, :]
    n = np.asarray(batch_num_nodes, dtype=np.int64)[:, None]
    return np.where(r < n, off[:-1, None] + r, -1).astype(np.int64)


def unbatch_pad(feat: torch.Tensor, batch_num_nodes: np.ndarray, max_node: int) -> torch.Tensor:
    """GraphModel.py:30-54,183 -> [B, max_node, F]."""
    m = torch.from_numpy(pad_truncate_map(batch_num_nodes, max_node))
    out = feat[m.clamp(min=0)]
    return out * (m >= 0).unsqueeze(-1).to(feat.dtype)


def segment_sum(feat: torch.Tensor, batch_num_nodes: np.ndarray) -> torch.Tensor:
    """Padded stack + sum(dim=1) of reveal/ggnn/model.py:26-28,46-56 == per-graph segment sum."""
    off = node_offsets(batch_num_nodes)
    out = torch.zeros((len(batch_num_nodes),) + tuple(feat.shape[1:]), dtype=feat.dtype)
    for k in range(len(batch_num_nodes)):
        out[k] = feat[off[k]:off[k + 1]].sum(0)
    return out


def mean_nodes(feat: torch.Tensor, batch_num_nodes: np.ndarray) -> torch.Tensor:
    """dgl.mean_nodes (GraphModel.py:298): per-graph mean, empty graph -> 0."""
    s = segment_sum(feat, batch_num_nodes)
    n = torch.from_numpy(np.maximum(np.asarray(batch_num_nodes), 1)).to(feat.dtype)
    return s / n.view(-1, *([1] * (feat.dim() - 1)))


@torch.no_grad()
def gat_conv(sd: Dict[str, torch.Tensor], prefix: str, src, dst, x: torch.Tensor, num_heads: int,
             out_feats: int, negative_slope: float = 0.2) -> torch.Tensor:
    """DGL GATConv forward as called at GraphModel.py:99-105,167-170 (eval: feat_drop off,
    attn_drop=0, residual=False, activation=None, bias=True) -> [N, H, out]."""
    N = x.shape[0]
    src_t = torch.as_tensor(src, dtype=torch.long)
    dst_t = torch.as_tensor(dst, dtype=torch.long)
    deg = torch.zeros(N, dtype=torch.long).index_add_(0, dst_t, torch.ones_like(dst_t))
    if (deg == 0).any():
        raise RuntimeError("There are 0-in-degree nodes in the graph (allow_zero_in_degree=False)")
    z = F.linear(x.float(), sd[prefix + "fc.weight"].float()).view(N, num_heads, out_feats)
    el = (z * sd[prefix + "attn_l"].float()).sum(-1)              # [N, H]
    er = (z * sd[prefix + "attn_r"].float()).sum(-1)
    e = F.leaky_relu(el[src_t] + er[dst_t], negative_slope)      # [E, H]
    emax = torch.full((N, num_heads), -float("inf")).scatter_reduce(0, dst_t[:, None].expand_as(e), e, "amax")
    p = torch.exp(e - emax[dst_t])
    den = torch.zeros(N, num_heads).index_add_(0, dst_t, p)
    alpha = p / den[dst_t]
    out = torch.zeros(N, num_heads, out_feats).index_add_(0, dst_t, alpha[..., None] * z[src_t])
    return out + sd[prefix + "bias"].float().view(1, num_heads, out_feats)


@torch.no_grad()
def gated_graph_conv(sd: Dict[str, torch.Tensor], prefix: str, src, dst, etypes, x: torch.Tensor,
                     out_feats: int, n_steps: int, n_etypes: int) -> torch.Tensor:
    """DGL GatedGraphConv forward as called at baselines/models/reveal/ggnn/model.py:15-16,23 and
    baselines/models/devign/model.py:15-16,35 -> [N, out]."""
    N = x.shape[0]
    src_t = torch.as_tensor(src, dtype=torch.long)
    dst_t = torch.as_tensor(dst, dtype=torch.long)
    et = torch.as_tensor(etypes, dtype=torch.long)
    if len(et) and (int(et.min()) < 0 or int(et.max()) >= n_etypes):
        raise AssertionError("edge type indices out of range [0, n_etypes)")
    h = torch.cat([x.float(), torch.zeros(N, out_feats - x.shape[1])], 1)
    W = torch.stack([sd[prefix + f"linears.{t}.weight"].float() for t in range(n_etypes)])   # [T, out, out]
    b = torch.stack([sd[prefix + f"linears.{t}.bias"].float() for t in range(n_etypes)])
    for _ in range(n_steps):
        a = torch.zeros(N, out_feats)
        for t in range(n_etypes):
            sel = et == t
            if sel.any():
                a.index_add_(0, dst_t[sel], F.linear(h[src_t[sel]], W[t], b[t]))
        gi = F.linear(a, sd[prefix + "gru.weight_ih"].float(), sd[prefix + "gru.bias_ih"].float())
        gh = F.linear(h, sd[prefix + "gru.weight_hh"].float(), sd[prefix + "gru.bias_hh"].float())
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        zg = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - zg) * n + zg * h
    return h
