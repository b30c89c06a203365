"""Oracle (test infrastructure): a SECOND, independently written restatement of the DGL operators on the hot path.

``oracle/dgl_ops.py`` is the checker of the CUDA graph kernels; DGL itself (``dgl-cu102==0.8.1``,
/root/reference/environment.yml:159) is absent, so that restatement is **parity unpinned**.  This file restates the
same documented semantics through different formulations -- a dense adjacency with multi-edge COUNTS for GATConv, a
plain per-edge Python loop for GatedGraphConv and for in-edge grouping, per-graph Python loops for batching -- so that
one misreading shared by the first restatement and the kernels written against it would show up as a disagreement
(tests/test_oracle_golden.py::test_second_dgl_restatement_*).  It pins nothing against DGL; it narrows what can be wrong.

Documented semantics restated (DGL 0.8 API reference):
  * GATConv:  z = fc(x) viewed [N, H, F]; e_ij = LeakyReLU(a_l . z_i + a_r . z_j) for EVERY edge i -> j (parallel edges
    count separately); alpha = softmax of e over the in-edges of j; out_j = sum_i alpha_ij z_i + bias.
    Call sites: /root/reference/mvuld/models/GraphModel.py:99-105,167-170.
  * GatedGraphConv:  h^0 = [x | 0]; each step a_v = sum over in-edges (u -> v, type t) of (W_t h_u + b_t);
    h = GRUCell(a, h).  Call sites: /root/reference/baselines/models/reveal/ggnn/model.py:15-23.
  * dgl.batch / add_self_loop / unbatch: node ids shifted by the running node count, edges concatenated graph by graph;
    self loops appended after a graph's own edges with zero edge data.  Call sites: mvuld/data/data_list.py:279,314.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


@torch.no_grad()
def gat_conv_dense(sd: Dict[str, torch.Tensor], prefix: str, src, dst, x: torch.Tensor, num_heads: int, out_feats: int,
                   negative_slope: float = 0.2) -> torch.Tensor:
    """Dense formulation: C[j, i] = number of edges i -> j; alpha[j, i] = C[j, i] exp(e[j, i]) / sum_k C[j, k] exp(e[j, k])
    (k parallel edges with the same score contribute k equal softmax terms)."""
    N = x.shape[0]
    C = torch.zeros(N, N, dtype=torch.float64)
    for s, d in zip(np.asarray(src).tolist(), np.asarray(dst).tolist()):
        C[d, s] += 1.0
    if bool((C.sum(1) == 0).any()):
        raise RuntimeError("There are 0-in-degree nodes in the graph (allow_zero_in_degree=False)")
    z = (x.double() @ sd[prefix + "fc.weight"].double().t()).view(N, num_heads, out_feats)
    el = (z * sd[prefix + "attn_l"].double()).sum(-1)                       # score part of the SOURCE node
    er = (z * sd[prefix + "attn_r"].double()).sum(-1)                       # score part of the DESTINATION node
    out = torch.zeros(N, num_heads, out_feats, dtype=torch.float64)
    for h in range(num_heads):
        e = er[:, h][:, None] + el[:, h][None, :]                           # e[j, i]
        e = torch.where(e > 0, e, negative_slope * e)
        e = e.masked_fill(C == 0, -float("inf"))
        w = C * torch.exp(e - e.max(dim=1, keepdim=True).values)
        alpha = w / w.sum(dim=1, keepdim=True)
        out[:, h] = alpha @ z[:, h]
    return (out + sd[prefix + "bias"].double().view(1, num_heads, out_feats)).float()


@torch.no_grad()
def gated_graph_conv_loop(sd: Dict[str, torch.Tensor], prefix: str, src, dst, etypes, x: torch.Tensor, out_feats: int,
                          n_steps: int, n_etypes: int) -> torch.Tensor:
    """Per-edge Python loop; the GRU cell written out gate by gate in float64."""
    N = x.shape[0]
    h = torch.zeros(N, out_feats, dtype=torch.float64)
    h[:, :x.shape[1]] = x.double()
    W = [sd[prefix + f"linears.{t}.weight"].double() for t in range(n_etypes)]
    b = [sd[prefix + f"linears.{t}.bias"].double() for t in range(n_etypes)]
    w_ih, w_hh = sd[prefix + "gru.weight_ih"].double(), sd[prefix + "gru.weight_hh"].double()
    b_ih, b_hh = sd[prefix + "gru.bias_ih"].double(), sd[prefix + "gru.bias_hh"].double()
    D = out_feats
    edges = list(zip(np.asarray(src).tolist(), np.asarray(dst).tolist(), torch.as_tensor(etypes).tolist()))
    for _ in range(n_steps):
        msg = [W[t] @ h[u] + b[t] for (u, _v, t) in edges]
        a = torch.zeros(N, D, dtype=torch.float64)
        for (_u, v, _t), m in zip(edges, msg):
            a[v] += m
        nh = torch.empty_like(h)
        for v in range(N):
            gi, gh = w_ih @ a[v] + b_ih, w_hh @ h[v] + b_hh
            r = torch.sigmoid(gi[:D] + gh[:D])
            zg = torch.sigmoid(gi[D:2 * D] + gh[D:2 * D])
            n = torch.tanh(gi[2 * D:] + r * gh[2 * D:])
            nh[v] = (1 - zg) * n + zg * h[v]
        h = nh
    return h.float()


def in_edges_loop(src, dst, num_nodes: int):
    """In-edge grouping by a Python loop: for every destination the (source, edge id) pairs in edge-id order ->
    (indptr, sources, edge ids), the CSR the kernels consume."""
    buckets = [[] for _ in range(num_nodes)]
    for eid, (s, d) in enumerate(zip(np.asarray(src).tolist(), np.asarray(dst).tolist())):
        buckets[d].append((s, eid))
    indptr, idx, eids = [0], [], []
    for bk in buckets:
        idx += [s for s, _ in bk]
        eids += [e for _, e in bk]
        indptr.append(len(idx))
    return np.asarray(indptr, np.int64), np.asarray(idx, np.int64), np.asarray(eids, np.int64)


def batch_with_self_loops_loop(graphs):
    """add_self_loop on every graph, then batch, written as one explicit loop over graphs and edges.
    graphs: sequence of (src, dst, num_nodes, etype or None) -> (src, dst, etype, batch_num_nodes, batch_num_edges)."""
    S, D, T, bnn, bne, base = [], [], [], [], [], 0
    for (src, dst, n, et) in graphs:
        src, dst = np.asarray(src).tolist(), np.asarray(dst).tolist()
        et = [0] * len(src) if et is None else torch.as_tensor(et).tolist()
        for s, d, t in zip(src, dst, et):
            S.append(s + base), D.append(d + base), T.append(t)
        for i in range(n):
            S.append(i + base), D.append(i + base), T.append(0)
        bnn.append(n), bne.append(len(src) + n)
        base += n
    return (np.asarray(S, np.int64), np.asarray(D, np.int64), np.asarray(T, np.int64), np.asarray(bnn, np.int64),
            np.asarray(bne, np.int64))


def unbatch_pad_loop(feat: torch.Tensor, batch_num_nodes, max_node: int) -> torch.Tensor:
    """GraphModel.py:30-54 as a loop: graph k's rows, truncated to max_node, zero rows after them."""
    out = torch.zeros(len(batch_num_nodes), max_node, feat.shape[1], dtype=feat.dtype)
    base = 0
    for k, n in enumerate(np.asarray(batch_num_nodes).tolist()):
        m = min(n, max_node)
        out[k, :m] = feat[base:base + m]
        base += n
    return out
