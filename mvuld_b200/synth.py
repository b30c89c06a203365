"""Seeded synthetic Big-Vul-shaped inputs and parity-mode weight randomisation.

There is no network or dataset here, so benchmarks and tests use synthetic data of the shapes the reference's data
layer produces (/root/reference/mvuld/data/data_list.py:276-314, mvuld/data/build.py:146-162,
mvuld/models/unixcoder.py:137-151).  Everything is generated on the CPU from an explicit ``torch.Generator`` so the
same seed gives the same tensors on any machine.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn

from . import graph as G

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def images(B: int, size: int = 448, seed: int = 12345) -> torch.Tensor:
    """Rendered-graph-like images: ~90 % white, dark strokes, ImageNet-normalised fp32 [B, 3, size, size]."""
    g = _gen(seed)
    white = torch.rand(B, 1, size, size, generator=g) < 0.9
    dark = torch.rand(B, 1, size, size, generator=g) * 0.3
    base = torch.where(white, torch.ones_like(dark), dark)
    x = (base + (torch.rand(B, 3, size, size, generator=g) - 0.5) * 0.04).clamp(0, 1)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def token_ids(B: int, L: int = 512, vocab: int = 51416, seed: int = 12345, pad: int = 1) -> torch.Tensor:
    """[CLS]=0, <encoder-only> placeholder id 6, [SEP]=2, tokens, [SEP]=2, PAD=1 ... (unixcoder.py:137-151)."""
    g = _gen(seed + 1)
    lens = torch.exp(torch.randn(B, generator=g) * 0.7 + 5.3).round().clamp(8, L).long()
    ids = torch.full((B, L), pad, dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        body = torch.randint(4, vocab, (n - 4,), generator=g)
        ids[b, :n] = torch.cat([torch.tensor([0, 6, 2]), body, torch.tensor([2])])
    return ids


def line_token_ids(n_lines: int, vocab: int = 51416, seed: int = 12345, pad: int = 1, L: int = 512) -> torch.Tensor:
    """Token ids of ``n_lines`` code lines as ``tokenize(..., max_length=512, padding=True)`` returns them
    (unixcoder.py:137-151): ``<s> <encoder-only> </s> tokens </s>`` padded to 512.  Line length (sub-word tokens) is
    log-normal around 12 (a statement of a C function), clipped to [1, 508]."""
    g = _gen(seed + 4)
    body = torch.exp(torch.randn(n_lines, generator=g) * 0.6 + math.log(12.0)).round().clamp(1, L - 4).long()
    ids = torch.full((n_lines, L), pad, dtype=torch.int64)
    for i in range(n_lines):
        nb = int(body[i])
        ids[i, :nb + 4] = torch.cat([torch.tensor([0, 6, 2]), torch.randint(4, vocab, (nb,), generator=g),
                                     torch.tensor([2])])
    return ids


def _num_nodes(B: int, g: torch.Generator, lo: int = 2, hi: int = 2000) -> List[int]:
    n = torch.exp(torch.randn(B, generator=g) * 0.6 + math.log(170.0)).round().clamp(lo, hi).long()
    return n.tolist()


def cpg_batch(B: int, seed: int = 12345, node_dim: int = 768) -> G.Graph:
    """Batched code-property graphs for the fusion model: typed edges {0 AST tree, 3 CFG chain, 1 CDG random}, self
    loops appended last (etype 0); ndata _UNIX_NODE_EMB, pos_emb, _FUNC_EMB."""
    g = _gen(seed + 2)
    graphs = []
    for n in _num_nodes(B, g):
        parent = (torch.rand(n - 1, generator=g) * torch.arange(1, n)).long()            # random spanning tree
        tree_src, tree_dst = parent, torch.arange(1, n)
        chain_src, chain_dst = torch.arange(0, n - 1), torch.arange(1, n)
        rnd_src = torch.randint(0, n, (n + 2,), generator=g)
        rnd_dst = torch.randint(0, n, (n + 2,), generator=g)
        src = torch.cat([tree_src, chain_src, rnd_src])
        dst = torch.cat([tree_dst, chain_dst, rnd_dst])
        gr = G.graph((src, dst), num_nodes=n)
        gr.edata["_ETYPE"] = torch.cat([torch.zeros(n - 1), torch.full((n - 1,), 3.0), torch.ones(n + 2)]).long()
        gr.ndata["_UNIX_NODE_EMB"] = torch.randn(n, node_dim, generator=g) * 0.5
        have = torch.rand(n, generator=g) < 0.7
        x0, y0 = torch.rand(n, generator=g) * 0.9, torch.rand(n, generator=g) * 0.9
        x1 = x0 + 0.02 + torch.rand(n, generator=g) * 0.08
        y1 = y0 + 0.01 + torch.rand(n, generator=g) * 0.04
        pos = torch.stack([x0, y0, x1, y1], 1) * have[:, None]
        gr.ndata["pos_emb"] = (pos * 1e5).round() / 1e5
        fvec = torch.randn(1, node_dim, generator=g) * 0.5
        gr.ndata["_FUNC_EMB"] = fvec.expand(n, node_dim).contiguous()
        graphs.append(G.add_self_loop(gr))
    return G.batch(graphs)


def ggnn_batch(B: int, seed: int = 12345, in_dim: int = 132, n_etypes: int = 4) -> G.Graph:
    """configs[2]: B graphs, E_k = 4 N_k random intra-graph typed edges + self loops (etype 0); _WORD2VEC [N, 132]
    = 32-way one-hot node type + 100-d N(0, 0.3) (baselines/models/devign/dataset.py:136-147)."""
    g = _gen(seed + 3)
    nn_ = torch.tensor(_num_nodes(B, g), dtype=torch.int64)
    off = torch.zeros(B, dtype=torch.int64)
    off[1:] = torch.cumsum(nn_, 0)[:-1]
    N = int(nn_.sum())
    # vectorised construction (819 200 nodes at B = 4096): per-graph edges then per-graph self loops, graph by graph
    gid = torch.repeat_interleave(torch.arange(B), nn_ * 4)
    src = (torch.rand(gid.numel(), generator=g) * nn_[gid]).long() + off[gid]
    dst = (torch.rand(gid.numel(), generator=g) * nn_[gid]).long() + off[gid]
    et = torch.randint(0, n_etypes, (gid.numel(),), generator=g)
    # interleave so that each graph's edge list is [4 N_k random edges..., N_k self loops] like dgl.batch would give
    e_cnt = nn_ * 4
    e_off = torch.zeros(B, dtype=torch.int64)
    e_off[1:] = torch.cumsum(e_cnt + nn_, 0)[:-1]
    E = int((e_cnt + nn_).sum())
    all_src = torch.empty(E, dtype=torch.int64)
    all_dst = torch.empty(E, dtype=torch.int64)
    all_et = torch.zeros(E, dtype=torch.int64)
    pos_rand = torch.arange(gid.numel()) - torch.repeat_interleave(torch.cumsum(e_cnt, 0) - e_cnt, e_cnt) + e_off[gid]
    all_src[pos_rand], all_dst[pos_rand], all_et[pos_rand] = src, dst, et
    nid = torch.arange(N)
    ngid = torch.repeat_interleave(torch.arange(B), nn_)
    pos_loop = nid - off[ngid] + e_off[ngid] + e_cnt[ngid]
    all_src[pos_loop], all_dst[pos_loop] = nid, nid
    out = G.Graph(all_src, all_dst, N, nn_, e_cnt + nn_)
    out.edata["_ETYPE"] = all_et
    onehot = torch.zeros(N, 32)
    onehot[nid, torch.randint(0, 32, (N,), generator=g)] = 1.0
    out.ndata["_WORD2VEC"] = torch.cat([onehot, torch.randn(N, in_dim - 32, generator=g) * 0.3], 1)
    return out


@torch.no_grad()
def randomize_for_parity(model: nn.Module, seed: int = 777) -> nn.Module:
    """Make a random-init comparison non-vacuous: the reference zero-initialises the
    res-post-norm LayerNorms (swin_transformer_v2.py:447-452) and Rs_GCN's output BatchNorm (Rs_GCN.py:33-34), which
    turns every Swin block and Rs_GCN block into the identity.  Re-randomise those, perturb logit_scale and the q/v
    biases, and give every BatchNorm non-trivial running statistics and affine parameters."""
    g = _gen(seed)
    rn = lambda t, mean, std: t.copy_(torch.randn(t.shape, generator=g) * std + mean)
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm1d):
            rn(m.running_mean, 0.0, 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            rn(m.weight, 1.0, 0.1)
            rn(m.bias, 0.0, 0.1)
        cls = type(m).__name__
        if cls == "SwinTransformerBlock":
            rn(m.norm1.weight, 1.0, 0.1)
            rn(m.norm1.bias, 0.0, 0.1)
            rn(m.norm2.weight, 1.0, 0.1)
            rn(m.norm2.bias, 0.0, 0.1)
        if cls == "WindowAttention":
            m.logit_scale.add_(torch.randn(m.logit_scale.shape, generator=g) * 0.3)
            rn(m.q_bias, 0.0, 0.02)
            rn(m.v_bias, 0.0, 0.02)
        if cls in ("GATConv",):
            rn(m.bias, 0.0, 0.05)
    for name, p in model.named_parameters():
        if name.endswith(".bias") and p.dim() == 1 and float(p.abs().sum()) == 0.0:
            rn(p, 0.0, 0.02)                      # linear biases are zero-initialised too
    if hasattr(model, "invalidate"):
        model.invalidate()
    for m in model.modules():
        if hasattr(m, "invalidate"):
            m.invalidate()
    return model
