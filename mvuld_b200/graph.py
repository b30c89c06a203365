"""Minimal batched-graph container standing in for ``dgl.DGLGraph`` on the MVulD hot path.

DGL (pinned ``dgl-cu102==0.8.1``, /root/reference/environment.yml:159) is not installable here, so the fusion model
takes this container; it exposes the handful of members the reference touches on ``g``
(mvuld/models/GraphModel.py:163-181, mvuld/main_bigvul.py:310): ``ndata`` / ``edata`` dicts, ``to(device)``,
``batch_num_nodes()``, ``edges()``, ``num_nodes()``.  ``from_dgl`` adapts a real DGLGraph when DGL is present.

Construction follows DGL semantics: ``graph`` keeps edge order (edge id = position),
``add_self_loop`` appends (i, i) after all edges and zero-fills new edata rows, ``batch`` shifts node ids by the running
node count and concatenates edges graph by graph.  These are host-side (numpy / torch CPU) like the reference's data
loader; the in-edge CSR the kernels need is built on the device (``mvuld_csr_from_coo``) and cached on the container.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib


class Graph:
    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: Optional[int] = None,
                 batch_num_nodes: Optional[torch.Tensor] = None, batch_num_edges: Optional[torch.Tensor] = None):
        src = torch.as_tensor(src, dtype=torch.int64)
        dst = torch.as_tensor(dst, dtype=torch.int64)
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src and dst must be 1-D tensors of equal length")
        if num_nodes is None:
            num_nodes = int(max(src.max().item(), dst.max().item()) + 1) if src.numel() else 0
        self._src, self._dst = src, dst
        self._num_nodes = int(num_nodes)
        self._bnn = batch_num_nodes if batch_num_nodes is not None else torch.tensor([self._num_nodes], dtype=torch.int64)
        self._bne = batch_num_edges if batch_num_edges is not None else torch.tensor([src.numel()], dtype=torch.int64)
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self._csr = None
        self._ocsr = None
        self._offsets = None

    # ---- the DGLGraph members the reference uses ----
    def num_nodes(self) -> int:
        return self._num_nodes

    number_of_nodes = num_nodes

    def num_edges(self) -> int:
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def batch_num_nodes(self) -> torch.Tensor:
        return self._bnn

    def batch_num_edges(self) -> torch.Tensor:
        return self._bne

    @property
    def batch_size(self) -> int:
        return int(self._bnn.numel())

    @property
    def device(self):
        return self._src.device

    def to(self, device, non_blocking: bool = False) -> "Graph":
        g = Graph.__new__(Graph)
        g._src = self._src.to(device, non_blocking=non_blocking)
        g._dst = self._dst.to(device, non_blocking=non_blocking)
        g._num_nodes = self._num_nodes
        g._bnn, g._bne = self._bnn, self._bne      # host-side metadata, as in DGL
        g.ndata = {k: v.to(device, non_blocking=non_blocking) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device, non_blocking=non_blocking) for k, v in self.edata.items()}
        g._csr = None
        g._ocsr = None
        g._offsets = None
        return g

    # ---- device-side derived structures ----
    def node_offsets(self) -> torch.Tensor:
        """int64 [B+1] on the graph's device: graph k owns node rows [off[k], off[k+1])."""
        if self._offsets is None or self._offsets.device != self.device:
            off = torch.zeros(self._bnn.numel() + 1, dtype=torch.int64)
            off[1:] = torch.cumsum(self._bnn.to(torch.int64), 0)
            self._offsets = off.to(self.device)
        return self._offsets

    def in_csr(self):
        """(indptr int32 [N+1], idx_src int32 [E], eids int32 [E]) sorted by (dst, edge id), built on the GPU."""
        if self._csr is None:
            indptr, idx_src, eids, status = _lib.csr_from_coo(self._src, self._dst, self._num_nodes)
            self._csr = (indptr, idx_src, eids, status)
        return self._csr[:3]

    def out_csr(self):
        """Out-edge CSR for the backward pass of GATConv: (out_indptr int32 [N+1], out_dst int32 [E], pos_in int32 [E])
        sorted by (src, in-CSR position); ``pos_in[k]`` is the in-CSR position of out-edge ``k`` (so per-edge values
        stored in in-CSR order can be gathered per source)."""
        if getattr(self, "_ocsr", None) is None:
            _, idx_src, eids = self.in_csr()
            dst_sorted = self._dst[eids.long()]                 # destination of each in-CSR position
            out_indptr, out_dst, pos_in, _ = _lib.csr_from_coo(dst_sorted, idx_src.long(), self._num_nodes)
            self._ocsr = (out_indptr, out_dst, pos_in)
        return self._ocsr

    def check_status(self):
        """Raise if the CSR build saw an endpoint outside [0, N) (one small D2H read)."""
        if self._csr is not None and int(self._csr[3].item()) != 0:
            raise ValueError("edge endpoint outside [0, num_nodes)")
        st = getattr(self, "_collate_status", None)
        if st is not None and int(st.item()) != 0:
            raise ValueError("batch_device: a local node id lies outside its graph")


def graph(edges, num_nodes: Optional[int] = None) -> Graph:
    """``dgl.graph((src, dst))`` (mvuld/data/data_list.py:279)."""
    src, dst = edges
    return Graph(src, dst, num_nodes)


def add_self_loop(g: Graph) -> Graph:
    """``dgl.add_self_loop`` (mvuld/data/data_list.py:314): (i, i) appended after all edges, edata zero-filled."""
    n = g.num_nodes()
    loop = torch.arange(n, dtype=torch.int64, device=g.device)
    out = Graph(torch.cat([g._src, loop]), torch.cat([g._dst, loop]), n)
    out.ndata = dict(g.ndata)
    for k, v in g.edata.items():
        out.edata[k] = torch.cat([v, torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)], 0)
    return out


def batch(graphs: Sequence[Graph]) -> Graph:
    """``dgl.batch`` as GraphDataLoader's collate applies it (mvuld/data/bigvul_dataset.py:177-205)."""
    if not graphs:
        raise ValueError("cannot batch an empty list of graphs")
    bnn = torch.tensor([g.num_nodes() for g in graphs], dtype=torch.int64)
    bne = torch.tensor([g.num_edges() for g in graphs], dtype=torch.int64)
    off = torch.zeros(len(graphs), dtype=torch.int64)
    off[1:] = torch.cumsum(bnn, 0)[:-1]
    src = torch.cat([g._src + o for g, o in zip(graphs, off.tolist())])
    dst = torch.cat([g._dst + o for g, o in zip(graphs, off.tolist())])
    out = Graph(src, dst, int(bnn.sum()), bnn, bne)
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return out


def batch_device(graphs: Sequence[Graph], device, add_self_loops: bool = False, pin: bool = True) -> Graph:
    """``dgl.batch`` (and optionally ``dgl.add_self_loop`` per graph first) with the index arithmetic on the GPU.

    The host only concatenates the graphs' raw LOCAL edge lists (int32) and node data; one kernel
    (``mvuld_collate_edges``) shifts node ids by the running node count, interleaves each graph's self loops after its
    own edges and zero-fills their edge data -- the edge order DGL produces, bit-exact with
    ``batch([add_self_loop(g) for g in graphs])``.  ``ndata`` / other ``edata`` are concatenated on the host
    (self-loop rows of extra edata are zero-filled there) and copied."""
    import numpy as np
    if not graphs:
        raise ValueError("cannot batch an empty list of graphs")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("batch_device collates on a CUDA device (no CPU fallback); use graph.batch on the host")
    B = len(graphs)
    bnn = torch.tensor([g.num_nodes() for g in graphs], dtype=torch.int64)
    bne_raw = torch.tensor([g.num_edges() for g in graphs], dtype=torch.int64)
    node_off = torch.zeros(B + 1, dtype=torch.int64)
    node_off[1:] = torch.cumsum(bnn, 0)
    edge_off = torch.zeros(B + 1, dtype=torch.int64)
    edge_off[1:] = torch.cumsum(bne_raw, 0)
    N, E_raw = int(node_off[-1]), int(edge_off[-1])
    total = E_raw + (N if add_self_loops else 0)
    stage = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    up = lambda t: stage(t).to(device, non_blocking=True)
    src_l = up(torch.cat([g._src for g in graphs]).to(torch.int32))
    dst_l = up(torch.cat([g._dst for g in graphs]).to(torch.int32))
    has_et = "_ETYPE" in graphs[0].edata
    et_l = up(torch.cat([g.edata["_ETYPE"] for g in graphs]).to(torch.int64)) if has_et else None
    src = torch.empty(total, dtype=torch.int64, device=device)
    dst = torch.empty(total, dtype=torch.int64, device=device)
    et = torch.empty(total, dtype=torch.int64, device=device) if has_et else None
    status = torch.zeros(1, dtype=torch.int32, device=device)
    _lib.call("mvuld_collate_edges", src_l, dst_l, et_l, up(edge_off), up(node_off), B, 1 if add_self_loops else 0,
              src, dst, et, total, status)
    out = Graph.__new__(Graph)
    out._src, out._dst, out._num_nodes = src, dst, N
    out._bnn = bnn
    out._bne = bne_raw + (bnn if add_self_loops else 0)
    out.ndata = {k: up(torch.cat([g.ndata[k] for g in graphs], 0)) for k in graphs[0].ndata}
    out.edata = {"_ETYPE": et} if has_et else {}
    for k in graphs[0].edata:
        if k == "_ETYPE":
            continue
        parts = []
        for g in graphs:
            v = g.edata[k]
            parts.append(v)
            if add_self_loops:
                parts.append(torch.zeros((g.num_nodes(),) + tuple(v.shape[1:]), dtype=v.dtype))
        out.edata[k] = up(torch.cat(parts, 0))
    out._csr = out._ocsr = out._offsets = None
    out._collate_status = status
    return out


def from_dgl(g) -> Graph:
    """Adapter for a real (batched) ``dgl.DGLGraph`` when DGL is installed."""
    src, dst = g.edges()
    out = Graph(src.to(torch.int64), dst.to(torch.int64), g.num_nodes(), g.batch_num_nodes().cpu().to(torch.int64),
                g.batch_num_edges().cpu().to(torch.int64))
    out.ndata = {k: v for k, v in g.ndata.items()}
    out.edata = {k: v for k, v in g.edata.items()}
    return out
