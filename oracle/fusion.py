"""Oracle (test infrastructure): functional fp32 restatement of the live fusion model, Rs_GCN and
the GGNN baseline, in eval mode (dropout off, BatchNorm on running statistics).

Reference: /root/reference/mvuld/models/GraphModel.py:74-211 (Multi_DefectModel_new_GCN),
/root/reference/mvuld/models/Rs_GCN.py:52-73, /root/reference/baselines/models/reveal/ggnn/model.py:20-31.
Graph ops come from oracle.dgl_ops (DGL restated; parity unpinned, see that file).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import dgl_ops


def _bn_eval(sd, prefix, x, dim, eps=1e-5):
    """BatchNorm1d in eval mode; channel axis = ``dim``."""
    shape = [1] * x.dim()
    shape[dim] = -1
    g = lambda k: sd[prefix + k].float().view(shape)
    return (x - g("running_mean")) / torch.sqrt(g("running_var") + eps) * g("weight") + g("bias")


@torch.no_grad()
def rs_gcn(sd: Dict[str, torch.Tensor], prefix: str, v: torch.Tensor):
    """Rs_GCN.py:52-73. v: [B, D, N] -> (v* [B, D, N], R/N [B, N, N]).  No softmax."""
    conv = lambda name, t: F.conv1d(t, sd[prefix + name + ".weight"].float(), sd[prefix + name + ".bias"].float())
    g_v = conv("g", v).permute(0, 2, 1)              # [B, N, C]
    theta = conv("theta", v).permute(0, 2, 1)        # [B, N, C]
    phi = conv("phi", v)                             # [B, C, N]
    R = theta @ phi
    R = R / R.size(-1)
    y = (R @ g_v).permute(0, 2, 1).contiguous()      # [B, C, N]
    wy = _bn_eval(sd, prefix + "W.1.", conv("W.0", y), 1)
    return wy + v, R


def l2norm_dim1(x: torch.Tensor) -> torch.Tensor:
    """GraphModel.py:74-79: X / sqrt(sum X^2 over dim=1) -- no eps; dim 1 is the node-slot axis."""
    return x / torch.pow(x, 2).sum(dim=1, keepdim=True).sqrt()


@torch.no_grad()
def fusion_forward(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                   func_text_embedding: torch.Tensor, max_node: int = 100, taps: dict | None = None,
                   head: str = "all", rs_gcn_modules=None) -> torch.Tensor:
    """GraphModel.py:150-211 -> logits [B, 2].  ``head``: "all" = the live model; "noFunc" = new_model.py:317-318
    (cat(image, graph)); "noGlobalImage" = new_model.py:196-197 (text * graph) -- same graph branch.
    ``rs_gcn_modules``: eight instances of the REFERENCE ``Rs_GCN`` module (oracle/_ref) to run in place of the
    restatement ``rs_gcn`` (bench.py's CPU arm)."""
    lin = lambda name, t: F.linear(t, sd[name + ".weight"].float(), sd[name + ".bias"].float())
    x = F.elu(lin("swinfc", _bn_eval(sd, "swinbn.", img_embedding.float(), 1)))
    t = F.elu(lin("fc_text", _bn_eval(sd, "bn_text.", func_text_embedding.float(), 1)))
    h = batch.ndata["_UNIX_NODE_EMB"].float()
    pos = batch.ndata["pos_emb"].float()
    h = dgl_ops.gat_conv(sd, "gat.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    if taps is not None:
        taps["gat1"] = h.clone()
    h = dgl_ops.gat_conv(sd, "gat2.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    if taps is not None:
        taps["gat2"] = h.clone()
    h = F.elu(lin("fc", h))
    for i in range(8):
        h = F.elu(lin(f"hidden.{i}", h))
    if taps is not None:
        taps["node_mlp"] = h.clone()
    # the h_func branch (GraphModel.py:172,177) never reaches the output and is not evaluated
    h_i = dgl_ops.unbatch_pad(h, batch.batch_num_nodes, max_node)            # [B, 100, 512]
    pos_i = dgl_ops.unbatch_pad(pos, batch.batch_num_nodes, max_node)        # [B, 100, 4]
    h_i = F.elu(lin("fc_gat", _bn_eval(sd, "bn_gat.", h_i, 1)))              # BN over the slot axis
    pos_i = F.elu(lin("fc_bbox", _bn_eval(sd, "bn_bbox.", pos_i, 1)))
    z = torch.cat([h_i, pos_i], 2).permute(0, 2, 1)                           # [B, 512, 100]
    if taps is not None:
        taps["gcn_in"] = z.clone()
    for k in range(1, 9):
        z, _ = rs_gcn(sd, f"Rs_GCN_{k}.", z) if rs_gcn_modules is None else rs_gcn_modules[k - 1](z)
    if taps is not None:
        taps["gcn_out"] = z.clone()
    z = l2norm_dim1(z.permute(0, 2, 1)).mean(dim=1)                           # [B, 512]
    feats = {"all": lambda: torch.cat([x, z, t], 1), "noFunc": lambda: torch.cat([x, z], 1),
             "noGlobalImage": lambda: t * z}[head]()
    return lin("final_fc", _bn_eval(sd, "final_fc_bn.", feats, 1))


@torch.no_grad()
def gat_variant_forward(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                        func_text_embedding: torch.Tensor) -> torch.Tensor:
    """GraphModel.py:263-304 (``Multi_DefectModel``, the RQ3 GAT ablation): GATConv x2 + node MLP, dgl.mean_nodes,
    ELU(hfc(hbn(.))), concat, final_fc(final_fc_bn(.)) -> logits [B, 2].  Eval mode."""
    lin = lambda name, t: F.linear(t, sd[name + ".weight"].float(), sd[name + ".bias"].float())
    x = F.elu(lin("swinfc", _bn_eval(sd, "swinbn.", img_embedding.float(), 1)))
    t = F.elu(lin("fc_text", _bn_eval(sd, "bn_text.", func_text_embedding.float(), 1)))
    h = batch.ndata["_UNIX_NODE_EMB"].float()
    h = dgl_ops.gat_conv(sd, "gat.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    h = dgl_ops.gat_conv(sd, "gat2.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    h = F.elu(lin("fc", h))
    for i in range(8):
        h = F.elu(lin(f"hidden.{i}", h))
    hf = dgl_ops.mean_nodes(h, batch.batch_num_nodes)
    hf = F.elu(lin("hfc", _bn_eval(sd, "hbn.", hf, 1)))
    feats = torch.cat([x, hf, t], 1)
    return lin("final_fc", _bn_eval(sd, "final_fc_bn.", feats, 1))


@torch.no_grad()
def ggnn_sum_forward(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", out_dim: int, n_steps: int,
                     n_etypes: int):
    """reveal/ggnn/model.py:20-31 -> (prob [B], logit [B,1], h_i_sum [B,out], node states [N,out])."""
    feats = batch.ndata["_WORD2VEC"].float()
    et = batch.edata["_ETYPE"]
    h = dgl_ops.gated_graph_conv(sd, "ggnn.", batch.src, batch.dst, et, feats, out_dim, n_steps, n_etypes)
    s = dgl_ops.segment_sum(h, batch.batch_num_nodes)
    logit = F.linear(s, sd["classifier.weight"].float(), sd["classifier.bias"].float())
    return torch.sigmoid(logit).squeeze(-1), logit, s, h


# RQ2 / RQ3 ablation classes without GATConv (GraphModel.py; commented alternatives at main_bigvul.py:126-145).
#   nodes    : "fconly" = ELU(fconly(h)); "fconly+hidden" adds the eight ELU(hidden[i](.)) layers
#   readout  : "mean_nodes" = ELU(hfc(hbn(dgl.mean_nodes))); "slots" = unbatch/pad to 100 slots + bn_gat + fc_gat
#   pos      : slots only -- concat ELU(fc_bbox(bn_bbox(pos slots))) (fc_gat is 512 -> 480) else fc_gat is 512 -> 512
#   gcn      : slots only -- the eight Rs_GCN blocks + l2norm over the slot axis before the slot mean
VARIANT_SPECS = {
    "Multi_DefectModel_noGraph": dict(nodes=None, readout=None),                                   # :306-359
    "Multi_DefectModel_000": dict(nodes="fconly", readout="mean_nodes"),                           # :362-430
    "Multi_DefectModel_001": dict(nodes="fconly", readout="slots", pos=False, gcn=True),           # :433-531
    "Multi_DefectModel_100": dict(nodes="fconly", readout="slots", pos=True, gcn=False),           # :534-615
    "Multi_DefectModel_NOGAT2": dict(nodes="fconly+hidden", readout="slots", pos=True, gcn=True),  # :1277-1384
}


@torch.no_grad()
def ablation_forward(name: str, sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                     func_text_embedding: torch.Tensor, max_node: int = 100) -> torch.Tensor:
    """Eval-mode forward of the GATConv-free ablation classes listed in VARIANT_SPECS -> logits [B, num_classes]."""
    if name in ("Multi_DefectModel_noFunc", "Multi_DefectModel_noGlobalImage"):
        return fusion_forward(sd, batch, img_embedding, func_text_embedding, max_node, head=name.split("_")[-1])
    spec = VARIANT_SPECS[name]
    lin = lambda key, t: F.linear(t, sd[key + ".weight"].float(), sd[key + ".bias"].float())
    x = F.elu(lin("swinfc", _bn_eval(sd, "swinbn.", img_embedding.float(), 1)))
    t = F.elu(lin("fc_text", _bn_eval(sd, "bn_text.", func_text_embedding.float(), 1)))
    if spec["nodes"] is None:
        feats = torch.cat([x, t], 1)                                            # GraphModel.py:356
        return lin("final_fc", _bn_eval(sd, "final_fc_bn.", feats, 1))
    h = F.elu(lin("fconly", batch.ndata["_UNIX_NODE_EMB"].float()))
    if spec["nodes"] == "fconly+hidden":
        for i in range(8):
            h = F.elu(lin(f"hidden.{i}", h))
    if spec["readout"] == "mean_nodes":
        hf = dgl_ops.mean_nodes(h, batch.batch_num_nodes)
        hf = F.elu(lin("hfc", _bn_eval(sd, "hbn.", hf, 1)))
    else:
        h_i = dgl_ops.unbatch_pad(h, batch.batch_num_nodes, max_node)          # [B, 100, 512]
        z = F.elu(lin("fc_gat", _bn_eval(sd, "bn_gat.", h_i, 1)))
        if spec["pos"]:
            pos_i = dgl_ops.unbatch_pad(batch.ndata["pos_emb"].float(), batch.batch_num_nodes, max_node)
            z = torch.cat([z, F.elu(lin("fc_bbox", _bn_eval(sd, "bn_bbox.", pos_i, 1)))], 2)
        if spec["gcn"]:
            z = z.permute(0, 2, 1)
            for k in range(1, 9):
                z, _ = rs_gcn(sd, f"Rs_GCN_{k}.", z)
            z = l2norm_dim1(z.permute(0, 2, 1))
        hf = z.mean(dim=1)
    feats = torch.cat([x, hf, t], 1)
    return lin("final_fc", _bn_eval(sd, "final_fc_bn.", feats, 1))
