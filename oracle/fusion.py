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


# The remaining classes of GraphModel.py (RQ3 "position / GAT / GCN on-off" grid) and the gating-fusion class of
# myModels.py.  Every one returns final_fc(final_fc_bn(.)) logits (the reference names the tensor ``all_feats``).
#   pre      : "gatpos" = per-node ELU(fc_gat 768->720) | ELU(fc_bbox 4->48) in front of the GATConvs (:790-792)
#   nodes    : "gat" = GATConv x2 + fc + hidden x 8; "raw" = the 768-wide line vectors go to the slots unchanged (:1014-1019);
#              "fconly+hidden" (:1130-1136); "fconly480|pos+hidden" = cat(ELU(fconly 768->480), ELU(fc_bbox 4->32)) +
#              hidden x 8 (:1241-1246)
#   posnodes : per-node ELU(fc_bbox 4->128) + eight ELU(pos_hidden[i]) layers (:1132,1138-1139)
#   slot     : what follows bn_gat on the [B, 100, F] slots: "fc_gat" | "hfc" (:815) | "elu" (:928)
#   pos      : None | "fc_bbox" (bn_bbox + fc_bbox 4->32 on the padded boxes) | "fc_bbox2" (bn_bbox + fc_bbox2 128->32, :1148)
#   gcn      : Rs_GCN x 8 + l2norm over the slot axis before the slot mean
VARIANT_SPECS2 = {
    "Multi_DefectModel_110": dict(nodes="gat", slot="fc_gat", pos="fc_bbox", gcn=False),                       # :618-718
    "Multi_DefectModel_GATPOS": dict(pre="gatpos", nodes="gat", slot="hfc", pos=None, gcn=False),              # :721-826
    "Multi_DefectModel_011": dict(nodes="gat", slot="elu", pos=None, gcn=True),                                # :830-948
    "Multi_DefectModel_NOGAT": dict(nodes="raw", slot="fc_gat", pos="fc_bbox", gcn=True),                      # :950-1050
    "Multi_DefectModel_NOGAT3": dict(nodes="fconly+hidden", posnodes=True, slot="fc_gat", pos="fc_bbox2", gcn=True),  # :1053-1170
    "Multi_DefectModel_NOGAT4": dict(nodes="fconly480|pos+hidden", slot="fc_gat", pos=None, gcn=True),         # :1173-1273
}


@torch.no_grad()
def variant2_forward(name: str, sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                     func_text_embedding: torch.Tensor, max_node: int = 100) -> torch.Tensor:
    """Eval-mode forward of the classes listed in VARIANT_SPECS2 -> logits [B, num_classes]."""
    spec = VARIANT_SPECS2[name]
    lin = lambda key, t: F.linear(t, sd[key + ".weight"].float(), sd[key + ".bias"].float())
    x = F.elu(lin("swinfc", _bn_eval(sd, "swinbn.", img_embedding.float(), 1)))
    t = F.elu(lin("fc_text", _bn_eval(sd, "bn_text.", func_text_embedding.float(), 1)))
    h = batch.ndata["_UNIX_NODE_EMB"].float()
    pos = batch.ndata["pos_emb"].float()
    pos_n = None
    if spec.get("pre") == "gatpos":
        h = torch.cat([F.elu(lin("fc_gat", h)), F.elu(lin("fc_bbox", pos))], 1)
    if spec["nodes"] == "gat":
        h = dgl_ops.gat_conv(sd, "gat.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
        h = dgl_ops.gat_conv(sd, "gat2.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
        h = F.elu(lin("fc", h))
    elif spec["nodes"] == "fconly+hidden":
        h = F.elu(lin("fconly", h))
    elif spec["nodes"] == "fconly480|pos+hidden":
        h = torch.cat([F.elu(lin("fconly", h)), F.elu(lin("fc_bbox", pos))], 1)
    if spec["nodes"] != "raw":
        for i in range(8):
            h = F.elu(lin(f"hidden.{i}", h))
    if spec.get("posnodes"):
        pos_n = F.elu(lin("fc_bbox", pos))
        for i in range(8):
            pos_n = F.elu(lin(f"pos_hidden.{i}", pos_n))
    h_i = _bn_eval(sd, "bn_gat.", dgl_ops.unbatch_pad(h, batch.batch_num_nodes, max_node), 1)
    z = F.elu({"fc_gat": lambda: lin("fc_gat", h_i), "hfc": lambda: lin("hfc", h_i), "elu": lambda: h_i}[spec["slot"]]())
    if spec["pos"] == "fc_bbox":
        pos_i = dgl_ops.unbatch_pad(pos, batch.batch_num_nodes, max_node)
        z = torch.cat([z, F.elu(lin("fc_bbox", _bn_eval(sd, "bn_bbox.", pos_i, 1)))], 2)
    elif spec["pos"] == "fc_bbox2":
        pos_i = dgl_ops.unbatch_pad(pos_n, batch.batch_num_nodes, max_node)
        z = torch.cat([z, F.elu(lin("fc_bbox2", _bn_eval(sd, "bn_bbox.", pos_i, 1)))], 2)
    if spec["gcn"]:
        z = z.permute(0, 2, 1)
        for k in range(1, 9):
            z, _ = rs_gcn(sd, f"Rs_GCN_{k}.", z)
        z = l2norm_dim1(z.permute(0, 2, 1))
    feats = torch.cat([x, z.mean(dim=1), t], 1)
    return lin("final_fc", _bn_eval(sd, "final_fc_bn.", feats, 1))


@torch.no_grad()
def gru_last_state(sd: Dict[str, torch.Tensor], prefix: str, seq: torch.Tensor) -> torch.Tensor:
    """torch.nn.GRU(batch_first=True, 1 layer), h_0 = 0 -> hidden state after the last step [B, H] (restated; pinned
    against torch.nn.GRU in tests/test_host_logic.py)."""
    w_ih, w_hh = sd[prefix + "weight_ih_l0"].float(), sd[prefix + "weight_hh_l0"].float()
    b_ih, b_hh = sd[prefix + "bias_ih_l0"].float(), sd[prefix + "bias_hh_l0"].float()
    B, T, _ = seq.shape
    H = w_hh.shape[1]
    h = seq.new_zeros(B, H)
    gi_all = seq.float() @ w_ih.t() + b_ih
    for step in range(T):
        gi, gh = gi_all[:, step], h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        zg = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - zg) * n + zg * h
    return h


@torch.no_grad()
def gating_forward(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                   func_text_embedding: torch.Tensor, projection: str = "gru", fusion: str = "attention") -> torch.Tensor:
    """myModels.py:343-428 (``Multi_DefectModel`` of that file; as shipped: projection_layer 'gru', fusion 'attention'):
    GATConv x2 + node MLP, unbatch padded to the LONGEST graph of the batch (:430-446, no truncation), projection over the
    node axis, ELU(hfc(hbn(.))), then the tanh / softmax gate of the image vector on the graph vector (:407-413)."""
    lin = lambda name, t: F.linear(t, sd[name + ".weight"].float(), sd[name + ".bias"].float())
    x = F.elu(lin("swinfc", _bn_eval(sd, "swinbn.", img_embedding.float(), 1)))
    t = F.elu(lin("fc_text", _bn_eval(sd, "bn_text.", func_text_embedding.float(), 1)))
    h = batch.ndata["_UNIX_NODE_EMB"].float()
    h = dgl_ops.gat_conv(sd, "gat.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    h = dgl_ops.gat_conv(sd, "gat2.", batch.src, batch.dst, h, 4, 512).reshape(h.shape[0], -1)
    h = F.elu(lin("fc", h))
    for i in range(8):
        h = F.elu(lin(f"hidden.{i}", h))
    max_len = int(max(batch.batch_num_nodes))
    h_i = dgl_ops.unbatch_pad(h, batch.batch_num_nodes, max_len)                 # [B, max_len, 512]
    B = x.shape[0]
    if projection == "gru":
        hv = gru_last_state(sd, "gru_local.", h_i)
    elif projection == "attention":
        a = F.softmax(F.leaky_relu(torch.bmm(x.reshape(B, 1, 512), h_i.permute(0, 2, 1))), dim=2)
        hv = torch.bmm(a, h_i).reshape(B, -1)
    else:
        hv = h_i.mean(dim=1)
    hv = F.elu(lin("hfc", _bn_eval(sd, "hbn.", hv, 1)))
    if fusion == "attention":
        feats = torch.cat([F.softmax(torch.tanh(x * hv), dim=1) * hv, t], 1)
    elif fusion == "dot":
        feats = torch.cat([x * hv, t], 1)
    else:
        feats = torch.cat([x, hv, t], 1)
    return lin("final_fc", _bn_eval(sd, "final_bn.", feats, 1))


def class_forward(key: str, sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img_embedding: torch.Tensor,
                  func_text_embedding: torch.Tensor) -> torch.Tensor:
    """Oracle forward of the fusion class ``key`` (keys of tests/golden/fusion_classes.pt, which holds the logits of
    the reference's own classes run with ``dgl`` stubbed by oracle.dgl_ops: tools/make_golden.py fusion_classes)."""
    if key == "Multi_DefectModel_new_GCN":
        return fusion_forward(sd, batch, img_embedding, func_text_embedding)
    if key == "Multi_DefectModel":
        return gat_variant_forward(sd, batch, img_embedding, func_text_embedding)
    if key == "myModels.Multi_DefectModel":
        return gating_forward(sd, batch, img_embedding, func_text_embedding)
    if key in VARIANT_SPECS2:
        return variant2_forward(key, sd, batch, img_embedding, func_text_embedding)
    return ablation_forward(key, sd, batch, img_embedding, func_text_embedding)
