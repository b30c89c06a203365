"""GATConv-free ablation classes of the fusion model, B200-native.

Interface mirrors of /root/reference/mvuld/models/GraphModel.py -- the RQ2 / RQ3 alternatives that
main_bigvul.py:126-145 keeps next to the live ``Multi_DefectModel_new_GCN``:

    Multi_DefectModel_noGraph   :306-359    image + text only, final_fc over 1024 features
    Multi_DefectModel_000       :362-430    ELU(fconly) nodes, dgl.mean_nodes, ELU(hfc(hbn))
    Multi_DefectModel_001       :433-531    ELU(fconly) nodes, 100 slots, bn_gat + fc_gat (512 -> 512), Rs_GCN x 8
    Multi_DefectModel_100       :534-615    ELU(fconly) nodes, 100 slots, fc_gat (480) | fc_bbox (32), slot mean
    Multi_DefectModel_NOGAT2    :1277-1384  fconly + hidden x 8 nodes, fc_gat | fc_bbox, Rs_GCN x 8 ("POS+GCN")

and of /root/reference/mvuld/models/new_model.py -- the live graph branch with one modality removed from the head:

    Multi_DefectModel_noFunc         :202-319   final_fc(final_fc_bn(cat(image, graph)))
    Multi_DefectModel_noGlobalImage  :81-199    final_fc(final_fc_bn(text * graph))

Same constructor ``(config, pretrained=True, attention=True)``, ``forward(g, img_embedding, func_text_embedding)``
and state-dict keys as the reference classes (modules a class declares but never runs -- ``hidden`` in _000 / _001 /
_100, ``ln_text``, ``hln``, ``hfc`` where unused -- are kept so checkpoints load with ``strict=True``).  Eval-mode
semantics as in graph_model.py: dropout off, BatchNorms folded; the ``h_func`` branch never reaches the output and is
not evaluated.  Every product runs on the kernels of the live model (tcgen05 GEMM with the ELU epilogue, unbatch/pad +
slot BatchNorm, pos branch, split-precision Rs_GCN chain, fused l2norm/mean/concat/BN/Linear head); no CPU fallback.
The classes of that file that return ``all_feats`` instead of logits (_110, _GATPOS, _011, _NOGAT) and the two with
extra position MLPs (_NOGAT3, _NOGAT4) are not built.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .graph import Graph
from .graph_model import (Multi_DefectModel_new_GCN, Rs_GCN, _bn_affine, _fold_bn_into_linear, plan_rs_gcn_chain,
                          run_rs_gcn_chain)


class _AblationBase(nn.Module):
    """Shared host logic; subclasses set NODES / READOUT / POS / GCN (oracle.fusion.VARIANT_SPECS has the same table)."""
    NODES = None          # None | "fconly" | "fconly+hidden"
    READOUT = None        # None | "mean_nodes" | "slots"
    POS = False
    GCN = False
    FC_GAT = None         # (in, out) of fc_gat when the class has one

    def __init__(self, config, pretrained=True, attention=True):
        super().__init__()
        self.num_features = 1024
        self.config = config
        self.num_classes = config.MODEL.NUM_CLASSES
        hfeat, embfeat = 512, 768
        self.fconly = nn.Linear(embfeat, hfeat)
        self.hidden = nn.ModuleList([nn.Linear(hfeat, hfeat) for _ in range(8)])
        if self.GCN:
            for k in range(1, 9):
                setattr(self, f"Rs_GCN_{k}", Rs_GCN(in_channels=512, inter_channels=512))
        self.bn_text = nn.BatchNorm1d(embfeat)
        self.ln_text = nn.LayerNorm(embfeat)
        self.fc_text = nn.Linear(embfeat, hfeat)
        if self.READOUT == "slots":
            self.max_node = 100
            self.bn_gat = nn.BatchNorm1d(self.max_node)
            self.fc_gat = nn.Linear(*self.FC_GAT)
            if self.POS:
                self.bn_bbox = nn.BatchNorm1d(self.max_node)
                self.fc_bbox = nn.Linear(4, 32)
        self.swinbn = nn.BatchNorm1d(self.num_features)
        self.swinfc = nn.Linear(self.num_features, hfeat)
        self.hbn = nn.BatchNorm1d(hfeat)
        self.hln = nn.LayerNorm(hfeat)
        self.hfc = nn.Linear(hfeat, hfeat)
        nfeat = hfeat * (2 if self.NODES is None else 3)
        self.final_fc = nn.Linear(nfeat, self.num_classes)
        self.final_fc_bn = nn.BatchNorm1d(nfeat)
        self._plan = None

    def invalidate(self):
        self._plan = None

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def prepare(self):
        dev = self.swinfc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        p = dict(dev=dev)
        w, b = _fold_bn_into_linear(self.swinbn, self.swinfc)
        p["img"] = (b16(w), f32(b))
        w, b = _fold_bn_into_linear(self.bn_text, self.fc_text)
        p["txt"] = (b16(w), f32(b))
        if self.NODES is not None:
            p["fconly"] = (b16(self.fconly.weight), f32(self.fconly.bias))
            p["hidden"] = [(b16(l.weight), f32(l.bias)) for l in self.hidden] if self.NODES == "fconly+hidden" else []
        if self.READOUT == "mean_nodes":
            w, b = _fold_bn_into_linear(self.hbn, self.hfc)
            p["h"] = (b16(w), f32(b))
        elif self.READOUT == "slots":
            s, t = _bn_affine(self.bn_gat)
            p["bn_gat"] = (f32(s), f32(t))
            p["fc_gat"] = (b16(self.fc_gat.weight), f32(self.fc_gat.bias))
            if self.POS:
                s, t = _bn_affine(self.bn_bbox)
                p["bn_bbox"] = (f32(s), f32(t))
                p["fc_bbox"] = (f32(self.fc_bbox.weight), f32(self.fc_bbox.bias))
            if self.GCN:
                p["gcn"] = plan_rs_gcn_chain([getattr(self, f"Rs_GCN_{k}") for k in range(1, 9)], dev)
        scale, shift = _bn_affine(self.final_fc_bn)
        wf = self.final_fc.weight.detach().float()
        p["final"] = (f32(wf * scale[None, :]), f32(self.final_fc.bias.detach().float() + wf @ shift))
        self._plan = p
        return self

    @torch.no_grad()
    def forward(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("mvuld_b200 fusion model implements the eval-mode forward: call model.eval()")
        if not img_embedding.is_cuda:
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        B = img_embedding.shape[0]
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32
        img_b, txt_b = e((B, 1024), bf), e((B, 768), bf)
        _lib.call("mvuld_f32_to_bf16", img_embedding.float().contiguous(), img_b, B * 1024)
        _lib.call("mvuld_f32_to_bf16", func_text_embedding.float().contiguous(), txt_b, B * 768)
        ximg, xtxt = e((B, 512), f32), e((B, 512), f32)
        _lib.gemm(img_b, p["img"][0], bias=p["img"][1], act=_lib.ACT_ELU, out_f32=ximg)
        _lib.gemm(txt_b, p["txt"][0], bias=p["txt"][1], act=_lib.ACT_ELU, out_f32=xtxt)
        logits = e((B, self.num_classes), f32)
        if self.NODES is None:                                         # GraphModel.py:356-358
            feats = torch.cat([ximg, xtxt], 1)
            _lib.call("mvuld_linear_small", feats, p["final"][0], p["final"][1], logits, None, B, self.num_classes, 1024)
            return logits

        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if g.device.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if g.batch_size != B:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {B}")
        N = g.num_nodes()
        h_in = g.ndata["_UNIX_NODE_EMB"]
        hb = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), hb, N * h_in.shape[1])
        a, a2, a32 = e((N, 512), bf), e((N, 512), bf), e((N, 512), f32)
        want32 = self.READOUT == "mean_nodes"                           # the node mean is taken on the fp32 copy
        _lib.gemm(hb, p["fconly"][0], bias=p["fconly"][1], act=_lib.ACT_ELU, out_bf16=a, out_f32=a32 if want32 else None)
        for (w, b) in p["hidden"]:
            _lib.gemm(a, w, bias=b, act=_lib.ACT_ELU, out_bf16=a2, out_f32=a32 if want32 else None)
            a, a2 = a2, a

        if self.READOUT == "mean_nodes":                               # dgl.mean_nodes + ELU(hfc(hbn(.)))
            bnn = g.batch_num_nodes().to(torch.int32)
            start = (torch.cumsum(bnn, 0, dtype=torch.int32) - bnn).to(dev)
            hmean, hmean_b = e((B, 512), f32), e((B, 512), bf)
            _lib.call("mvuld_seq_segment_mean", a32, start, bnn.to(dev), None, hmean, B, 512)
            _lib.call("mvuld_f32_to_bf16", hmean, hmean_b, B * 512)
            hfeat = e((B, 512), f32)
            _lib.gemm(hmean_b, p["h"][0], bias=p["h"][1], act=_lib.ACT_ELU, out_f32=hfeat)
            feats = torch.cat([ximg, hfeat, xtxt], 1)
            _lib.call("mvuld_linear_small", feats, p["final"][0], p["final"][1], logits, None, B, self.num_classes, 1536)
            g.check_status()
            return logits

        # slots: unbatch -> pad / truncate to max_node -> slot BN -> fc_gat (| fc_bbox)
        pos = g.ndata["pos_emb"].float().contiguous()
        g.ndata['HGATOUTPUT'] = a                                      # side effects of the reference forward (bf16 here)
        g.ndata['HFGATOUTPUT'] = pos
        n = self.max_node
        offsets = g.node_offsets()
        hp = e((B * n, 512), bf)
        _lib.call("mvuld_unbatch_pad_bn", a, offsets, p["bn_gat"][0], p["bn_gat"][1], hp, None, B, n, 512)
        z32, zb = e((B * n, 512), f32), e((B * n, 512), bf)
        _lib.gemm(hp, p["fc_gat"][0], bias=p["fc_gat"][1], act=_lib.ACT_ELU, out_bf16=zb, out_f32=z32)
        if self.POS:
            _lib.call("mvuld_pos_branch", pos, offsets, p["bn_bbox"][0], p["bn_bbox"][1], p["fc_bbox"][0],
                      p["fc_bbox"][1], z32, zb, B, n, 32, 512, 480)
        if self.GCN:                                                   # Rs_GCN x 8, l2norm over slots, mean, head
            run_rs_gcn_chain(p["gcn"], z32, B, n)
            _lib.call("mvuld_fusion_head", z32, ximg, xtxt, p["final"][0], p["final"][1], logits, None, B, n, 512,
                      self.num_classes)
        else:                                                          # plain mean over the slots (GraphModel.py:610)
            start = torch.arange(B, device=dev, dtype=torch.int32) * n
            length = torch.full((B,), n, device=dev, dtype=torch.int32)
            hfeat = e((B, 512), f32)
            _lib.call("mvuld_seq_segment_mean", z32, start, length, None, hfeat, B, 512)
            feats = torch.cat([ximg, hfeat, xtxt], 1)
            _lib.call("mvuld_linear_small", feats, p["final"][0], p["final"][1], logits, None, B, self.num_classes, 1536)
        g.check_status()
        return logits


class Multi_DefectModel_noGraph(_AblationBase):
    """GraphModel.py:306-359 (RQ2: no graph feature)."""
    NODES, READOUT = None, None


class Multi_DefectModel_000(_AblationBase):
    """GraphModel.py:362-430."""
    NODES, READOUT = "fconly", "mean_nodes"


class Multi_DefectModel_001(_AblationBase):
    """GraphModel.py:433-531."""
    NODES, READOUT, POS, GCN, FC_GAT = "fconly", "slots", False, True, (512, 512)


class Multi_DefectModel_100(_AblationBase):
    """GraphModel.py:534-615."""
    NODES, READOUT, POS, GCN, FC_GAT = "fconly", "slots", True, False, (512, 480)


class Multi_DefectModel_NOGAT2(_AblationBase):
    """GraphModel.py:1277-1384 (RQ3 "POS+GCN 101", main_bigvul.py:140)."""
    NODES, READOUT, POS, GCN, FC_GAT = "fconly+hidden", "slots", True, True, (512, 480)


class Multi_DefectModel_noFunc(Multi_DefectModel_new_GCN):
    """new_model.py:202-319 (RQ2: no function text): the live model's GATConv x2 / node MLP / slots / Rs_GCN x 8 graph
    branch; head = final_fc(final_fc_bn(cat(image, graph))) over 1024 features (one fused kernel, mode 1)."""
    HEAD_MODE, HEAD_FEATS = 1, 2


class Multi_DefectModel_noGlobalImage(Multi_DefectModel_new_GCN):
    """new_model.py:81-199 (RQ2: no global image): same graph branch; head = final_fc(final_fc_bn(text * graph)) over
    512 features (elementwise product of the text projection and the graph readout, mode 2)."""
    HEAD_MODE, HEAD_FEATS = 2, 1


ABLATIONS = {c.__name__: c for c in (Multi_DefectModel_noGraph, Multi_DefectModel_000, Multi_DefectModel_001,
                                     Multi_DefectModel_100, Multi_DefectModel_NOGAT2, Multi_DefectModel_noFunc,
                                     Multi_DefectModel_noGlobalImage)}
