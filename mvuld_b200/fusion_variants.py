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
The remaining classes of GraphModel.py -- the RQ3 "position / GAT / GCN on-off" grid -- are the table-driven
``_GridBase`` family below:

    Multi_DefectModel_110      :618-718    GATConv x2 + node MLP, 100 slots, fc_gat (480) | fc_bbox (32), slot mean
    Multi_DefectModel_GATPOS   :721-826    per-node ELU(fc_gat 768->720) | ELU(fc_bbox 4->48) in front of the GATConvs,
                                           100 slots, ELU(hfc(bn_gat(.))), slot mean
    Multi_DefectModel_011      :830-948    GATConv x2 + node MLP, 100 slots, ELU(bn_gat(.)), Rs_GCN x 8
    Multi_DefectModel_NOGAT    :950-1050   raw 768-wide line vectors into the slots, fc_gat (768->480) | fc_bbox, Rs_GCN x 8
    Multi_DefectModel_NOGAT3   :1053-1170  fconly + hidden x 8 nodes, a second per-node MLP on the boxes (fc_bbox 4->128,
                                           pos_hidden x 8), fc_gat (480) | fc_bbox2 (128->32), Rs_GCN x 8
    Multi_DefectModel_NOGAT4   :1173-1273  cat(ELU(fconly 768->480), ELU(fc_bbox 4->32)) + hidden x 8, fc_gat (512), Rs_GCN x 8

(the reference names their return value ``all_feats``; it is ``final_fc(final_fc_bn(.))``, i.e. logits.)  The gating-
fusion class of myModels.py lives in ``mvuld_b200/my_models.py``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .graph import Graph
from .graph_model import (GATConv, Multi_DefectModel_new_GCN, Rs_GCN, _bn_affine, _fold_bn_into_linear,
                          plan_rs_gcn_chain, run_rs_gcn_chain)


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
            _lib.call("mvuld_fusion_head_mode", z32, ximg, xtxt, p["final"][0], p["final"][1], logits, None, B, n, 512,
                      self.num_classes, 0)
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


def plan_gat(m: GATConv, dev):
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    return dict(w=m.fc.weight.detach().to(device=dev, dtype=torch.bfloat16).contiguous(), al=f32(m.attn_l.view(-1)),
                ar=f32(m.attn_r.view(-1)), bias=f32(m.bias), H=m._heads, F=m._out, slope=float(m.negative_slope))


def run_gat_nodes(p: dict, g: Graph, hb: torch.Tensor) -> torch.Tensor:
    """GATConv x2 -> ELU(fc) -> 8 x ELU(hidden[i]) on bf16 node features (GraphModel.py:163-177) -> bf16 [N, 512]."""
    dev, N = hb.device, hb.shape[0]
    e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
    bf, f32 = torch.bfloat16, torch.float32
    indptr, idx_src, _ = g.in_csr()
    zero_deg = torch.zeros(1, device=dev, dtype=torch.int32)
    for name in ("gat", "gat2"):
        gp = p[name]
        H, F = gp["H"], gp["F"]
        z = e((N, H * F), bf)
        _lib.gemm(hb, gp["w"], out_bf16=z)
        el, er = e((N, H), f32), e((N, H), f32)
        _lib.call("mvuld_gat_scores", z, gp["al"], gp["ar"], el, er, N, H, F)
        hb = e((N, H * F), bf)
        _lib.call("mvuld_gat_aggregate", z, el, er, indptr, idx_src, gp["bias"], hb, N, H, F, gp["slope"], zero_deg)
    a, a2 = e((N, 512), bf), e((N, 512), bf)
    _lib.gemm(hb, p["fc"][0], bias=p["fc"][1], act=_lib.ACT_ELU, out_bf16=a)
    for (w, b) in p["hidden"]:
        _lib.gemm(a, w, bias=b, act=_lib.ACT_ELU, out_bf16=a2)
        a, a2 = a2, a
    if int(zero_deg.item()):
        raise RuntimeError("There are 0-in-degree nodes in the graph (DGL GATConv raises here; add self loops)")
    return a


class _GridBase(nn.Module):
    """Shared host logic of the RQ3 grid classes; oracle.fusion.VARIANT_SPECS2 holds the same table."""
    PRE = None            # None | "gatpos"
    NODES = "gat"         # "gat" | "raw" | "fconly+hidden" | "fconly480|pos+hidden"
    POSNODES = False      # NOGAT3: per-node MLP on the boxes
    SLOT = "fc_gat"       # "fc_gat" | "hfc" | "elu"
    POS = None            # None | "fc_bbox" | "fc_bbox2"
    GCN = False
    FC_GAT = None         # (in, out) when declared
    FC_BBOX = None        # (in, out) when declared
    FCONLY_OUT = 512
    EXTRA = ()            # declared-but-unused modules kept for strict state-dict loading: "hbn_hfc", "ln", "fconly"

    def __init__(self, config, pretrained=True, attention=True):
        super().__init__()
        self.num_features = 1024
        self.config = config
        self.num_classes = config.MODEL.NUM_CLASSES
        hfeat, embfeat, numheads = 512, 768, 4
        if self.NODES == "gat":
            self.gat = GATConv(embfeat, hfeat, numheads, feat_drop=0.1)
            self.gat2 = GATConv(hfeat * numheads, hfeat, numheads, feat_drop=0.1)
            self.fc = nn.Linear(hfeat * numheads, hfeat)
        if self.NODES != "raw":
            self.fconly = nn.Linear(embfeat, self.FCONLY_OUT)
            self.hidden = nn.ModuleList([nn.Linear(hfeat, hfeat) for _ in range(8)])
        if self.POSNODES:
            self.pos_hidden = nn.ModuleList([nn.Linear(128, 128) for _ in range(8)])
        if self.GCN:
            for k in range(1, 9):
                setattr(self, f"Rs_GCN_{k}", Rs_GCN(in_channels=512, inter_channels=512))
        self.bn_text = nn.BatchNorm1d(embfeat)
        if "ln" in self.EXTRA:
            self.ln_text = nn.LayerNorm(embfeat)
        self.fc_text = nn.Linear(embfeat, hfeat)
        self.max_node = 100
        self.bn_gat = nn.BatchNorm1d(self.max_node)
        if self.FC_GAT:
            self.fc_gat = nn.Linear(*self.FC_GAT)
        if self.POS is not None or self.PRE == "gatpos":
            self.bn_bbox = nn.BatchNorm1d(self.max_node)
        if self.FC_BBOX:
            self.fc_bbox = nn.Linear(*self.FC_BBOX)
        if self.POS == "fc_bbox2":
            self.fc_bbox2 = nn.Linear(128, 32)
        self.swinbn = nn.BatchNorm1d(self.num_features)
        self.swinfc = nn.Linear(self.num_features, hfeat)
        if "hbn_hfc" in self.EXTRA or self.SLOT == "hfc":
            self.hbn = nn.BatchNorm1d(hfeat)
            if "ln" in self.EXTRA:
                self.hln = nn.LayerNorm(hfeat)
            self.hfc = nn.Linear(hfeat, hfeat)
        self.final_fc = nn.Linear(hfeat * 3, self.num_classes)
        self.final_fc_bn = nn.BatchNorm1d(hfeat * 3)
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
        lin16 = lambda m: (b16(m.weight), f32(m.bias))
        p = dict(dev=dev)
        w, b = _fold_bn_into_linear(self.swinbn, self.swinfc)
        p["img"] = (b16(w), f32(b))
        w, b = _fold_bn_into_linear(self.bn_text, self.fc_text)
        p["txt"] = (b16(w), f32(b))
        if self.NODES == "gat":
            p["gat"], p["gat2"] = plan_gat(self.gat, dev), plan_gat(self.gat2, dev)
            p["fc"] = lin16(self.fc)
        if self.NODES in ("fconly+hidden", "fconly480|pos+hidden"):
            p["fconly"] = lin16(self.fconly)
        if self.NODES != "raw":
            p["hidden"] = [lin16(l) for l in self.hidden]
        if self.POSNODES:
            p["pos_hidden"] = [lin16(l) for l in self.pos_hidden]
        s, t = _bn_affine(self.bn_gat)
        p["bn_gat"] = (f32(s), f32(t))
        if self.SLOT == "fc_gat" or self.PRE == "gatpos":
            p["fc_gat"] = lin16(self.fc_gat)
        if self.SLOT == "hfc":
            p["hfc"] = lin16(self.hfc)
        if self.FC_BBOX:
            p["fc_bbox"] = (f32(self.fc_bbox.weight), f32(self.fc_bbox.bias))
        if self.POS is not None:
            s, t = _bn_affine(self.bn_bbox)
            p["bn_bbox"] = (f32(s), f32(t))
        if self.POS == "fc_bbox2":
            p["fc_bbox2"] = lin16(self.fc_bbox2)
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
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if g.device.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        B, N, n = img_embedding.shape[0], g.num_nodes(), self.max_node
        if g.batch_size != B:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {B}")
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32
        img_b, txt_b = e((B, 1024), bf), e((B, 768), bf)
        _lib.call("mvuld_f32_to_bf16", img_embedding.float().contiguous(), img_b, B * 1024)
        _lib.call("mvuld_f32_to_bf16", func_text_embedding.float().contiguous(), txt_b, B * 768)
        ximg, xtxt = e((B, 512), f32), e((B, 512), f32)
        _lib.gemm(img_b, p["img"][0], bias=p["img"][1], act=_lib.ACT_ELU, out_f32=ximg)
        _lib.gemm(txt_b, p["txt"][0], bias=p["txt"][1], act=_lib.ACT_ELU, out_f32=xtxt)

        h_in = g.ndata["_UNIX_NODE_EMB"]
        pos = g.ndata["pos_emb"].float().contiguous()
        hb = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), hb, N * h_in.shape[1])
        if self.PRE == "gatpos":                                       # GraphModel.py:790-792
            wg, bg = p["fc_gat"]
            cat = e((N, wg.shape[0] + self.FC_BBOX[1]), bf)
            _lib.gemm(hb, wg, bias=bg, act=_lib.ACT_ELU, out_bf16=cat[:, :wg.shape[0]])
            _lib.call("mvuld_node_linear4", pos, p["fc_bbox"][0], p["fc_bbox"][1], cat, N, self.FC_BBOX[1],
                      cat.shape[1], wg.shape[0])
            hb = cat
        if self.NODES == "gat":
            a = run_gat_nodes(p, g, hb)
        elif self.NODES == "raw":
            a = hb
        else:
            a, a2 = e((N, 512), bf), e((N, 512), bf)
            wo, bo = p["fconly"]
            _lib.gemm(hb, wo, bias=bo, act=_lib.ACT_ELU, out_bf16=a[:, :wo.shape[0]])
            if self.NODES == "fconly480|pos+hidden":                   # GraphModel.py:1241-1243
                _lib.call("mvuld_node_linear4", pos, p["fc_bbox"][0], p["fc_bbox"][1], a, N, self.FC_BBOX[1], 512,
                          wo.shape[0])
            for (w, b) in p["hidden"]:
                _lib.gemm(a, w, bias=b, act=_lib.ACT_ELU, out_bf16=a2)
                a, a2 = a2, a
        g.ndata['HGATOUTPUT'] = a                                      # side effects of the reference forward (bf16 here)
        g.ndata['HFGATOUTPUT'] = pos
        offsets = g.node_offsets()
        F_in = a.shape[1]
        z32, zb = e((B * n, 512), f32), e((B * n, 512), bf)
        if self.SLOT == "elu":                                         # GraphModel.py:928
            _lib.call("mvuld_unbatch_pad_bn_elu", a, offsets, p["bn_gat"][0], p["bn_gat"][1], z32, zb, B, n, F_in)
        else:
            hp = e((B * n, F_in), bf)
            _lib.call("mvuld_unbatch_pad_bn", a, offsets, p["bn_gat"][0], p["bn_gat"][1], hp, None, B, n, F_in)
            w, b = p["fc_gat"] if self.SLOT == "fc_gat" else p["hfc"]
            _lib.gemm(hp, w, bias=b, act=_lib.ACT_ELU, out_bf16=zb[:, :w.shape[0]], out_f32=z32[:, :w.shape[0]])
        if self.POS == "fc_bbox":
            _lib.call("mvuld_pos_branch", pos, offsets, p["bn_bbox"][0], p["bn_bbox"][1], p["fc_bbox"][0],
                      p["fc_bbox"][1], z32, zb, B, n, 32, 512, 480)
        elif self.POS == "fc_bbox2":                                   # GraphModel.py:1132,1138-1139,1148
            pn, pn2 = e((N, 128), bf), e((N, 128), bf)
            _lib.call("mvuld_node_linear4", pos, p["fc_bbox"][0], p["fc_bbox"][1], pn, N, 128, 128, 0)
            for (w, b) in p["pos_hidden"]:
                _lib.gemm(pn, w, bias=b, act=_lib.ACT_ELU, out_bf16=pn2)
                pn, pn2 = pn2, pn
            pp = e((B * n, 128), bf)
            _lib.call("mvuld_unbatch_pad_bn", pn, offsets, p["bn_bbox"][0], p["bn_bbox"][1], pp, None, B, n, 128)
            w, b = p["fc_bbox2"]
            _lib.gemm(pp, w, bias=b, act=_lib.ACT_ELU, out_bf16=zb[:, 480:], out_f32=z32[:, 480:])
        logits = e((B, self.num_classes), f32)
        if self.GCN:
            run_rs_gcn_chain(p["gcn"], z32, B, n)
            _lib.call("mvuld_fusion_head_mode", z32, ximg, xtxt, p["final"][0], p["final"][1], logits, None, B, n, 512,
                      self.num_classes, 0)
        else:
            start = torch.arange(B, device=dev, dtype=torch.int32) * n
            length = torch.full((B,), n, device=dev, dtype=torch.int32)
            hfeat = e((B, 512), f32)
            _lib.call("mvuld_seq_segment_mean", z32, start, length, None, hfeat, B, 512)
            feats = torch.cat([ximg, hfeat, xtxt], 1)
            _lib.call("mvuld_linear_small", feats, p["final"][0], p["final"][1], logits, None, B, self.num_classes, 1536)
        g.check_status()
        return logits


class Multi_DefectModel_110(_GridBase):
    """GraphModel.py:618-718."""
    NODES, SLOT, POS, GCN, FC_GAT, FC_BBOX = "gat", "fc_gat", "fc_bbox", False, (512, 480), (4, 32)


class Multi_DefectModel_GATPOS(_GridBase):
    """GraphModel.py:721-826 (RQ3: positions fed to the GATConvs)."""
    PRE, NODES, SLOT, POS, GCN, FC_GAT, FC_BBOX = "gatpos", "gat", "hfc", None, False, (768, 720), (4, 48)


class Multi_DefectModel_011(_GridBase):
    """GraphModel.py:830-948."""
    NODES, SLOT, POS, GCN, FC_GAT, EXTRA = "gat", "elu", None, True, (512, 512), ("hbn_hfc", "ln")


class Multi_DefectModel_NOGAT(_GridBase):
    """GraphModel.py:950-1050."""
    NODES, SLOT, POS, GCN, FC_GAT, FC_BBOX, EXTRA = "raw", "fc_gat", "fc_bbox", True, (768, 480), (4, 32), ("hbn_hfc", "ln")


class Multi_DefectModel_NOGAT3(_GridBase):
    """GraphModel.py:1053-1170."""
    NODES, POSNODES, SLOT, POS, GCN = "fconly+hidden", True, "fc_gat", "fc_bbox2", True
    FC_GAT, FC_BBOX, EXTRA = (512, 480), (4, 128), ("hbn_hfc", "ln")


class Multi_DefectModel_NOGAT4(_GridBase):
    """GraphModel.py:1173-1273."""
    NODES, SLOT, POS, GCN, FC_GAT, FC_BBOX, FCONLY_OUT = "fconly480|pos+hidden", "fc_gat", None, True, (512, 512), (4, 32), 480
    EXTRA = ("hbn_hfc", "ln")


GRID_VARIANTS = {c.__name__: c for c in (Multi_DefectModel_110, Multi_DefectModel_GATPOS, Multi_DefectModel_011,
                                         Multi_DefectModel_NOGAT, Multi_DefectModel_NOGAT3, Multi_DefectModel_NOGAT4)}

ABLATIONS = {c.__name__: c for c in (Multi_DefectModel_noGraph, Multi_DefectModel_000, Multi_DefectModel_001,
                                     Multi_DefectModel_100, Multi_DefectModel_NOGAT2, Multi_DefectModel_noFunc,
                                     Multi_DefectModel_noGlobalImage)}
