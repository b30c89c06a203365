"""The gating-fusion class of /root/reference/mvuld/models/myModels.py:280-446 (first citation),
B200-native.

``Multi_DefectModel`` of that file, as shipped (``projection_layer = 'gru'``, ``fusion = 'attention'``, :321-322):

    x   = ELU(swinfc(swinbn(image)))                      t = ELU(fc_text(bn_text(text)))                 (:345-350)
    h   = GATConv x2 -> ELU(fc) -> 8 x ELU(hidden[i])     (the h_func copy never reaches the output)      (:361-375)
    h_i = unbatch, zero-padded to the LONGEST graph of the batch (no truncation)                          (:377-382,430-446)
    h_i = last hidden state of nn.GRU(512, 512, 1, batch_first=True) over the node axis, padding included (:385-388)
    h_i = ELU(hfc(hbn(h_i)))                                                                             (:397)
    out = final_fc(final_bn(cat(softmax(tanh(x * h_i), dim=1) * h_i, t)))                                (:402-413)

``projection_layer`` 'attention' / 'mean' and ``fusion`` 'dot' / 'concat' (the other branches of the same forward) are
selectable through the same attributes.  Same constructor, ``forward`` signature and state-dict keys as the reference
class.  Eval-mode semantics (dropout off, BatchNorms folded).  Products on the tcgen05 GEMM; the recurrence is one
cooperative kernel (``mvuld_gru_sequence``) with W_hh resident in shared memory in fp32; no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .fusion_variants import plan_gat, run_gat_nodes
from .graph import Graph
from .graph_model import GATConv, _bn_affine, _fold_bn_into_linear


class Multi_DefectModel(nn.Module):
    def __init__(self, config, pretrained=True, attention=True):
        super().__init__()
        self.num_features = 1024
        self.config = config
        self.num_classes = config.MODEL.NUM_CLASSES
        hfeat, embfeat, numheads = 512, 768, 4
        self.gat = GATConv(embfeat, hfeat, numheads, feat_drop=0.2)
        self.gat2 = GATConv(hfeat * numheads, hfeat, numheads, feat_drop=0.2)
        self.fc = nn.Linear(hfeat * numheads, hfeat)
        self.fconly = nn.Linear(embfeat, hfeat)
        self.hidden = nn.ModuleList([nn.Linear(hfeat, hfeat) for _ in range(8)])
        self.bn_text = nn.BatchNorm1d(embfeat)
        self.fc_text = nn.Linear(embfeat, hfeat)
        self.swinbn = nn.BatchNorm1d(self.num_features)
        self.swinfc = nn.Linear(self.num_features, hfeat)
        self.hbn = nn.BatchNorm1d(hfeat)
        self.hfc = nn.Linear(hfeat, hfeat)
        self.projection_layer = 'gru'      # 'gru' | 'attention' | 'mean'
        self.fusion = 'attention'          # 'attention' | 'dot' | 'concat'
        self.gru_local = nn.GRU(hfeat, hfeat, 1, batch_first=True)
        nfeat = hfeat * (3 if self.fusion == 'concat' else 2)
        self.final_bn = nn.BatchNorm1d(nfeat)
        self.final_fc = nn.Linear(nfeat, self.num_classes)
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
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        p = dict(dev=dev, projection=self.projection_layer, fusion=self.fusion)
        for key, bn, lin in (("img", self.swinbn, self.swinfc), ("txt", self.bn_text, self.fc_text),
                             ("h", self.hbn, self.hfc)):
            w, b = _fold_bn_into_linear(bn, lin)
            p[key] = (b16(w), f32(b))
        p["gat"], p["gat2"] = plan_gat(self.gat, dev), plan_gat(self.gat2, dev)
        p["fc"] = (b16(self.fc.weight), f32(self.fc.bias))
        p["hidden"] = [(b16(l.weight), f32(l.bias)) for l in self.hidden]
        g = self.gru_local
        p["gru"] = dict(w_ih=b16(g.weight_ih_l0), b_ih=f32(g.bias_ih_l0), w_hh=f32(g.weight_hh_l0), b_hh=f32(g.bias_hh_l0))
        scale, shift = _bn_affine(self.final_bn)
        wf = self.final_fc.weight.detach().float()
        p["final"] = (f32(wf * scale[None, :]), f32(self.final_fc.bias.detach().float() + wf @ shift))
        self._plan = p
        return self

    @torch.no_grad()
    def forward(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
        """myModels.py:343-428."""
        if self.training:
            raise RuntimeError("mvuld_b200 fusion model implements the eval-mode forward: call model.eval()")
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if not img_embedding.is_cuda or g.device.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if self._plan is None or self._plan["projection"] != self.projection_layer or self._plan["fusion"] != self.fusion:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        B, N = img_embedding.shape[0], g.num_nodes()
        if g.batch_size != B:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {B}")
        if p["fusion"] not in ("attention", "dot", "concat") or p["projection"] not in ("gru", "attention", "mean"):
            raise ValueError("Error: Last Layer Fusion selected not implemented")           # myModels.py:340
        if self.final_fc.weight.shape[1] != 512 * (3 if p["fusion"] == "concat" else 2):
            raise ValueError("final_bn / final_fc were built for another `fusion` setting (myModels.py:328-338)")
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32
        img_b, txt_b = e((B, 1024), bf), e((B, 768), bf)
        _lib.call("mvuld_f32_to_bf16", img_embedding.float().contiguous(), img_b, B * 1024)
        _lib.call("mvuld_f32_to_bf16", func_text_embedding.float().contiguous(), txt_b, B * 768)
        ximg, xtxt = e((B, 512), f32), e((B, 512), f32)
        _lib.gemm(img_b, p["img"][0], bias=p["img"][1], act=_lib.ACT_ELU, out_f32=ximg)
        _lib.gemm(txt_b, p["txt"][0], bias=p["txt"][1], act=_lib.ACT_ELU, out_f32=xtxt)
        h_in = g.ndata["_UNIX_NODE_EMB"]
        hb = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), hb, N * h_in.shape[1])
        a = run_gat_nodes(p, g, hb)
        g.ndata['HGATOUTPUT'] = a
        # unbatch -> zero-pad to the longest graph of the batch (myModels.py:430-446): identity affine, no truncation
        T = int(g.batch_num_nodes().max())
        ones, zeros = torch.ones(T, device=dev, dtype=f32), torch.zeros(T, device=dev, dtype=f32)
        hp = e((B * T, 512), bf)
        _lib.call("mvuld_unbatch_pad_bn", a, g.node_offsets(), ones, zeros, hp, None, B, T, 512)
        hv = e((B, 512), f32)
        if p["projection"] == "gru":
            gp = p["gru"]
            gi = e((B * T, 1536), f32)
            _lib.gemm(hp, gp["w_ih"], bias=gp["b_ih"], out_f32=gi)
            ws = torch.empty(int(_lib.load().mvuld_gru_sequence_workspace(B, 512)), device=dev, dtype=torch.uint8)
            _lib.call("mvuld_gru_sequence", gi, gp["w_hh"], gp["b_hh"], hv, ws, B, T, 512)
        elif p["projection"] == "mean":
            start = torch.arange(B, device=dev, dtype=torch.int32) * T
            hp32 = hp.float()
            _lib.call("mvuld_seq_segment_mean", hp32, start, torch.full((B,), T, device=dev, dtype=torch.int32), None,
                      hv, B, 512)
        else:
            raise NotImplementedError("mvuld_b200 my_models.Multi_DefectModel: projection_layer 'attention' is not built "
                                      "(the reference ships 'gru')")
        hv_b = e((B, 512), bf)
        _lib.call("mvuld_f32_to_bf16", hv, hv_b, B * 512)
        hproj = e((B, 512), f32)
        _lib.gemm(hv_b, p["h"][0], bias=p["h"][1], act=_lib.ACT_ELU, out_f32=hproj)
        K = self.final_fc.weight.shape[1]
        if p["fusion"] in ("attention", "dot"):
            feats = e((B, K), f32)
            _lib.call("mvuld_gate_fusion", ximg, hproj, feats, B, 512, K, 0, 0 if p["fusion"] == "attention" else 1)
            feats[:, 512:] = xtxt
        else:
            feats = torch.cat([ximg, hproj, xtxt], 1)
        logits = e((B, self.num_classes), f32)
        _lib.call("mvuld_linear_small", feats.contiguous(), p["final"][0], p["final"][1], logits, None, B, self.num_classes, K)
        g.check_status()
        return logits
