"""Live fusion model ``Multi_DefectModel_new_GCN`` and the GGNN baseline branch, B200-native.

Interface mirror of /root/reference/mvuld/models/GraphModel.py:81-211 (constructor ``(config, pretrained=True,
attention=True)``, ``forward(g, img_embedding, func_text_embedding) -> logits [B, num_classes]``, same state-dict
keys incl. the DGL ``GATConv`` ones ``gat.fc.weight / gat.attn_l / gat.attn_r / gat.bias``), of
/root/reference/mvuld/models/Rs_GCN.py:7-73 (``Rs_GCN`` parameter names) and of
/root/reference/baselines/models/reveal/ggnn/model.py:8-31 (``GGNNSum``; DGL ``GatedGraphConv`` keys
``ggnn.linears.{t}.weight``, ``ggnn.gru.weight_ih`` ...).

``g`` is a :class:`mvuld_b200.graph.Graph` (or a real DGLGraph through ``graph.from_dgl``).  Eval-mode semantics:
dropout off, every BatchNorm uses its running statistics and is folded into the adjacent linear layer / a per-slot
affine when the weights are packed.  The dead ``h_func`` branch (GraphModel.py:172,177) is not evaluated.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .graph import Graph


# ------------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------------
class GATConv(nn.Module):
    """DGL ``GATConv`` parameters (fc without bias, attn_l / attn_r [1, H, F], bias [H*F]); DGL's reset_parameters."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., negative_slope=0.2):
        super().__init__()
        self._in, self._out, self._heads, self.negative_slope = in_feats, out_feats, num_heads, negative_slope
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.bias = nn.Parameter(torch.zeros(num_heads * out_feats))
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)


class Rs_GCN(nn.Module):
    """Rs_GCN.py:9-50 parameters: g / theta / phi = Conv1d(k=1), W = Sequential(Conv1d(k=1), BatchNorm1d) zero-init."""

    def __init__(self, in_channels, inter_channels, bn_layer=True):
        super().__init__()
        if not bn_layer:
            raise NotImplementedError("mvuld_b200 Rs_GCN: bn_layer=True only (as the fusion model builds it)")
        self.in_channels, self.inter_channels = in_channels, inter_channels or max(in_channels // 2, 1)
        self.g = nn.Conv1d(in_channels, self.inter_channels, 1)
        self.W = nn.Sequential(nn.Conv1d(self.inter_channels, in_channels, 1), nn.BatchNorm1d(in_channels))
        nn.init.constant_(self.W[1].weight, 0)
        nn.init.constant_(self.W[1].bias, 0)
        self.theta = nn.Conv1d(in_channels, self.inter_channels, 1)
        self.phi = nn.Conv1d(in_channels, self.inter_channels, 1)


def _bn_affine(bn: nn.BatchNorm1d):
    """eval-mode BatchNorm as y = x * scale + shift (fp32, on the parameters' device)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _fold_bn_into_linear(bn: nn.BatchNorm1d, lin: nn.Linear):
    """lin(bn(x)) == x @ W'.T + b' with W' = W * scale, b' = b + W @ shift."""
    scale, shift = _bn_affine(bn)
    w = lin.weight.detach().float()
    return w * scale[None, :], lin.bias.detach().float() + w @ shift


@torch.no_grad()
def plan_rs_gcn_chain(blocks, dev):
    """Packed weights of a chain of Rs_GCN blocks: (theta | phi | g) as one projection, the BatchNorm after the 1x1
    convolution W folded on its output side, both as bf16x3 split weights (W_hi | W_hi | W_lo) -- fp32-class 1x1
    convolutions, see mvuld_rs_gcn_affinity_f32."""
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    plan = []
    for m in blocks:
        wcat = torch.cat([m.theta.weight[:, :, 0], m.phi.weight[:, :, 0], m.g.weight[:, :, 0]], 0)
        bcat = torch.cat([m.theta.bias, m.phi.bias, m.g.bias], 0)
        scale, shift = _bn_affine(m.W[1])
        ww = m.W[0].weight.detach().float()[:, :, 0] * scale[:, None]
        wb = m.W[0].bias.detach().float() * scale + shift
        split = []
        for w32 in (f32(wcat), f32(ww)):
            w3 = torch.empty(w32.shape[0], 3 * w32.shape[1], device=dev, dtype=torch.bfloat16)
            _lib.call("mvuld_split3_bf16", w32, w32.shape[1], w3, w32.shape[0], w32.shape[1], 1)
            split.append(w3)
        plan.append(dict(wcat3=split[0], bcat=f32(bcat), ww3=split[1], wb=f32(wb)))
    return plan


@torch.no_grad()
def run_rs_gcn_chain(plan, z32: torch.Tensor, B: int, n: int) -> torch.Tensor:
    """Rs_GCN.py:52-73 applied block after block, in place on the token-major fp32 [B*n, 512] tensor.  The residual
    stream stays fp32 and every product uses bf16x3 split operands: no softmax bounds the affinity and a BatchNorm
    follows, so plain bf16 operands cost 1-2 % per block (measured in train mode)."""
    dev = z32.device
    z3 = torch.empty((B * n, 1536), device=dev, dtype=torch.bfloat16)
    y3 = torch.empty((B * n, 1536), device=dev, dtype=torch.bfloat16)
    tpg = torch.empty((B * n, 1536), device=dev, dtype=torch.float32)
    for gc in plan:
        _lib.call("mvuld_split3_bf16", z32, 512, z3, B * n, 512, 0)
        _lib.gemm(z3, gc["wcat3"], bias=gc["bcat"], out_f32=tpg)
        _lib.call("mvuld_rs_gcn_affinity_f32", tpg, y3, None, B, n, 512)
        _lib.gemm(y3, gc["ww3"], bias=gc["wb"], res=z32, out_f32=z32)
    return z32


class Multi_DefectModel_new_GCN(nn.Module):
    HEAD_MODE = 0          # mvuld_fusion_head_mode: 0 cat(img, graph, txt); the RQ2 subclasses in fusion_variants.py use 1 / 2
    HEAD_FEATS = 3         # width of final_fc / final_fc_bn in units of hfeat

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
        for k in range(1, 9):
            setattr(self, f"Rs_GCN_{k}", Rs_GCN(in_channels=512, inter_channels=512))
        self.bn_text = nn.BatchNorm1d(embfeat)
        self.ln_text = nn.LayerNorm(embfeat)
        self.fc_text = nn.Linear(embfeat, hfeat)
        self.max_node = 100
        self.bn_gat = nn.BatchNorm1d(self.max_node)
        self.fc_gat = nn.Linear(512, 480)
        self.bn_bbox = nn.BatchNorm1d(self.max_node)
        self.fc_bbox = nn.Linear(4, 32)
        self.swinbn = nn.BatchNorm1d(self.num_features)
        self.swinfc = nn.Linear(self.num_features, hfeat)
        self.hbn = nn.BatchNorm1d(hfeat)
        self.hln = nn.LayerNorm(hfeat)
        self.hfc = nn.Linear(hfeat, hfeat)
        self.final_fc = nn.Linear(hfeat * self.HEAD_FEATS, self.num_classes)
        self.final_fc_bn = nn.BatchNorm1d(hfeat * self.HEAD_FEATS)
        self._plan = None
        # Input validity (DGL raises on 0-in-degree nodes; edge endpoints must lie in [0, N)) is detected by the kernels
        # and read back here.  With defer_checks = True the read-back (a host synchronisation per call) is postponed to
        # raise_if_invalid(), so a pipelined caller (mvuld_b200.prefetch) can keep several steps in flight.
        self.defer_checks = False
        self._pending = []

    def invalidate(self):
        self._plan = None

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        self._engine = None                 # the flat training buffers belong to the old device / dtype
        return super()._apply(fn, *a, **k)

    def raise_if_invalid(self):
        """Raise for any input problem recorded since the last call (one host synchronisation)."""
        pending, self._pending = self._pending, []
        for zero_deg, g in pending:
            if int(zero_deg.item()) != 0:
                raise RuntimeError("There are 0-in-degree nodes in the graph (GATConv allow_zero_in_degree=False); "
                                   "add self-loops with mvuld_b200.graph.add_self_loop")
            g.check_status()

    @torch.no_grad()
    def prepare(self):
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        p = dict(dev=dev)
        w, b = _fold_bn_into_linear(self.swinbn, self.swinfc)
        p["img"] = (b16(w), f32(b))
        w, b = _fold_bn_into_linear(self.bn_text, self.fc_text)
        p["txt"] = (b16(w), f32(b))
        for name in ("gat", "gat2"):
            m = getattr(self, name)
            p[name] = dict(w=b16(m.fc.weight), al=f32(m.attn_l.view(-1)), ar=f32(m.attn_r.view(-1)), bias=f32(m.bias),
                           H=m._heads, F=m._out, slope=float(m.negative_slope))
        p["fc"] = (b16(self.fc.weight), f32(self.fc.bias))
        p["hidden"] = [(b16(l.weight), f32(l.bias)) for l in self.hidden]
        s, t = _bn_affine(self.bn_gat)
        p["bn_gat"] = (f32(s), f32(t))
        p["fc_gat"] = (b16(self.fc_gat.weight), f32(self.fc_gat.bias))
        s, t = _bn_affine(self.bn_bbox)
        p["bn_bbox"] = (f32(s), f32(t))
        p["fc_bbox"] = (f32(self.fc_bbox.weight), f32(self.fc_bbox.bias))
        p["gcn"] = plan_rs_gcn_chain([getattr(self, f"Rs_GCN_{k}") for k in range(1, 9)], dev)
        scale, shift = _bn_affine(self.final_fc_bn)
        wf = self.final_fc.weight.detach().float()
        p["final"] = (f32(wf * scale[None, :]), f32(self.final_fc.bias.detach().float() + wf @ shift))
        self._plan = p
        return self

    @torch.no_grad()
    def graph_features(self, g: Graph) -> torch.Tensor:
        """The graph half of GraphModel.py:150-211 -- GATConv x2, node MLP, unbatch / pad to ``max_node`` slots, slot
        BatchNorms, fc_gat | fc_bbox, the eight Rs_GCN blocks -> fp32 [B * max_node, 512].  It needs nothing from the
        image / text branches, so a caller may run it on another CUDA stream while they compute (``MVulD.forward``)."""
        if self.training:
            raise RuntimeError("mvuld_b200 fusion model implements the eval-mode forward: call model.eval()")
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        if g.device.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        B, N = g.batch_size, g.num_nodes()
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32

        # graph branch (GraphModel.py:163-177)
        indptr, idx_src, _ = g.in_csr()
        offsets = g.node_offsets()
        h_in = g.ndata["_UNIX_NODE_EMB"]
        pos = g.ndata["pos_emb"].float().contiguous()
        hb = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), hb, N * h_in.shape[1])
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
        a = e((N, 512), bf)
        _lib.gemm(hb, p["fc"][0], bias=p["fc"][1], act=_lib.ACT_ELU, out_bf16=a)
        a2 = e((N, 512), bf)
        for (w, b) in p["hidden"]:
            _lib.gemm(a, w, bias=b, act=_lib.ACT_ELU, out_bf16=a2)
            a, a2 = a2, a
        g.ndata['HGATOUTPUT'] = a                      # side effects of GraphModel.py:180-181 (bf16 here)
        g.ndata['HFGATOUTPUT'] = pos

        # unbatch -> pad/truncate to max_node -> slot BN -> fc_gat / fc_bbox -> concat (GraphModel.py:182-189)
        n = self.max_node
        hp = e((B * n, 512), bf)
        _lib.call("mvuld_unbatch_pad_bn", a, offsets, p["bn_gat"][0], p["bn_gat"][1], hp, None, B, n, 512)
        z32, zb = e((B * n, 512), f32), e((B * n, 512), bf)
        _lib.gemm(hp, p["fc_gat"][0], bias=p["fc_gat"][1], act=_lib.ACT_ELU, out_bf16=zb, out_f32=z32)
        _lib.call("mvuld_pos_branch", pos, offsets, p["bn_bbox"][0], p["bn_bbox"][1], p["fc_bbox"][0],
                  p["fc_bbox"][1], z32, zb, B, n, 32, 512, 480)

        # 8 x Rs_GCN on the token-major [B*n, 512] tensor (GraphModel.py:190-198)
        run_rs_gcn_chain(p["gcn"], z32, B, n)
        self._pending.append((zero_deg, g))
        return z32

    @torch.no_grad()
    def head(self, z32: torch.Tensor, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
        """The rest of GraphModel.py:150-211: image / text projections (:153-159), l2norm over the slot axis + mean +
        concat + BatchNorm + final_fc (:200-209) -> logits [B, num_classes]."""
        if not img_embedding.is_cuda:
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        B, n = img_embedding.shape[0], self.max_node
        if z32.shape[0] != B * n:
            raise ValueError(f"graph batch size {z32.shape[0] // n} != embedding batch size {B}")
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32
        img_b, txt_b = e((B, 1024), bf), e((B, 768), bf)
        _lib.call("mvuld_f32_to_bf16", img_embedding.float().contiguous(), img_b, B * 1024)
        _lib.call("mvuld_f32_to_bf16", func_text_embedding.float().contiguous(), txt_b, B * 768)
        ximg, xtxt = e((B, 512), f32), e((B, 512), f32)
        _lib.gemm(img_b, p["img"][0], bias=p["img"][1], act=_lib.ACT_ELU, out_f32=ximg)
        _lib.gemm(txt_b, p["txt"][0], bias=p["txt"][1], act=_lib.ACT_ELU, out_f32=xtxt)
        logits = e((B, self.num_classes), f32)
        _lib.call("mvuld_fusion_head_mode", z32, ximg, xtxt, p["final"][0], p["final"][1], logits, None, B, n, 512,
                  self.num_classes, self.HEAD_MODE)
        if not self.defer_checks:
            self.raise_if_invalid()
        elif len(self._pending) > 64:
            self.raise_if_invalid()                      # bound the backlog of a caller that never asks
        return logits

    def train(self, mode: bool = True):
        """Switching between train and eval drops the packed eval-mode plan: an optimiser may have changed the
        parameters, and train-mode forwards move the BatchNorm running statistics the plan folds in."""
        self._plan = None
        return super().train(mode)

    def train_engine(self):
        """The flat-buffer training engine behind the train-mode ``forward`` (``mvuld_b200.train.FusionTrainer`` used
        for its forward / backward launch sequences only; created on first use, re-created when the model moves)."""
        eng = getattr(self, "_engine", None)
        if eng is None or eng.flat_p.device != self.fc.weight.device:
            from .train import FusionTrainer
            eng = FusionTrainer(self, world_size=1)
            self._engine = eng
        return eng

    def forward(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
        """GraphModel.py:150-211.  ``model.eval()``: the folded inference path, no autograd graph.  ``model.train()``:
        the train-mode forward (dropout, BatchNorm on batch statistics) as ONE autograd node whose backward is the
        hand-written backward launch sequence, so the reference loop body (main_bigvul.py:328-342:
        ``outputs = model(...)``, ``loss_scaler(loss, optimizer, ...)`` with ``ACCUMULATION_STEPS``, any
        ``torch.optim`` optimiser, DDP) runs against this module unchanged."""
        if not img_embedding.is_cuda:
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if g.batch_size != img_embedding.shape[0]:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {img_embedding.shape[0]}")
        if self.training:
            if type(self).HEAD_MODE != 0 or type(self).HEAD_FEATS != 3:
                raise RuntimeError("mvuld_b200: train-mode forward is built for Multi_DefectModel_new_GCN only")
            from .autograd import fusion_train_forward
            return fusion_train_forward(self, g, img_embedding, func_text_embedding)
        with torch.no_grad():
            return self.head(self.graph_features(g), img_embedding, func_text_embedding)


class Multi_DefectModel(nn.Module):
    """The RQ3 "GAT" ablation (mvuld/models/GraphModel.py:214-304, commented alternative at main_bigvul.py:141): the
    same GATConv x2 + node MLP as the live model, then ``dgl.mean_nodes`` over each graph, ``ELU(hfc(hbn(.)))``, concat
    with the image / text projections, ``final_fc(final_fc_bn(.))`` -- no slot padding, no Rs_GCN.  Same constructor,
    ``forward`` signature and state-dict keys as the reference class; eval-mode semantics (BatchNorms folded)."""

    def __init__(self, config, pretrained=True, attention=True):
        super().__init__()
        self.num_features = 1024
        self.config = config
        self.num_classes = config.MODEL.NUM_CLASSES
        hfeat, embfeat, numheads = 512, 768, 4
        self.gat = GATConv(embfeat, hfeat, numheads, feat_drop=0.1)
        self.gat2 = GATConv(hfeat * numheads, hfeat, numheads, feat_drop=0.1)
        self.fc = nn.Linear(hfeat * numheads, hfeat)
        self.fconly = nn.Linear(embfeat, hfeat)
        self.hidden = nn.ModuleList([nn.Linear(hfeat, hfeat) for _ in range(8)])
        self.bn_text = nn.BatchNorm1d(embfeat)
        self.fc_text = nn.Linear(embfeat, hfeat)
        self.swinbn = nn.BatchNorm1d(self.num_features)
        self.swinfc = nn.Linear(self.num_features, hfeat)
        self.hbn = nn.BatchNorm1d(hfeat)
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
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 fusion model runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        p = dict(dev=dev)
        for key, bn, lin in (("img", self.swinbn, self.swinfc), ("txt", self.bn_text, self.fc_text),
                             ("h", self.hbn, self.hfc)):
            w, b = _fold_bn_into_linear(bn, lin)
            p[key] = (b16(w), f32(b))
        for name in ("gat", "gat2"):
            m = getattr(self, name)
            p[name] = dict(w=b16(m.fc.weight), al=f32(m.attn_l.view(-1)), ar=f32(m.attn_r.view(-1)), bias=f32(m.bias),
                           H=m._heads, F=m._out, slope=float(m.negative_slope))
        p["fc"] = (b16(self.fc.weight), f32(self.fc.bias))
        p["hidden"] = [(b16(l.weight), f32(l.bias)) for l in self.hidden]
        scale, shift = _bn_affine(self.final_fc_bn)
        wf = self.final_fc.weight.detach().float()
        p["final"] = (f32(wf * scale[None, :]), f32(self.final_fc.bias.detach().float() + wf @ shift))
        self._plan = p
        return self

    @torch.no_grad()
    def forward(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
        """GraphModel.py:263-304."""
        if self.training:
            raise RuntimeError("mvuld_b200 fusion model implements the eval-mode forward: call model.eval()")
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if not img_embedding.is_cuda:
            raise RuntimeError("mvuld_b200 fusion model takes CUDA tensors (no CPU fallback)")
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        B, N = img_embedding.shape[0], g.num_nodes()
        if g.batch_size != B:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {B}")
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        bf, f32 = torch.bfloat16, torch.float32
        feats = e((B, 1536), f32)                                   # cat(x, h_feature, text) (GraphModel.py:300)
        img_b, txt_b = e((B, 1024), bf), e((B, 768), bf)
        _lib.call("mvuld_f32_to_bf16", img_embedding.float().contiguous(), img_b, B * 1024)
        _lib.call("mvuld_f32_to_bf16", func_text_embedding.float().contiguous(), txt_b, B * 768)
        _lib.gemm(img_b, p["img"][0], bias=p["img"][1], act=_lib.ACT_ELU, out_f32=feats[:, 0:512])
        _lib.gemm(txt_b, p["txt"][0], bias=p["txt"][1], act=_lib.ACT_ELU, out_f32=feats[:, 1024:1536])
        indptr, idx_src, _ = g.in_csr()
        h_in = g.ndata["_UNIX_NODE_EMB"]
        hb = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), hb, N * h_in.shape[1])
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
        a32 = e((N, 512), f32)
        for i, (w, b) in enumerate(p["hidden"]):
            last = i == len(p["hidden"]) - 1
            _lib.gemm(a, w, bias=b, act=_lib.ACT_ELU, out_bf16=a2, out_f32=a32 if last else None)
            a, a2 = a2, a
        # dgl.mean_nodes (GraphModel.py:296-298): per-graph mean over its node rows
        bnn = g.batch_num_nodes().to(torch.int32)
        start = (torch.cumsum(bnn, 0, dtype=torch.int32) - bnn).to(dev)
        hmean = e((B, 512), f32)
        _lib.call("mvuld_seq_segment_mean", a32, start, bnn.to(dev), None, hmean, B, 512)
        hmean_b = e((B, 512), bf)
        _lib.call("mvuld_f32_to_bf16", hmean, hmean_b, B * 512)
        _lib.gemm(hmean_b, p["h"][0], bias=p["h"][1], act=_lib.ACT_ELU, out_f32=feats[:, 512:1024])
        logits = e((B, self.num_classes), f32)
        _lib.call("mvuld_linear_small", feats, p["final"][0], p["final"][1], logits, None, B, self.num_classes, 1536)
        if int(zero_deg.item()) != 0:
            raise RuntimeError("There are 0-in-degree nodes in the graph (GATConv allow_zero_in_degree=False); "
                               "add self-loops with mvuld_b200.graph.add_self_loop")
        g.check_status()
        return logits


class GatedGraphConv(nn.Module):
    """DGL ``GatedGraphConv`` parameters: linears[t] = Linear(out, out), gru = GRUCell(out, out)."""

    def __init__(self, in_feats, out_feats, n_steps, n_etypes, bias=True):
        super().__init__()
        self._in_feats, self._out_feats, self._n_steps, self._n_etypes = in_feats, out_feats, n_steps, n_etypes
        self.linears = nn.ModuleList([nn.Linear(out_feats, out_feats) for _ in range(n_etypes)])
        self.gru = nn.GRUCell(out_feats, out_feats, bias=bias)
        gain = nn.init.calculate_gain('relu')
        for lin in self.linears:
            nn.init.xavier_normal_(lin.weight, gain=gain)
            nn.init.zeros_(lin.bias)


class GGNNSum(nn.Module):
    """baselines/models/reveal/ggnn/model.py:8-31."""

    def __init__(self, input_dim, output_dim, max_edge_types=3, num_steps=8):
        super().__init__()
        if input_dim > output_dim:
            raise ValueError("GatedGraphConv requires in_feats <= out_feats")
        if output_dim % 8 != 0 or output_dim > 256:
            raise NotImplementedError("mvuld_b200 GGNN: out_feats must be a multiple of 8 and <= 256")
        self.inp_dim, self.out_dim = input_dim, output_dim
        self.max_edge_types, self.num_timesteps = max_edge_types, num_steps
        self.ggnn = GatedGraphConv(input_dim, output_dim, num_steps, max_edge_types)
        self.classifier = nn.Linear(output_dim, 1)
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
        dev = self.classifier.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 GGNN runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        gg = self.ggnn
        D = self.out_dim
        # GRUCell as one GEMM over [a | h] with the four pre-activations of feature j in adjacent output columns
        # (mvuld_gemm_gru): rows 4j .. 4j+3 = (W_ir | W_hr), (W_iz | W_hz), (W_in | 0), (0 | W_hn)
        wih, whh = gg.gru.weight_ih.detach().float(), gg.gru.weight_hh.detach().float()     # [3D, D], gate order r, z, n
        wg = torch.zeros(D, 4, 2 * D, device=wih.device)
        wg[:, 0, :D], wg[:, 0, D:] = wih[:D], whh[:D]
        wg[:, 1, :D], wg[:, 1, D:] = wih[D:2 * D], whh[D:2 * D]
        wg[:, 2, :D] = wih[2 * D:]
        wg[:, 3, D:] = whh[2 * D:]
        if gg.gru.bias_ih is not None:
            bih, bhh = gg.gru.bias_ih.detach().float(), gg.gru.bias_hh.detach().float()
        else:
            bih = bhh = torch.zeros(3 * D, device=wih.device)
        b4 = torch.stack([bih[:D] + bhh[:D], bih[D:2 * D] + bhh[D:2 * D], bih[2 * D:], bhh[2 * D:]], 1)
        self._plan = dict(
            dev=dev,
            wmsg=b16(torch.cat([l.weight for l in gg.linears], 0)),     # [T*D, D]: row t*D + o = linears[t].weight[o]
            bmsg=f32(torch.cat([l.bias for l in gg.linears], 0)),
            wg=b16(wg.reshape(4 * D, 2 * D)), b4=f32(b4.reshape(-1)),
            wc=f32(self.classifier.weight), bc=f32(self.classifier.bias))
        return self

    @torch.no_grad()
    def node_states(self, g: Graph) -> torch.Tensor:
        """GatedGraphConv forward -> fp32 [N, out] (``g.ndata['GGNNOUTPUT']`` in the reference)."""
        if self.training:
            raise RuntimeError("mvuld_b200 GGNN implements the eval-mode forward: call model.eval()")
        if self._plan is None:
            self.prepare()
        p = self._plan
        dev = p["dev"]
        feats = g.ndata['_WORD2VEC'].float().contiguous()
        et = g.edata["_ETYPE"].to(torch.int64).contiguous()
        N, D, T = g.num_nodes(), self.out_dim, self.max_edge_types
        indptr, idx_src, eids = g.in_csr()
        status = torch.zeros(1, device=dev, dtype=torch.int32)
        et_sorted = torch.empty(et.numel(), device=dev, dtype=torch.uint8)
        _lib.call("mvuld_gather_etype", et, eids, et.numel(), T, et_sorted, status)
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        h32 = e((N, D), torch.float32)
        # two [a | h] operand buffers (bf16 [N, 2D]): step t reads X[t % 2] and writes the new state into the h half of
        # the other one (tiles of the same rows are still loading the old state while the epilogue runs)
        X = [e((N, 2 * D), torch.bfloat16), e((N, 2 * D), torch.bfloat16)]
        _lib.call("mvuld_ggnn_init", feats, h32, _lib._Raw(X[0][:, D:]), 2 * D, N, feats.shape[1], D)
        msgs = e((N, T * D), torch.bfloat16)
        for step in range(self.num_timesteps):
            cur, nxt = X[step % 2], X[(step + 1) % 2]
            _lib.gemm(cur[:, D:], p["wmsg"], bias=p["bmsg"], out_bf16=msgs)
            _lib.call("mvuld_ggnn_gather_sum", msgs, indptr, idx_src, et_sorted, cur, 2 * D, N, T, D)
            _lib.call("mvuld_gemm_gru", cur, 2 * D, p["wg"], 2 * D, N, D, 2 * D, p["b4"], h32,
                      _lib._Raw(nxt[:, D:]), 2 * D)
        if int(status.item()) != 0:
            raise AssertionError("edge type indices out of range [0, n_etypes)")
        g.check_status()
        return h32

    @torch.no_grad()
    def forward(self, g: Graph, dataset=None, cuda=False):
        """reveal/ggnn/model.py:20-31 -> (sigmoid(logit) [B], logit [B, 1])."""
        h = self.node_states(g)
        g.ndata['GGNNOUTPUT'] = h
        p = self._plan
        B = g.batch_size
        s = torch.empty(B, self.out_dim, device=h.device, dtype=torch.float32)
        _lib.call("mvuld_segment_sum", h, g.node_offsets(), s, B, self.out_dim)
        logit = torch.empty(B, 1, device=h.device, dtype=torch.float32)
        prob = torch.empty(B, 1, device=h.device, dtype=torch.float32)
        _lib.call("mvuld_linear_small", s, p["wc"], p["bc"], logit, prob, B, 1, self.out_dim)
        self._last_sum = s
        return prob.squeeze(-1), logit

    def save_ggnn_output(self, g, dataset=None, cuda=False):
        """reveal/ggnn/model.py:33-44 -> (prob, h_i_sum)."""
        prob, _ = self.forward(g)
        return prob, self._last_sum
