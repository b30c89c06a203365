"""Training step of the fusion model, B200-native (BASELINE.json configs[4], reference-faithful variant).

What the reference trains (/root/reference/mvuld/main_bigvul.py:294-342): only ``Multi_DefectModel_new_GCN`` -- the
SwinV2 / UniXcoder vectors come from frozen encoders (cached offline, data_list.py:179-211,292-313).  One step is

    model.train(); outputs = model(g, img_embedding, func_text_embedding)     # GraphModel.py:150-211, train mode
    loss = CrossEntropyLoss()(outputs, targets)                                # main_bigvul.py:298,332
    loss.backward(); clip_grad_norm_(parameters, 5.0); AdamW.step()            # utils_multi.py:225-240, optimizer.py:11-33
    lr_scheduler.step_update(...)                                              # lr_scheduler.py:12-30 (cosine + warm-up)

under DDP with per-rank BatchNorm statistics and ``find_unused_parameters=True`` (main_bigvul.py:162-164: the dead
``h_func`` branch leaves ``fconly / ln_text / hbn / hln / hfc`` without gradients; AdamW skips them).

Here the step is a fixed sequence of C-ABI launches (include/mvuld_b200.h): every dense forward / backward product runs
on the tcgen05 GEMM (dX = dY W on a transposed weight copy, dW = dY^T X on transposed activations, fp32 result written
straight into the flat gradient buffer), the sparse GATConv backward, BatchNorm (batch statistics) forward / backward,
Rs_GCN affinity backward, the l2norm / mean / cross-entropy head and the clipped AdamW update are the kernels of
csrc/train.cu.  Parameters live in ONE flat fp32 buffer ordered by backward completion, so data-parallel training is a
handful of bucketed NCCL all-reduces over contiguous slices, launched as soon as the backward pass has produced a
bucket (overlapping the rest of the backward pass), then one gradient-norm reduction and one AdamW launch.
No CPU path: tensors must be CUDA tensors and ``libmvuld_b200.so`` must be present.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .graph import Graph
from .graph_model import Multi_DefectModel_new_GCN

_ALIGN = 64          # elements: every parameter starts on a 256-byte boundary of the flat buffers (TMA needs 16 B)


def trainable_parameter_order(model: Multi_DefectModel_new_GCN) -> List[str]:
    """Names of the parameters that receive a gradient, in the order the backward pass finishes them."""
    names = ["final_fc.weight", "final_fc.bias", "final_fc_bn.weight", "final_fc_bn.bias",
             "swinfc.weight", "swinfc.bias", "swinbn.weight", "swinbn.bias",
             "fc_text.weight", "fc_text.bias", "bn_text.weight", "bn_text.bias"]
    for k in range(8, 0, -1):
        p = f"Rs_GCN_{k}."
        names += [p + "W.1.weight", p + "W.1.bias", p + "W.0.weight", p + "W.0.bias",
                  p + "theta.weight", p + "phi.weight", p + "g.weight", p + "theta.bias", p + "phi.bias", p + "g.bias"]
    names += ["fc_gat.weight", "fc_gat.bias", "bn_gat.weight", "bn_gat.bias",
              "fc_bbox.weight", "fc_bbox.bias", "bn_bbox.weight", "bn_bbox.bias"]
    for i in range(7, -1, -1):
        names += [f"hidden.{i}.weight", f"hidden.{i}.bias"]
    names += ["fc.weight", "fc.bias", "gat2.bias", "gat2.attn_l", "gat2.attn_r", "gat2.fc.weight",
              "gat.bias", "gat.attn_l", "gat.attn_r", "gat.fc.weight"]
    have = dict(model.named_parameters())
    missing = [n for n in names if n not in have]
    if missing:
        raise KeyError(f"fusion model lacks parameters {missing}")
    return names


def plan_layout(shapes: Sequence[Tuple[str, int]], align: int = _ALIGN) -> Tuple[Dict[str, int], int]:
    """Offsets (elements) of each parameter in the flat buffers; every parameter is padded to ``align`` elements."""
    off, cur = {}, 0
    for name, numel in shapes:
        off[name] = cur
        cur += (numel + align - 1) // align * align
    return off, cur


def plan_buckets(ends: Sequence[int], bucket_elems: int) -> List[Tuple[int, int]]:
    """Split [0, ends[-1]) into contiguous buckets that close on parameter boundaries once they hold at least
    ``bucket_elems`` elements (the tail bucket may be smaller).  ``ends`` = exclusive end offset of each parameter in
    backward-completion order."""
    out, lo = [], 0
    for e in ends:
        if e - lo >= bucket_elems:
            out.append((lo, e))
            lo = e
    if ends and lo < ends[-1]:
        out.append((lo, ends[-1]))
    return out


def cosine_lr(step: int, total_steps: int, warmup_steps: int, base_lr: float, min_lr: float, warmup_lr: float) -> float:
    """timm ``CosineLRScheduler(t_initial=total, lr_min, warmup_lr_init, warmup_t, cycle_limit=1, t_in_epochs=False)``
    as lr_scheduler.py:19-30 builds it (timm default ``warmup_prefix=False``)."""
    if step < warmup_steps:
        return warmup_lr + step * (base_lr - warmup_lr) / max(warmup_steps, 1)
    if step >= total_steps:
        return min_lr
    return min_lr + 0.5 * (base_lr - min_lr) * (1.0 + math.cos(math.pi * step / total_steps))


def allreduce_buckets(flat: torch.Tensor, buckets: Sequence[Tuple[int, int]], group=None, async_op: bool = False):
    """SUM all-reduce of each bucket slice of ``flat`` (DDP's bucketed gradient reduction, main_bigvul.py:162-164)."""
    import torch.distributed as dist
    works = [dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=async_op) for lo, hi in buckets]
    return works


class FusionTrainer:
    """Owns the flat parameter / gradient / AdamW state of a ``Multi_DefectModel_new_GCN`` and runs training steps."""

    def __init__(self, model: Multi_DefectModel_new_GCN, lr: float = 5e-5, weight_decay: float = 0.005,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, clip_grad: float = 5.0,
                 dropout: Optional[float] = None, seed: int = 12345, process_group=None, world_size: Optional[int] = None,
                 bucket_mb: float = 8.0, bn_momentum: float = 0.1):
        dev = model.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 FusionTrainer runs on CUDA only (no CPU fallback): move the model to the GPU")
        _lib.load()
        self.model, self.dev = model, dev
        self.lr, self.wd, self.betas, self.eps, self.clip = float(lr), float(weight_decay), betas, float(eps), float(clip_grad)
        self.p_drop = 0.2 if dropout is None else float(dropout)     # gatdrop = mlpdropout = hdropout = 0.2 (GraphModel.py:93-95)
        self.seed, self.step_count, self.momentum = int(seed), 0, float(bn_momentum)
        self.group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world = int(world_size)

        self.names = trainable_parameter_order(model)
        params = dict(model.named_parameters())
        self.offsets, self.total = plan_layout([(n, params[n].numel()) for n in self.names])
        self.shapes = {n: tuple(params[n].shape) for n in self.names}
        f32 = dict(device=dev, dtype=torch.float32)
        self.flat_p = torch.zeros(self.total, **f32)
        self.flat_g = torch.zeros(self.total, **f32)
        self.flat_m = torch.zeros(self.total, **f32)
        self.flat_v = torch.zeros(self.total, **f32)
        self.flat_w16 = torch.zeros(self.total, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            for n in self.names:                     # re-point the module parameters at the flat buffer
                view = self._view(self.flat_p, n)
                view.copy_(params[n].detach().float())
                params[n].data = view
        # AdamW parameter groups (optimizer.py:35-50): 1-D tensors and biases are not decayed
        seg_end, seg_wd = [], []
        for n in self.names:
            no_decay = len(self.shapes[n]) == 1 or n.endswith(".bias")
            seg_end.append(self.offsets[n] + (params[n].numel() + _ALIGN - 1) // _ALIGN * _ALIGN)
            seg_wd.append(0.0 if no_decay else self.wd)
        self.seg_end = torch.tensor(seg_end, dtype=torch.int64, device=dev)
        self.seg_wd = torch.tensor(seg_wd, dtype=torch.float32, device=dev)
        self.buckets = plan_buckets(seg_end, int(bucket_mb * (1 << 20) / 4))
        self.gnorm_sq = torch.zeros(1, **f32)
        self.gnorm_partials = torch.zeros(1184, **f32)
        self.loss_buf = torch.zeros(1, **f32)
        if self.world > 1:
            self.sync_replicas()               # DDP broadcasts rank 0's module state at construction
        else:
            self._refresh_shadows()
        self.last = {}
        self.debug_taps = None      # tests set this to a dict to receive intermediate activations

    # ----------------------------------------------------------------------------------------------------
    def _view(self, flat: torch.Tensor, name: str, shape=None) -> torch.Tensor:
        shp = self.shapes[name] if shape is None else shape
        n = 1
        for d in shp:
            n *= d
        o = self.offsets[name]
        return flat[o:o + n].view(shp)

    def _mat(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        """2-D [out, in] view of a weight (Conv1d k=1 weights drop the trailing axis, attn vectors flatten)."""
        shp = self.shapes[name]
        if len(shp) == 3 and shp[2] == 1:
            return self._view(flat, name, (shp[0], shp[1]))
        return self._view(flat, name)

    def _gcn_cat(self, flat: torch.Tensor, k: int, what: str) -> torch.Tensor:
        """theta | phi | g laid out back to back: one [1536, 512] weight / [1536] bias view (Rs_GCN.py:54-58)."""
        p = f"Rs_GCN_{k}."
        o = self.offsets[p + f"theta.{what}"]
        if what == "weight":
            return flat[o:o + 3 * 512 * 512].view(1536, 512)
        return flat[o:o + 1536]

    @torch.no_grad()
    def _refresh_shadows(self):
        """bf16 copies of every weight (GEMM operands) and their transposes (dX = dY W needs W^T as the K-major operand)."""
        _lib.call("mvuld_f32_to_bf16", self.flat_p, self.flat_w16, self.total)
        first = not hasattr(self, "_wt_table")
        if first:
            # sources are views of the flat bf16 / fp32 buffers (fixed addresses), destinations are allocated once: the
            # 29 transposes of a step are ONE launch over a device-side table
            self.wt, self.w3, self._w3_src, pairs = {}, {}, [], []
            lin = ["swinfc", "fc_text", "fc_gat", "fc", "gat2.fc"] + [f"hidden.{i}" for i in range(8)]
            srcs = [(name, self._mat(self.flat_w16, name + ".weight")) for name in lin]
            for k in range(1, 9):
                srcs.append((f"gcn{k}.cat", self._gcn_cat(self.flat_w16, k, "weight")))
                srcs.append((f"gcn{k}.W0", self._mat(self.flat_w16, f"Rs_GCN_{k}.W.0.weight")))
                # forward operands of the Rs_GCN 1x1 convolutions: bf16x3 split (W_hi | W_hi | W_lo) of the fp32 weights
                for key, w32 in ((f"gcn{k}.cat", self._gcn_cat(self.flat_p, k, "weight")),
                                 (f"gcn{k}.W0", self._mat(self.flat_p, f"Rs_GCN_{k}.W.0.weight"))):
                    self.w3[key] = torch.empty(w32.shape[0], 3 * w32.shape[1], device=self.dev, dtype=torch.bfloat16)
                    self._w3_src.append((w32, self.w3[key]))
            for name, w in srcs:
                R, C = w.shape
                self.wt[name] = torch.empty(C, (R + 7) // 8 * 8, device=self.dev, dtype=torch.bfloat16)
                pairs.append((w, self.wt[name]))
            self._wt_table = _lib.TransposeTable(pairs)
        self._wt_table.run()
        for w32, w3 in self._w3_src:
            _lib.call("mvuld_split3_bf16", w32, w32.shape[1], w3, w32.shape[0], w32.shape[1], 1)
        self.model.invalidate()

    def _transpose(self, x: torch.Tensor) -> torch.Tensor:
        """bf16 [R, C] (any row stride) -> [C, Rp] with Rp = R rounded up to 8 (zero filled): the K-major operand of a product
        that reduces over R."""
        R, C = x.shape
        assert x.stride(1) == 1
        Rp = (R + 7) // 8 * 8
        out = torch.empty(C, Rp, device=self.dev, dtype=torch.bfloat16)
        _lib.call("mvuld_transpose_bf16", _lib._Raw(x), x.stride(0), out, R, C, Rp)
        return out

    # ----------------------------------------------------------------------------------------------------
    def _grad_w(self, dy: torch.Tensor, x: torch.Tensor, gname: str, rows_out: Optional[int] = None,
                out: Optional[torch.Tensor] = None, G: Optional[torch.Tensor] = None):
        """dW [out, in] = dY^T X on the tensor cores, fp32 straight into the flat gradient buffer: ``mvuld_gemm_dw`` reads
        both operands row-major as they are (MN-major tcgen05 operands) and splits the M rows over CTAs."""
        dst = self._mat(self.flat_g if G is None else G, gname) if out is None else out
        _lib.gemm_dw(dy, x, dst, n_out=rows_out)

    def _seed(self, layer: int) -> int:
        return (self.seed * 1000003 + self.step_count * 257 + layer) & 0x7FFFFFFFFFFFFFFF

    # ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor,
                         targets: torch.Tensor, on_bucket=None):
        """Train-mode forward (GraphModel.py:150-211) + backward of the mean cross-entropy; fills ``flat_g`` (scaled
        by 1 / world so that a SUM all-reduce yields DDP's mean).  Returns (loss [1] fp32 device tensor, logits)."""
        if not targets.is_cuda:
            raise RuntimeError("mvuld_b200 FusionTrainer takes CUDA tensors (no CPU fallback)")
        logits, ctx = self.forward_train(g, img_embedding, func_text_embedding)
        B, C = logits.shape
        dlogits = torch.empty((B, C), device=self.dev, dtype=torch.float32)
        self.loss_buf.zero_()
        _lib.call("mvuld_ce_loss", logits, targets.to(torch.int64).contiguous(), self.loss_buf, dlogits, B, C,
                  1.0 / (B * self.world))
        self.flat_g.zero_()
        self.backward_train(ctx, dlogits, self.flat_g, on_bucket)
        return self.loss_buf, logits

    @torch.no_grad()
    def forward_train(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor,
                      seed_index: Optional[int] = None):
        """Train-mode forward of GraphModel.py:150-211 (dropout on, BatchNorm on batch statistics with the running
        statistics updated).  Returns (logits [B, num_classes] fp32, ctx): ``ctx`` holds what ``backward_train`` needs.
        ``seed_index`` selects the dropout masks of this pass (default: the optimiser step count)."""
        if not isinstance(g, Graph):
            from .graph import from_dgl
            g = from_dgl(g)
        if not (img_embedding.is_cuda and func_text_embedding.is_cuda):
            raise RuntimeError("mvuld_b200 FusionTrainer takes CUDA tensors (no CPU fallback)")
        m, dev = self.model, self.dev
        B, N, n = img_embedding.shape[0], g.num_nodes(), m.max_node
        if g.batch_size != B:
            raise ValueError(f"graph batch size {g.batch_size} != embedding batch size {B}")
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training (BatchNorm1d)")
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        P, W16 = self.flat_p, self.flat_w16
        pv = lambda name: self._view(P, name)
        w16 = lambda name: self._mat(W16, name)
        p, mom = self.p_drop, self.momentum
        R = B * n
        seed_base = self.seed * 1000003 + (self.step_count if seed_index is None else int(seed_index)) * 257
        _seed = lambda layer: (seed_base + layer) & 0x7FFFFFFFFFFFFFFF

        def bn_cols(prefix, x, R_, C_, res=None, ldr=0, y32=None, ldy=None, yb=None):
            bn = getattr(m, prefix) if "." not in prefix else m.get_submodule(prefix)
            mean, rstd = e((C_,), f32), e((C_,), f32)
            _lib.call("mvuld_bn_cols_fwd", x, pv(prefix + ".weight"), pv(prefix + ".bias"), float(bn.eps),
                      _lib._Raw(res) if res is not None else None, ldr,
                      _lib._Raw(y32) if y32 is not None else None, (ldy if ldy is not None else C_), yb, mean, rstd,
                      bn.running_mean, bn.running_var, mom, R_, C_)
            bn.num_batches_tracked += 1
            return mean, rstd

        # ---------------- forward ----------------
        feats = e((B, 1536), f32)                                        # cat(x, h_feature, text) (GraphModel.py:207)
        img32 = img_embedding.float().contiguous()
        txt32 = func_text_embedding.float().contiguous()
        img_n, txt_n = e((B, 1024), bf), e((B, 768), bf)
        img_stat = bn_cols("swinbn", img32, B, 1024, yb=img_n)
        txt_stat = bn_cols("bn_text", txt32, B, 768, yb=txt_n)
        _lib.gemm(img_n, w16("swinfc.weight"), bias=pv("swinfc.bias"), act=_lib.ACT_ELU, out_f32=feats[:, 0:512])
        _lib.gemm(txt_n, w16("fc_text.weight"), bias=pv("fc_text.bias"), act=_lib.ACT_ELU, out_f32=feats[:, 1024:1536])

        indptr, idx_src, _ = g.in_csr()
        out_indptr, out_dst, pos_in = g.out_csr()
        E = idx_src.numel()
        offsets = g.node_offsets()
        h_in = g.ndata["_UNIX_NODE_EMB"]
        pos = g.ndata["pos_emb"].float().contiguous()
        x0 = e((N, h_in.shape[1]), bf)
        _lib.call("mvuld_f32_to_bf16", h_in.float().contiguous(), x0, N * h_in.shape[1])
        zero_deg = torch.zeros(1, device=dev, dtype=torch.int32)
        gat_saved = []
        hcur = x0
        for li, name in enumerate(("gat", "gat2")):
            mod = getattr(m, name)
            H, F = mod._heads, mod._out
            if p > 0:                                                     # GATConv feat_drop on the input features
                xd = e(hcur.shape, bf)
                _lib.call("mvuld_dropout_bf16", hcur, xd, hcur.numel(), _seed(li), p)
            else:
                xd = hcur
            z = e((N, H * F), bf)
            _lib.gemm(xd, w16(name + ".fc.weight"), out_bf16=z)
            el, er = e((N, H), f32), e((N, H), f32)
            _lib.call("mvuld_gat_scores", z, pv(name + ".attn_l").view(-1), pv(name + ".attn_r").view(-1), el, er, N, H, F)
            hout = e((N, H * F), bf)
            _lib.call("mvuld_gat_aggregate", z, el, er, indptr, idx_src, pv(name + ".bias"), hout, N, H, F,
                      float(mod.negative_slope), zero_deg)
            gat_saved.append((xd, z, el, er, H, F, float(mod.negative_slope)))
            hcur = hout
            if self.debug_taps is not None:
                self.debug_taps["gat1" if li == 0 else "gat2"] = hout.clone()
                self.debug_taps["z1" if li == 0 else "z2"] = z.clone()
        acts = [hcur]                                                     # inputs of fc, hidden.0 ... hidden.7
        for li, name in enumerate(["fc"] + [f"hidden.{i}" for i in range(8)]):
            a = e((N, 512), bf)
            _lib.gemm(acts[-1], w16(name + ".weight"), bias=pv(name + ".bias"), act=_lib.ACT_ELU, out_bf16=a)
            if p > 0:                                                     # mlpdropout / hdropout (GraphModel.py:171,176)
                _lib.call("mvuld_dropout_bf16", a, a, a.numel(), _seed(8 + li), p)
            acts.append(a)
            if self.debug_taps is not None and li == 0:
                self.debug_taps["fc"] = a.clone()
        g.ndata['HGATOUTPUT'] = acts[-1]
        g.ndata['HFGATOUTPUT'] = pos

        ones, zeros = torch.ones(n, device=dev, dtype=f32), torch.zeros(n, device=dev, dtype=f32)
        hp = e((R, 512), bf)
        _lib.call("mvuld_unbatch_pad_bn", acts[-1], offsets, ones, zeros, hp, None, B, n, 512)
        hpn = e((R, 512), bf)
        gat_mean, gat_rstd = e((n,), f32), e((n,), f32)
        _lib.call("mvuld_bn_slot_fwd", hp, pv("bn_gat.weight"), pv("bn_gat.bias"), float(m.bn_gat.eps), hpn, gat_mean,
                  gat_rstd, m.bn_gat.running_mean, m.bn_gat.running_var, mom, B, n, 512)
        m.bn_gat.num_batches_tracked += 1
        z32, zb0 = e((R, 512), f32), e((R, 512), bf)
        _lib.gemm(hpn, w16("fc_gat.weight"), bias=pv("fc_gat.bias"), act=_lib.ACT_ELU, out_bf16=zb0, out_f32=z32)
        box_scale, box_shift, box_mean, box_rstd = e((n,), f32), e((n,), f32), e((n,), f32), e((n,), f32)
        _lib.call("mvuld_pos_slot_stats", pos, offsets, pv("bn_bbox.weight"), pv("bn_bbox.bias"), float(m.bn_bbox.eps),
                  box_scale, box_shift, box_mean, box_rstd, m.bn_bbox.running_mean, m.bn_bbox.running_var, mom, B, n)
        m.bn_bbox.num_batches_tracked += 1
        _lib.call("mvuld_pos_branch", pos, offsets, box_scale, box_shift, pv("fc_bbox.weight"), pv("fc_bbox.bias"),
                  z32, zb0, B, n, 32, 512, 480)

        if self.debug_taps is not None:
            self.debug_taps.update(node_mlp=acts[-1].clone(), gcn_in=z32.clone())
        gcn_saved = []
        for k in range(1, 9):
            # fp32-class block (bf16x3 split operands, see mvuld_rs_gcn_affinity_f32): Rs_GCN.py:52-73 in train mode
            pre = f"Rs_GCN_{k}."
            z3, y3 = e((R, 1536), bf), e((R, 1536), bf)
            tpg32, tpg, w0 = e((R, 1536), f32), e((R, 1536), bf), e((R, 512), f32)
            _lib.call("mvuld_split3_bf16", z32, 512, z3, R, 512, 0)
            _lib.gemm(z3, self.w3[f"gcn{k}.cat"], bias=self._gcn_cat(P, k, "bias"), out_bf16=tpg, out_f32=tpg32)
            _lib.call("mvuld_rs_gcn_affinity_f32", tpg32, y3, None, B, n, 512)
            _lib.gemm(y3, self.w3[f"gcn{k}.W0"], bias=pv(pre + "W.0.bias"), out_f32=w0)
            mean, rstd = bn_cols(pre + "W.1", w0, R, 512, res=z32, ldr=512, y32=z32, ldy=512, yb=None)
            gcn_saved.append((z3[:, :512], tpg, y3[:, :512], w0, mean, rstd))
            if self.debug_taps is not None:
                self.debug_taps[f"gcn_{k}"] = z32.clone()
        inv_s = e((B, 512), f32)
        _lib.call("mvuld_l2norm_mean_fwd", z32, _lib._Raw(feats[:, 512:1024]), 1536, inv_s, B, n, 512)
        fn = e((B, 1536), f32)
        fin_mean, fin_rstd = bn_cols("final_fc_bn", feats, B, 1536, y32=fn)
        C = m.num_classes
        logits = e((B, C), f32)
        _lib.call("mvuld_linear_small", fn, pv("final_fc.weight"), pv("final_fc.bias"), logits, None, B, C, 1536)
        self.last = dict(zero_deg=zero_deg, graph=g, feats=feats)
        ctx = dict(B=B, N=N, n=n, R=R, E=E, C=C, seed_base=seed_base, p=p, feats=feats, fn=fn, fin=(fin_mean, fin_rstd),
                   img=(img32, img_n, img_stat), txt=(txt32, txt_n, txt_stat), z32=z32, inv_s=inv_s, gcn_saved=gcn_saved,
                   zb0=zb0, hpn=hpn, hp=hp, gat_stat=(gat_mean, gat_rstd), pos=pos, offsets=offsets,
                   box_stat=(box_mean, box_rstd), acts=acts, gat_saved=gat_saved, csr=(indptr, idx_src),
                   out_csr=(out_indptr, out_dst, pos_in))
        return logits, ctx

    @torch.no_grad()
    def backward_train(self, ctx: dict, dlogits: torch.Tensor, G: torch.Tensor, on_bucket=None,
                       input_grads: bool = False):
        """Backward of ``forward_train`` from ``dlogits`` [B, num_classes] fp32: parameter gradients are WRITTEN into
        the flat buffer ``G`` (same layout as ``flat_p``; it must be zero on entry).  With ``input_grads`` also returns
        (d img_embedding [B, 1024], d func_text_embedding [B, 768]) fp32 -- what trainable encoders consume."""
        m, dev = self.model, self.dev
        B, N, n, R, E, C = ctx["B"], ctx["N"], ctx["n"], ctx["R"], ctx["E"], ctx["C"]
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        P = self.flat_p
        pv = lambda name: self._view(P, name)
        gv = lambda name: self._view(G, name)
        p = ctx["p"]
        _seed = lambda layer: (ctx["seed_base"] + layer) & 0x7FFFFFFFFFFFFFFF
        feats, fn, (fin_mean, fin_rstd) = ctx["feats"], ctx["fn"], ctx["fin"]
        (img32, img_n, img_stat), (txt32, txt_n, txt_stat) = ctx["img"], ctx["txt"]
        z32, inv_s, gcn_saved, zb0, hpn, hp = ctx["z32"], ctx["inv_s"], ctx["gcn_saved"], ctx["zb0"], ctx["hpn"], ctx["hp"]
        (gat_mean, gat_rstd), pos, offsets, (box_mean, box_rstd) = ctx["gat_stat"], ctx["pos"], ctx["offsets"], ctx["box_stat"]
        acts, gat_saved, (indptr, idx_src), (out_indptr, out_dst, pos_in) = ctx["acts"], ctx["gat_saved"], ctx["csr"], ctx["out_csr"]
        dlogits = dlogits.to(torch.float32).contiguous()
        bucket_i = 0
        d_inputs = []

        def ready(name):
            """Every gradient up to and including ``name`` is final: hand closed buckets to the reducer."""
            nonlocal bucket_i
            if on_bucket is None:
                return
            end = self.offsets[name] + (self._view(P, name).numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            while bucket_i < len(self.buckets) and self.buckets[bucket_i][1] <= end:
                on_bucket(self.buckets[bucket_i])
                bucket_i += 1

        dfn = e((B, 1536), f32)
        _lib.call("mvuld_linear_small_bwd", fn, pv("final_fc.weight"), dlogits, dfn, gv("final_fc.weight"),
                  gv("final_fc.bias"), B, C, 1536)
        dfeats = e((B, 1536), f32)
        _lib.call("mvuld_bn_cols_bwd", feats, dfn, 1536, pv("final_fc_bn.weight"), fin_mean, fin_rstd, dfeats, None,
                  gv("final_fc_bn.weight"), gv("final_fc_bn.bias"), B, 1536)
        ready("final_fc_bn.bias")

        # image / text projections: ELU(fc(bn(.)))
        for (lin, bnname, col0, xin, xn, K, stat) in (("swinfc", "swinbn", 0, img32, img_n, 1024, img_stat),
                                                      ("fc_text", "bn_text", 1024, txt32, txt_n, 768, txt_stat)):
            dpre = e((B, 512), bf)
            _lib.call("mvuld_elu_bwd_rows", _lib._Raw(dfeats[:, col0:col0 + 512]), 1536,
                      _lib._Raw(feats[:, col0:col0 + 512]), 1536, dpre, 512, B, 512)
            self._grad_w(dpre, xn, lin + ".weight", G=G)
            _lib.colsum(dpre, 1, 512, gv(lin + ".bias"), B, 512)
            dxn = e((B, K), f32)
            _lib.gemm(dpre, self.wt[lin][:, :512], out_f32=dxn)
            dxin = e((B, K), f32) if input_grads else None
            _lib.call("mvuld_bn_cols_bwd", xin, dxn, K, pv(bnname + ".weight"), stat[0], stat[1], dxin, None,
                      gv(bnname + ".weight"), gv(bnname + ".bias"), B, K)
            d_inputs.append(dxin)
        ready("bn_text.bias")

        # l2norm + mean, then the eight Rs_GCN blocks in reverse
        dz32, dzb = e((R, 512), f32), e((R, 512), bf)
        _lib.call("mvuld_l2norm_mean_bwd", z32, inv_s, _lib._Raw(dfeats[:, 512:1024]), 1536, dz32, dzb, B, n, 512)
        for k in range(8, 0, -1):
            pre = f"Rs_GCN_{k}."
            zin, tpg, y, w0, mean, rstd = gcn_saved[k - 1]
            dw0 = e((R, 512), bf)
            _lib.call("mvuld_bn_cols_bwd", w0, dz32, 512, pv(pre + "W.1.weight"), mean, rstd, None, dw0,
                      gv(pre + "W.1.weight"), gv(pre + "W.1.bias"), R, 512)
            self._grad_w(dw0, y, pre + "W.0.weight", G=G)
            _lib.colsum(dw0, 1, 512, gv(pre + "W.0.bias"), R, 512)
            dy = e((R, 512), bf)
            _lib.gemm(dw0, self.wt[f"gcn{k}.W0"][:, :512], out_bf16=dy)
            dtpg = e((R, 1536), bf)
            _lib.call("mvuld_rs_gcn_affinity_bwd", tpg, dy, dtpg, B, n, 512)
            self._grad_w(dtpg, zin, "", out=self._gcn_cat(G, k, "weight"), G=G)
            _lib.colsum(dtpg, 1, 1536, self._gcn_cat(G, k, "bias"), R, 1536)
            _lib.gemm(dtpg, self.wt[f"gcn{k}.cat"][:, :1536], res=dz32, out_bf16=dzb, out_f32=dz32)   # + residual path
            ready(pre + "g.bias")

        if self.debug_taps is not None:
            self.debug_taps["d_gcn_in"] = dz32.clone()
        # concat(h_i, pos_i) -> ELU -> fc_gat / fc_bbox -> slot BatchNorms -> unbatch
        dpre = e((R, 512), bf)
        _lib.call("mvuld_elu_bwd", dzb, zb0, dpre, R * 512, 0, 0, 0.0)
        self._grad_w(dpre, hpn, "fc_gat.weight", rows_out=480, G=G)
        _lib.colsum(dpre, 1, 512, gv("fc_gat.bias"), R, 480)
        dhpn = e((R, 512), bf)
        _lib.gemm(dpre[:, :480], self.wt["fc_gat"][:, :480], out_bf16=dhpn)
        dhp = e((R, 512), bf)
        _lib.call("mvuld_bn_slot_bwd", hp, dhpn, pv("bn_gat.weight"), gat_mean, gat_rstd, dhp, gv("bn_gat.weight"),
                  gv("bn_gat.bias"), B, n, 512)
        _lib.call("mvuld_pos_branch_bwd", pos, offsets, box_mean, box_rstd, pv("bn_bbox.weight"), pv("bn_bbox.bias"),
                  pv("fc_bbox.weight"), dpre, gv("fc_bbox.weight"), gv("fc_bbox.bias"), gv("bn_bbox.weight"),
                  gv("bn_bbox.bias"), e((n, 160), f32), B, n, 32, 512, 480)
        dh = e((N, 512), bf)
        _lib.call("mvuld_unbatch_pad_bwd", dhp, offsets, dh, B, n, 512)
        ready("bn_bbox.bias")

        # node MLP in reverse: hidden.7 ... hidden.0, fc
        mlp = ["fc"] + [f"hidden.{i}" for i in range(8)]
        for li in range(8, -1, -1):
            name = mlp[li]
            dpre = e((N, 512), bf)
            _lib.call("mvuld_elu_bwd", dh, acts[li + 1], dpre, N * 512, 0, _seed(8 + li), p)
            self._grad_w(dpre, acts[li], name + ".weight", G=G)
            _lib.colsum(dpre, 1, 512, gv(name + ".bias"), N, 512)
            K = acts[li].shape[1]
            dh = e((N, K), bf)
            _lib.gemm(dpre, self.wt[name][:, :512], out_bf16=dh)
            ready(name + ".bias")

        # GATConv x2 in reverse
        for li in (1, 0):
            name = ("gat", "gat2")[li]
            xd, z, el, er, H, F, slope = gat_saved[li]
            _lib.colsum(dh, 1, H * F, gv(name + ".bias"), N, H * F)
            alpha_e, ds_e = e((E, H), f32), e((E, H), f32)
            dl, dr = e((N, H), f32), e((N, H), f32)
            dz = e((N, H * F), bf)
            _lib.call("mvuld_gat_bwd", z, dh, el, er, indptr, idx_src, out_indptr, out_dst, pos_in,
                      pv(name + ".attn_l").view(-1), pv(name + ".attn_r").view(-1), alpha_e, ds_e, dl, dr, dz,
                      gv(name + ".attn_l").view(-1), gv(name + ".attn_r").view(-1), N, H, F, slope)
            self._grad_w(dz, xd, name + ".fc.weight", G=G)
            if li == 1:
                dxd = e((N, xd.shape[1]), bf)
                _lib.gemm(dz, self.wt["gat2.fc"][:, :H * F], out_bf16=dxd)
                if p > 0:
                    _lib.call("mvuld_dropout_bf16", dxd, dxd, dxd.numel(), _seed(li), p)
                dh = dxd
            ready(name + ".fc.weight")
        if on_bucket is not None:
            while bucket_i < len(self.buckets):
                on_bucket(self.buckets[bucket_i])
                bucket_i += 1
        if input_grads:
            return d_inputs[0], d_inputs[1]
        return None

    # ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, g: Graph, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor, targets: torch.Tensor,
             lr: Optional[float] = None, check: bool = True):
        """One optimiser step of main_bigvul.py:306-342.  Returns (loss, logits): loss is a [1] fp32 device tensor
        holding this rank's mean cross-entropy (no host synchronisation unless ``check``)."""
        self.step_count += 1                   # counts THIS step: dropout seeds and AdamW's bias correction use it
        works = []
        if self.world > 1:
            import torch.distributed as dist
            on_bucket = lambda b: works.append(dist.all_reduce(self.flat_g[b[0]:b[1]], op=dist.ReduceOp.SUM,
                                                               group=self.group, async_op=True))
        else:
            on_bucket = None
        try:
            loss, logits = self.forward_backward(g, img_embedding, func_text_embedding, targets, on_bucket)
        except Exception:
            self.step_count -= 1               # a rejected batch must not advance the schedule or the seeds
            raise
        for w in works:
            w.wait()
        self.apply_update(lr, advance=False)
        if self.world > 1:
            loss = loss * self.world           # forward_backward scaled the cross-entropy by 1 / world
        if check:
            if int(self.last["zero_deg"].item()) != 0:
                raise RuntimeError("There are 0-in-degree nodes in the graph (GATConv allow_zero_in_degree=False); "
                                   "add self-loops with mvuld_b200.graph.add_self_loop")
            self.last["graph"].check_status()
        return loss, logits

    @torch.no_grad()
    def apply_update(self, lr: Optional[float] = None, advance: bool = True):
        """``clip_grad_norm_(clip)`` + AdamW on whatever ``flat_g`` holds (utils_multi.py:225-240, optimizer.py:28-30), then
        the bf16 / transposed weight copies.  ``advance``: count a new optimiser step first (``step`` has already)."""
        if advance:
            self.step_count += 1
        self.gnorm_sq.zero_()
        _lib.call("mvuld_sumsq_f32", self.flat_g, self.total, self.gnorm_partials, self.gnorm_sq)
        _lib.call("mvuld_adamw", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.total, self.seg_end,
                  self.seg_wd, int(self.seg_end.numel()), self.gnorm_sq, self.clip, float(self.lr if lr is None else lr),
                  float(self.betas[0]), float(self.betas[1]), self.eps, self.step_count)
        self._refresh_shadows()

    @torch.no_grad()
    def sync_replicas(self, src: int = 0):
        """Data-parallel start-up / resume: every rank takes rank ``src``'s parameters, AdamW moments, step count and
        BatchNorm running statistics, as DDP's constructor broadcast does for the module state
        (main_bigvul.py:162-164).  A no-op on one GPU."""
        if self.world <= 1:
            return
        import torch.distributed as dist
        meta = torch.tensor([float(self.step_count)], device=self.dev, dtype=torch.float64)
        for t in (self.flat_p, self.flat_m, self.flat_v, meta):
            dist.broadcast(t, src=src, group=self.group)
        self.step_count = int(meta.item())
        for b in self.model.buffers():
            if b.is_floating_point() or b.dtype in (torch.int64, torch.int32):
                dist.broadcast(b, src=src, group=self.group)
        self._refresh_shadows()

    def refresh(self):
        """Re-derive the bf16 / transposed / split weight copies after the parameters were written from outside
        (``model.load_state_dict`` copies into the flat buffer the parameters are views of); under data parallelism the
        replicas are re-synchronised from rank 0 first (a checkpoint loaded on one rank only must not fork the replicas)."""
        if self.world > 1:
            self.sync_replicas()
        else:
            self._refresh_shadows()

    # ------------------------------------------------------------------------------------------------
    # optimiser state in torch.optim.AdamW's own layout, so checkpoints move both ways between this trainer and the
    # reference's ``optimizer.state_dict()`` / ``optimizer.load_state_dict()`` (utils_multi.py:7-32,125-137)
    # ------------------------------------------------------------------------------------------------
    def _torch_param_index(self) -> Tuple[Dict[str, int], int, int]:
        """Parameter name -> index in ``build_optimizer``'s groups (optimizer.py:35-50): every ``requires_grad``
        parameter in ``named_parameters()`` order, the decayed group first, then the 1-D / bias group."""
        decay, no_decay = [], []
        for n, prm in self.model.named_parameters():
            if not prm.requires_grad:
                continue
            (no_decay if (prm.dim() == 1 or n.endswith(".bias")) else decay).append(n)
        return {n: i for i, n in enumerate(decay + no_decay)}, len(decay), len(no_decay)

    def state_dict(self) -> dict:
        """``torch.optim.AdamW.state_dict()`` layout: ``state[idx] = {step, exp_avg, exp_avg_sq}`` for every parameter that
        has received a gradient, ``param_groups`` = [decayed, not decayed].  Trainer-only settings ride along under the
        extra key ``mvuld_b200`` (``Optimizer.load_state_dict`` reads ``state`` and ``param_groups`` only)."""
        index, n_decay, n_no = self._torch_param_index()
        state = {}
        if self.step_count > 0:
            m, v = self.flat_m.detach().cpu(), self.flat_v.detach().cpu()
            for n in self.names:
                state[index[n]] = {"step": torch.tensor(float(self.step_count)),
                                   "exp_avg": self._view(m, n).clone(), "exp_avg_sq": self._view(v, n).clone()}
        common = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, amsgrad=False, maximize=False, foreach=None,
                      capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True)
        groups = [dict(common, weight_decay=self.wd, params=list(range(n_decay))),
                  dict(common, weight_decay=0.0, params=list(range(n_decay, n_decay + n_no)))]
        return {"state": state, "param_groups": groups,
                "mvuld_b200": dict(kind="mvuld_b200.FusionTrainer", step=self.step_count, clip_grad=self.clip,
                                   dropout=self.p_drop, seed=self.seed)}

    def load_state_dict(self, sd: dict):
        """Accepts a ``torch.optim.AdamW`` state dict over the same model (the reference's checkpoints) or one written
        by ``state_dict()`` above."""
        if "param_groups" not in sd or "state" not in sd:
            raise ValueError("optimizer state is not a torch.optim.AdamW state dict ({'state', 'param_groups'})")
        index, n_decay, n_no = self._torch_param_index()
        groups = sd["param_groups"]
        if len(groups) != 2 or len(groups[0]["params"]) != n_decay or len(groups[1]["params"]) != n_no:
            raise ValueError("optimizer state does not belong to an AdamW over this model's parameter groups "
                             f"(expected {n_decay} decayed + {n_no} non-decayed parameters)")
        ids = list(groups[0]["params"]) + list(groups[1]["params"])          # position -> saved id
        g0 = groups[0]
        self.lr, self.wd, self.betas, self.eps = float(g0["lr"]), float(g0["weight_decay"]), tuple(g0["betas"]), float(g0["eps"])
        self.flat_m.zero_()
        self.flat_v.zero_()
        steps = set()
        for n in self.names:
            st = sd["state"].get(ids[index[n]])
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != self.shapes[n]:
                raise ValueError(f"optimizer state of {n}: shape {tuple(st['exp_avg'].shape)} != {self.shapes[n]}")
            self._view(self.flat_m, n).copy_(st["exp_avg"].to(torch.float32))
            self._view(self.flat_v, n).copy_(st["exp_avg_sq"].to(torch.float32))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"optimizer state holds different step counts per parameter ({sorted(steps)}); the flat "
                             "AdamW launch applies one bias correction")
        extra = sd.get("mvuld_b200", {})
        self.step_count = int(extra.get("step", steps.pop() if steps else 0))
        self.clip = float(extra.get("clip_grad", self.clip))
        self.p_drop = float(extra.get("dropout", self.p_drop))
        self.seed = int(extra.get("seed", self.seed))
        wd = [0.0 if (len(self.shapes[n]) == 1 or n.endswith(".bias")) else self.wd for n in self.names]
        self.seg_wd.copy_(torch.tensor(wd, dtype=torch.float32))

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (what clip_grad_norm_ returns, utils_multi.py:233)."""
        return self.gnorm_sq.sqrt()

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {n: self._view(self.flat_g, n) for n in self.names}
