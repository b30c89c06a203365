"""Training step of the text encoder: RoBERTa-base (UniXcoder) forward + backward + AdamW as fixed launch sequences.

Reference: autograd through HF ``RobertaModel`` as ``/root/reference/mvuld/models/unixcoder.py:33-38`` calls it
(``encoder(ids, attention_mask=ids.ne(1))[0]``, masked mean over the valid tokens) -- the text half of BASELINE.json
configs[4] in its primary reading.  The backward pass is hand written on the same kernels as the image encoder's:
dense products and weight gradients on tcgen05 (``gemm.cu`` / ``gemm_dw.cu``), self-attention backward on tcgen05
(``attention_bwd.cu::seq_attn_bwd_kernel``), LayerNorm(x + residual) / GELU backward row kernels, embedding-table
gradients as ordered row sums.  Dropout 0 (as in the pinned oracle cases); padded ``[B, L]`` rows (pad tokens a suffix).
No CPU path: CUDA tensors only.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from .train import _ALIGN, plan_buckets, plan_layout
from .unixcoder import LOG2E, RobertaEncoder


def _parameter_order(enc: RobertaEncoder) -> List[str]:
    """Trained parameters in the order the backward pass completes them; query | key | value adjacent so that one
    [3H, H] view serves the fused qkv product.  The pooler takes no part in get_repr (no gradient, left out)."""
    names = []
    for i in range(len(enc.encoder.layer) - 1, -1, -1):
        p = f"encoder.layer.{i}."
        names += [p + "output.LayerNorm.weight", p + "output.LayerNorm.bias", p + "output.dense.weight",
                  p + "output.dense.bias", p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                  p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
                  p + "attention.output.dense.weight", p + "attention.output.dense.bias",
                  p + "attention.self.query.weight", p + "attention.self.key.weight", p + "attention.self.value.weight",
                  p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias"]
    names += ["embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias", "embeddings.token_type_embeddings.weight",
              "embeddings.position_embeddings.weight", "embeddings.word_embeddings.weight"]
    have = dict(enc.named_parameters())
    missing = [n for n in names if n not in have]
    if missing:
        raise KeyError(f"encoder lacks parameters {missing}")
    return names


class RobertaTrainer:
    """Owns flat parameter / gradient / AdamW buffers of a ``RobertaEncoder`` and runs training steps on it."""

    def __init__(self, encoder: RobertaEncoder, lr: float = 2e-5, weight_decay: float = 0.01,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, clip_grad: float = 1.0,
                 process_group=None, world_size: Optional[int] = None, bucket_mb: float = 25.0):
        dev = encoder.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 RobertaTrainer runs on CUDA only (no CPU fallback): move the model to the GPU")
        _lib.load()
        self.enc, self.cfg, self.dev = encoder, encoder.config, dev
        self.lr, self.wd, self.betas, self.eps, self.clip = float(lr), float(weight_decay), betas, float(eps), float(clip_grad)
        self.group = process_group
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world = int(world_size)
        self.step_count = 0
        params = dict(encoder.named_parameters())
        self.names = _parameter_order(encoder)
        self.offsets, self.total = plan_layout([(n, params[n].numel()) for n in self.names])
        self.shapes = {n: tuple(params[n].shape) for n in self.names}
        f32 = dict(device=dev, dtype=torch.float32)
        self.flat_p = torch.zeros(self.total, **f32)
        self.flat_g = torch.zeros(self.total, **f32)
        self.flat_m = torch.zeros(self.total, **f32)
        self.flat_v = torch.zeros(self.total, **f32)
        self.flat_w16 = torch.zeros(self.total, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            for n in self.names:
                view = self._view(self.flat_p, n)
                view.copy_(params[n].detach().float())
                params[n].data = view
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
        H = self.cfg.hidden_size
        if (H * H) % _ALIGN or H % _ALIGN:
            raise NotImplementedError("RobertaTrainer: hidden_size must be a multiple of 64 (fused qkv view of the flat buffer)")
        if self.world > 1:
            self.sync_replicas()
        else:
            self._refresh()

    # ------------------------------------------------------------------------------------------------
    def _view(self, flat: torch.Tensor, name: str, shape=None) -> torch.Tensor:
        shp = self.shapes[name] if shape is None else shape
        o = self.offsets[name]
        return flat[o:o + math.prod(shp)].view(shp)

    def _qkv(self, flat: torch.Tensor, i: int, what: str) -> torch.Tensor:
        """query | key | value of layer i laid out back to back: one [3H, H] weight / [3H] bias view."""
        H = self.cfg.hidden_size
        o = self.offsets[f"encoder.layer.{i}.attention.self.query.{what}"]
        return flat[o:o + 3 * H * H].view(3 * H, H) if what == "weight" else flat[o:o + 3 * H]

    def _transpose(self, x: torch.Tensor) -> torch.Tensor:
        R, C = x.shape
        Rp = (R + 7) // 8 * 8
        out = torch.empty(C, Rp, device=self.dev, dtype=torch.bfloat16)
        _lib.call("mvuld_transpose_bf16", _lib._Raw(x), x.stride(0), out, R, C, Rp)
        return out

    @torch.no_grad()
    def _refresh(self):
        """bf16 copies of the weight matrices (forward operands) and their transposes (dX = dY W)."""
        _lib.call("mvuld_f32_to_bf16", self.flat_p, self.flat_w16, self.total)
        if not hasattr(self, "_wt_table"):
            # the bf16 weights are views of flat_w16 (fixed addresses) and the transposed copies are allocated once: the
            # 48 transposes of a step are ONE launch over a device-side table
            self.w, self.wt = [], []
            for i in range(len(self.enc.encoder.layer)):
                p = f"encoder.layer.{i}."
                w = dict(qkv=self._qkv(self.flat_w16, i, "weight"),
                         o=self._view(self.flat_w16, p + "attention.output.dense.weight"),
                         i=self._view(self.flat_w16, p + "intermediate.dense.weight"),
                         o2=self._view(self.flat_w16, p + "output.dense.weight"))
                self.w.append(w)
                self.wt.append({k: torch.empty(v.shape[1], (v.shape[0] + 7) // 8 * 8, device=self.dev, dtype=torch.bfloat16)
                                for k, v in w.items()})
            self._wt_table = _lib.TransposeTable([(w[k], wt[k]) for w, wt in zip(self.w, self.wt) for k in w])
        self._wt_table.run()
        self.enc.invalidate()

    def refresh(self):
        if self.world > 1:
            self.sync_replicas()
        else:
            self._refresh()

    @torch.no_grad()
    def sync_replicas(self, src: int = 0):
        if self.world > 1:
            import torch.distributed as dist
            meta = torch.tensor([float(self.step_count)], device=self.dev, dtype=torch.float64)
            for t in (self.flat_p, self.flat_m, self.flat_v, meta):
                dist.broadcast(t, src=src, group=self.group)
            self.step_count = int(meta.item())
        self._refresh()

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_train(self, source_ids: torch.Tensor):
        """Training-mode ``get_xcode_vec`` (unixcoder.py:33-38) -> (sentence vectors fp32 [B, H], ctx)."""
        if not source_ids.is_cuda:
            raise RuntimeError("mvuld_b200 RobertaTrainer takes CUDA tensors (no CPU fallback)")
        cfg, dev = self.cfg, self.dev
        ids = source_ids.to(torch.int64).contiguous()
        B, L = ids.shape
        if L > 512 or L % 8 != 0:
            raise ValueError("sequence length must be a multiple of 8 and <= 512")
        H, nH, I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        M = B * L
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        pv = lambda n: self._view(self.flat_p, n)
        pos, length, ok = e((M,), torch.int32), e((B,), torch.int32), torch.ones(1, device=dev, dtype=torch.int32)
        _lib.call("mvuld_seq_positions", ids, B, L, int(cfg.pad_token_id), pos, length, ok)
        x32, xb, ysum = e((M, H), f32), e((M, H), bf), e((M, H), bf)
        _lib.call("mvuld_roberta_embed_train", ids, pos, pv("embeddings.word_embeddings.weight"),
                  pv("embeddings.position_embeddings.weight"), pv("embeddings.token_type_embeddings.weight")[0].contiguous(),
                  pv("embeddings.LayerNorm.weight"), pv("embeddings.LayerNorm.bias"), x32, xb, ysum, M, H,
                  float(cfg.layer_norm_eps))
        qmul = LOG2E / math.sqrt(H // nH)
        eps = float(cfg.layer_norm_eps)
        ctx = dict(B=B, L=L, ids=ids, pos=pos, len=length, ok=ok, ysum=ysum, layers=[])
        for i in range(len(self.enc.encoder.layer)):
            p, w = f"encoder.layer.{i}.", self.w[i]
            q, k, v = e((M, H), bf), e((M, H), bf), e((M, H), bf)
            _lib.call("mvuld_heads_qkv", xb, w["qkv"], self._qkv(self.flat_p, i, "bias"), q, k, v, B, L, H, nH, qmul)
            att, lse = e((M, H), bf), e((B * nH, L), f32)
            _lib.call("mvuld_seq_attention_train", q, k, v, length, att, lse, B, L, nH, H // nH)
            y1 = e((M, H), bf)
            _lib.gemm(att, w["o"], bias=pv(p + "attention.output.dense.bias"), out_bf16=y1)
            x32_1, xb1 = e((M, H), f32), e((M, H), bf)
            _lib.call("mvuld_ln_rows", y1, x32, pv(p + "attention.output.LayerNorm.weight"),
                      pv(p + "attention.output.LayerNorm.bias"), x32_1, xb1, M, H, eps, 2)
            pre, hid = e((M, I), bf), e((M, I), bf)
            _lib.gemm(xb1, w["i"], bias=pv(p + "intermediate.dense.bias"), out_bf16=pre)
            _lib.call("mvuld_gelu_fwd", pre, hid, M * I)
            y2 = e((M, H), bf)
            _lib.gemm(hid, w["o2"], bias=pv(p + "output.dense.bias"), out_bf16=y2)
            x32_2, xb2 = e((M, H), f32), e((M, H), bf)
            _lib.call("mvuld_ln_rows", y2, x32_1, pv(p + "output.LayerNorm.weight"), pv(p + "output.LayerNorm.bias"),
                      x32_2, xb2, M, H, eps, 2)
            ctx["layers"].append(dict(xb_in=xb, x32_in=x32, q=q, k=k, v=v, lse=lse, att=att, y1=y1, x32_1=x32_1, xb1=xb1,
                                      pre=pre, hid=hid, y2=y2))
            x32, xb = x32_2, xb2
        sent = e((B, H), f32)
        _lib.call("mvuld_masked_mean", x32, length, sent, B, L, H)
        return sent, ctx

    @torch.no_grad()
    def backward_train(self, ctx: dict, dsent: torch.Tensor, on_bucket=None):
        """Backward of ``forward_train`` for the cotangent ``dsent`` [B, H]; fills ``flat_g`` (zero it first)."""
        cfg, dev, G = self.cfg, self.dev, self.flat_g
        B, L = ctx["B"], ctx["L"]
        H, nH, I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        M = B * L
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        pv = lambda n: self._view(self.flat_p, n)
        gv = lambda n: self._view(G, n)
        eps = float(cfg.layer_norm_eps)
        done, bucket_i = set(), 0

        def ready(*names):
            nonlocal bucket_i
            if on_bucket is None:
                return
            done.update(names)
            while bucket_i < len(self.buckets):
                lo, hi = self.buckets[bucket_i]
                if not all(n in done for n in self.names if lo <= self.offsets[n] < hi):
                    break
                on_bucket(self.buckets[bucket_i])
                bucket_i += 1

        def ln_bwd(y, shortcut, gname, bname, dout, bias_name=None):
            """LN(y + shortcut) backward -> (dv fp32: gradient of the sum, i.e. of both addends; dv bf16); the gradient
            of the dense bias in front of the LayerNorm (column sums of dv) comes out of the same kernel."""
            dvb, dv32 = e((M, H), bf), e((M, H), f32)
            _lib.call("mvuld_ln_rows_bwd", y, shortcut, pv(gname), dout, dvb, dv32, gv(gname), gv(bname),
                      gv(bias_name) if bias_name else None, _lib.ln_rows_bwd_partials(M, H, dev), M, H, eps,
                      2 if shortcut is not None else 0)
            return dv32, dvb

        dx = e((M, H), f32)
        _lib.call("mvuld_masked_mean_bwd", dsent.to(f32).contiguous(), ctx["len"], dx, B, L, H)
        for i in range(len(self.enc.encoder.layer) - 1, -1, -1):
            p, s, wt = f"encoder.layer.{i}.", ctx["layers"][i], self.wt[i]
            # ---- x2 = LN(dense2(gelu(dense1(x1))) + x1) ----
            d32, db16 = ln_bwd(s["y2"], s["x32_1"], p + "output.LayerNorm.weight", p + "output.LayerNorm.bias", dx,
                               bias_name=p + "output.dense.bias")
            _lib.gemm_dw(db16, s["hid"], gv(p + "output.dense.weight"))
            dhid = e((M, I), bf)
            _lib.gemm(db16, wt["o2"], out_bf16=dhid)
            dpre = e((M, I), bf)
            _lib.gelu_bwd_colsum(s["pre"].view(M, I), dhid, dpre, gv(p + "intermediate.dense.bias"))
            _lib.gemm_dw(dpre, s["xb1"], gv(p + "intermediate.dense.weight"))
            _lib.gemm(dpre, wt["i"], res=d32, out_f32=d32)                      # dx1 = d(sum) + dpre W_i
            del dhid, dpre
            # ---- x1 = LN(dense(attention(x0)) + x0) ----
            d32b, db16 = ln_bwd(s["y1"], s["x32_in"], p + "attention.output.LayerNorm.weight",
                                p + "attention.output.LayerNorm.bias", d32, bias_name=p + "attention.output.dense.bias")
            _lib.gemm_dw(db16, s["att"], gv(p + "attention.output.dense.weight"))
            datt = e((M, H), bf)
            _lib.gemm(db16, wt["o"], out_bf16=datt)
            dOh, ld = e((M, H), bf), e((B * nH, L, 2), f32)
            _lib.call("mvuld_seq_attention_bwd_prep", datt, s["att"], s["lse"], dOh, ld, B, L, nH)
            dq = torch.zeros(B * nH, L, H // nH, device=dev, dtype=f32)
            dk, dv = torch.zeros_like(dq), torch.zeros_like(dq)
            _lib.call("mvuld_seq_attention_bwd", s["q"], s["k"], s["v"], dOh, ld, ctx["len"], dq, dk, dv, B, L, nH, H // nH)
            dqkv = e((M, 3 * H), bf)
            _lib.call("mvuld_seq_qkv_bwd", dq, dk, dv, dqkv, B, L, nH, H // nH)
            del dq, dk, dv, dOh
            _lib.gemm_dw(dqkv, s["xb_in"], self._qkv(G, i, "weight"))
            _lib.colsum(dqkv, 1, 3 * H, self._qkv(G, i, "bias"), M, 3 * H)
            _lib.gemm(dqkv, wt["qkv"], res=d32b, out_f32=d32b)                  # dx0 = d(sum) + dqkv W_qkv
            dx = d32b
            ctx["layers"][i] = None
            ready(*[n for n in self.names if n.startswith(p)])
        # ---- embeddings: LayerNorm(word[ids] + position[pos] + type[0]) ----
        dsum, _ = ln_bwd(ctx["ysum"], None, "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias", dx)
        pad = int(cfg.pad_token_id)
        rows = torch.arange(M, device=dev, dtype=torch.int64)
        for name, index, n_index in (("embeddings.word_embeddings.weight", ctx["ids"].view(-1), cfg.vocab_size),
                                     ("embeddings.position_embeddings.weight", ctx["pos"].to(torch.int64),
                                      cfg.max_position_embeddings)):
            indptr, order, _eids, _status = _lib.csr_from_coo(rows, index.contiguous(), int(n_index))
            _lib.call("mvuld_embed_grad_rows", dsum, indptr, order, gv(name), int(n_index), H, pad)
        _lib.colsum(dsum, 0, H, gv("embeddings.token_type_embeddings.weight")[0], M, H)
        ready(*[n for n in self.names if n.startswith("embeddings.")])
        if on_bucket is not None:
            while bucket_i < len(self.buckets):
                on_bucket(self.buckets[bucket_i])
                bucket_i += 1

    @torch.no_grad()
    def apply_update(self, lr: Optional[float] = None, advance: bool = True):
        if advance:
            self.step_count += 1
        self.gnorm_sq.zero_()
        _lib.call("mvuld_sumsq_f32", self.flat_g, self.total, self.gnorm_partials, self.gnorm_sq)
        _lib.call("mvuld_adamw", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.total, self.seg_end,
                  self.seg_wd, int(self.seg_end.numel()), self.gnorm_sq, self.clip, float(self.lr if lr is None else lr),
                  float(self.betas[0]), float(self.betas[1]), self.eps, self.step_count)
        self._refresh()

    def grad_norm(self) -> torch.Tensor:
        return self.gnorm_sq.sqrt()

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {n: self._view(self.flat_g, n) for n in self.names}
