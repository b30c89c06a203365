"""Joint training step of the composed model (BASELINE.json configs[4], primary reading: the encoders train too).

The reference trains its pieces in separate stages -- the SwinV2 image classifier (/root/reference/mvuld/main.py:251-300),
the fusion model on cached encoder vectors (mvuld/main_bigvul.py:294-342) -- because it never runs the modalities in one
pass.  With the encoders on the same GPU the two backward passes chain:

    img = swin.forward_features(image)                       (SwinTrainer.forward_train, activations kept)
    txt = unix.get_repr(token_ids)[0]                        (RobertaTrainer.forward_train, activations kept;
                                                              ``train_text=False``: frozen, eval forward on packed rows)
    logits = fusion(g, img, txt); loss = CE(logits, y)       (FusionTrainer.forward_train)
    d img, d txt = fusion backward (input gradients)  ->  SwinV2 backward, RoBERTa backward
    one clip_grad_norm_ over ALL parameter sets (fixed summation order), AdamW per parameter set (the fusion model's
    decay 0.005 / the image encoder's 0.05, config.py TRAIN.WEIGHT_DECAY / the text encoder's 0.01)

Data parallel: each trainer's flat fp32 gradient buffer is all-reduced bucket by bucket while the backward pass runs.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from .roberta_train import RobertaTrainer
from .swin_train import SwinTrainer
from .train import FusionTrainer


class MVulDTrainer:
    def __init__(self, model, lr: float = 5e-5, fusion_weight_decay: float = 0.005, swin_weight_decay: float = 0.05,
                 clip_grad: float = 5.0, dropout: Optional[float] = None, seed: int = 12345, process_group=None,
                 world_size: Optional[int] = None, bucket_mb: float = 25.0, train_text: bool = True,
                 text_weight_decay: float = 0.01):
        self.model = model
        self.fusion = FusionTrainer(model.fusion, lr=lr, weight_decay=fusion_weight_decay, clip_grad=clip_grad,
                                    dropout=dropout, seed=seed, process_group=process_group, world_size=world_size,
                                    bucket_mb=bucket_mb)
        self.swin = SwinTrainer(model.swin, lr=lr, weight_decay=swin_weight_decay, clip_grad=clip_grad,
                                process_group=process_group, world_size=world_size, bucket_mb=bucket_mb)
        self.text = RobertaTrainer(model.unix.encoder, lr=lr, weight_decay=text_weight_decay, clip_grad=clip_grad,
                                   process_group=process_group, world_size=world_size, bucket_mb=bucket_mb) \
            if train_text else None
        self.trainers = [t for t in (self.fusion, self.swin, self.text) if t is not None]
        self.world, self.group, self.clip, self.lr = self.fusion.world, process_group, float(clip_grad), float(lr)
        self.dev = self.fusion.dev
        self.gnorm_sq = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self.loss_buf = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self.step_count = 0

    @property
    def num_parameters(self) -> int:
        return sum(math.prod(t.shapes[n]) for t in self.trainers for n in t.names)

    @property
    def buckets(self):
        return [b for t in self.trainers for b in t.buckets]

    @torch.no_grad()
    def forward_backward(self, g, image: torch.Tensor, token_ids, targets: torch.Tensor, works=None):
        if not targets.is_cuda:
            raise RuntimeError("mvuld_b200 MVulDTrainer takes CUDA tensors (no CPU fallback)")
        f, s, t = self.fusion, self.swin, self.text
        bucket_cb = None
        if self.world > 1 and works is not None:
            import torch.distributed as dist

            def make(flat):
                return lambda b: works.append(dist.all_reduce(flat[b[0]:b[1]], op=dist.ReduceOp.SUM, group=self.group,
                                                              async_op=True))
            bucket_cb = (make(f.flat_g), make(s.flat_g), make(t.flat_g) if t is not None else None)
        img, sctx = s.forward_train(image)
        if t is not None:
            if not torch.is_tensor(token_ids):
                raise TypeError("MVulDTrainer(train_text=True) takes the tokenizer's [B, L] id tensor (padded rows)")
            txt, tctx = t.forward_train(token_ids)
        else:
            txt, _ = self.model.unix.get_repr(token_ids)
        logits, fctx = f.forward_train(g, img, txt)
        B, C = logits.shape
        dlogits = torch.empty((B, C), device=self.dev, dtype=torch.float32)
        self.loss_buf.zero_()
        _lib.call("mvuld_ce_loss", logits, targets.to(torch.int64).contiguous(), self.loss_buf, dlogits, B, C,
                  1.0 / (B * self.world))
        f.flat_g.zero_()
        s.flat_g.zero_()
        d_img, d_txt = f.backward_train(fctx, dlogits, f.flat_g, bucket_cb[0] if bucket_cb else None, input_grads=True)
        if t is not None:                       # the shorter backward first: its buckets reduce under the SwinV2 backward
            t.flat_g.zero_()
            t.backward_train(tctx, d_txt, bucket_cb[2] if bucket_cb else None)
        s.backward_train(sctx, d_img, bucket_cb[1] if bucket_cb else None)
        return self.loss_buf, logits

    @torch.no_grad()
    def step(self, g, image: torch.Tensor, token_ids, targets: torch.Tensor, lr: Optional[float] = None):
        """One optimiser step.  Returns (loss [1] fp32 device tensor of this rank, logits)."""
        f, s = self.fusion, self.swin
        self.step_count += 1
        for t in self.trainers:
            t.step_count = self.step_count
        works = []
        try:
            loss, logits = self.forward_backward(g, image, token_ids, targets, works)
        except Exception:
            self.step_count -= 1
            for t in self.trainers:
                t.step_count = self.step_count
            raise
        for w in works:
            w.wait()
        # clip_grad_norm_ over every trained parameter: one norm from both flat buffers (fixed summation order)
        self.gnorm_sq.zero_()
        for t in self.trainers:
            _lib.call("mvuld_sumsq_f32", t.flat_g, t.total, t.gnorm_partials, self.gnorm_sq)
        cur = float(self.lr if lr is None else lr)
        for t in self.trainers:
            _lib.call("mvuld_adamw", t.flat_p, t.flat_g, t.flat_m, t.flat_v, t.total, t.seg_end, t.seg_wd,
                      int(t.seg_end.numel()), self.gnorm_sq, self.clip, cur, float(t.betas[0]), float(t.betas[1]), t.eps,
                      self.step_count)
        f._refresh_shadows()
        s._refresh()
        if self.text is not None:
            self.text._refresh()
        return (loss * self.world if self.world > 1 else loss), logits

    def grad_norm(self) -> torch.Tensor:
        return self.gnorm_sq.sqrt()
