"""Autograd boundary of the fusion model (last row).

The reference trains ``Multi_DefectModel_new_GCN`` through plain autograd (/root/reference/mvuld/main_bigvul.py:328-342:
``outputs = model(g, img_embedding, func_text_embedding)``; ``loss = criterion(outputs, targets) / ACCUMULATION_STEPS``;
``loss_scaler(loss, optimizer, clip_grad=..., parameters=model.parameters(), update_grad=...)`` ->
``GradScaler.scale(loss).backward()``, ``clip_grad_norm_``, ``optimizer.step()``, utils_multi.py:225-240).  Here the
whole train-mode forward is ONE autograd node: its forward is ``FusionTrainer.forward_train`` (a fixed sequence of
C-ABI launches), its backward ``FusionTrainer.backward_train`` -- the same kernels the fast path
(``FusionTrainer.step``) runs.  Parameter gradients come back as views of a fresh flat fp32 buffer in the layout of the
flat parameter buffer, so ``param.grad`` accumulates over ``ACCUMULATION_STEPS`` exactly as with eager modules, the
parameters without a gradient in the reference forward (the dead ``h_func`` branch) get ``None``, and the image / text
embeddings receive gradients when they require them (trainable encoders, BASELINE.json configs[4] primary reading).
No CPU path: CUDA tensors only.
"""
from __future__ import annotations

import torch


class _FusionTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, g, seed_index, img_embedding, func_text_embedding, *params):
        logits, saved = engine.forward_train(g, img_embedding, func_text_embedding, seed_index=seed_index)
        ctx.engine, ctx.saved = engine, saved
        ctx.input_grads = bool(ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        ctx.n_params = len(params)
        return logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogits):
        eng = ctx.engine
        # a fresh buffer per backward pass: autograd may adopt the returned views as ``param.grad`` without a copy
        G = torch.zeros(eng.total, device=eng.dev, dtype=torch.float32)
        d_in = eng.backward_train(ctx.saved, dlogits, G, None, input_grads=ctx.input_grads)
        ctx.saved = None
        grads = tuple(eng._view(G, n) for n in eng.names)
        d_img = d_in[0] if (ctx.input_grads and ctx.needs_input_grad[3]) else None
        d_txt = d_in[1] if (ctx.input_grads and ctx.needs_input_grad[4]) else None
        return (None, None, None, d_img, d_txt) + grads


def fusion_train_forward(model, g, img_embedding: torch.Tensor, func_text_embedding: torch.Tensor) -> torch.Tensor:
    """Train-mode ``Multi_DefectModel_new_GCN.forward`` with an autograd graph (see the module docstring)."""
    eng = model.train_engine()
    params = dict(model.named_parameters())
    plist = [params[n] for n in eng.names]
    # an optimiser updates the parameters in place (they are views of the engine's flat fp32 buffer): refresh the
    # bf16 / transposed / split GEMM operands when any version counter moved
    versions = tuple(p._version for p in plist)
    if versions != getattr(eng, "_seen_versions", None):
        eng.refresh()
        eng._seen_versions = tuple(p._version for p in plist)
    eng.autograd_calls = getattr(eng, "autograd_calls", 0) + 1
    model._plan = None                                   # BatchNorm running statistics move: the eval plan is stale
    if not torch.is_grad_enabled():
        logits, _ = eng.forward_train(g, img_embedding, func_text_embedding, seed_index=eng.autograd_calls)
        return logits
    out = _FusionTrainFunction.apply(eng, g, eng.autograd_calls, img_embedding, func_text_embedding, *plist)
    if not model.defer_checks:
        if int(eng.last["zero_deg"].item()) != 0:
            raise RuntimeError("There are 0-in-degree nodes in the graph (GATConv allow_zero_in_degree=False); "
                               "add self-loops with mvuld_b200.graph.add_self_loop")
        eng.last["graph"].check_status()
    return out
