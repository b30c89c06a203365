"""GPU: the reference training loop body run UNCHANGED against ``mvuld_b200.Multi_DefectModel_new_GCN``.

/root/reference/mvuld/main_bigvul.py:294-342 (``train_one_epoch``): ``model.train()``, ``optimizer.zero_grad()``,
``outputs = model(g, img_embedding, func_text_embedding)`` under ``autocast(enabled=False)``,
``loss = criterion(outputs, targets) / ACCUMULATION_STEPS``, then ``loss_scaler(loss, optimizer, clip_grad=...,
parameters=model.parameters(), update_grad=(idx + 1) % ACCUMULATION_STEPS == 0)`` -- utils_multi.py:225-240
(``GradScaler.scale(loss).backward()``, ``unscale_``, ``clip_grad_norm_``, ``scaler.step(optimizer)``).  The loop body
below is that code; ``LossScaler`` restates the 15-line scaler class (test infrastructure).  The gradients autograd
accumulates are compared with the flat-buffer fast path (``FusionTrainer.forward_backward``, itself checked against the
fp32 autograd oracle in tests/test_gpu_train.py), the update with ``torch.optim.AdamW`` semantics.
"""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import mvuld_b200 as mv                          # noqa: E402
from mvuld_b200 import synth, train             # noqa: E402
from tests import cases                          # noqa: E402

DEV = "cuda"


class LossScaler:
    """utils_multi.py:219-247 (NativeScalerWithGradNormCount), restated."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scaler = torch.amp.GradScaler("cuda")

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if update_grad:
            self._scaler.unscale_(optimizer)
            if clip_grad is not None:
                norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
            else:
                norm = torch.norm(torch.stack([torch.norm(p.grad.detach(), 2.0) for p in parameters if p.grad is not None]), 2.0)
            self._scaler.step(optimizer)
            self._scaler.update()
        else:
            norm = None
        return norm

    def state_dict(self):
        return self._scaler.state_dict()


def _batches(n, B=6):
    out = []
    for i in range(n):
        g = synth.cpg_batch(B, seed=cases.SEED + i)
        r = torch.Generator().manual_seed(cases.SEED + 9 + i)
        img, txt = torch.randn(B, 1024, generator=r), torch.randn(B, 768, generator=r) * 0.5
        out.append((g, img, txt, torch.randint(0, 2, (B,), generator=r)))
    return out


def _build_optimizer(model, lr, wd):
    """optimizer.py:11-50: AdamW, 1-D tensors and biases not decayed."""
    has_decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (len(p.shape) == 1 or name.endswith(".bias")) else has_decay).append(p)
    return torch.optim.AdamW([{"params": has_decay}, {"params": no_decay, "weight_decay": 0.}], eps=1e-8,
                             betas=(0.9, 0.999), lr=lr, weight_decay=wd)


def test_reference_loop_body_with_accumulation_matches_fast_path():
    config = types.SimpleNamespace(TRAIN=types.SimpleNamespace(ACCUMULATION_STEPS=2, CLIP_GRAD=5.0))
    model = cases.make_fusion().to(DEV)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    optimizer = _build_optimizer(model, lr=1e-3, wd=0.005)
    loss_scaler = LossScaler()
    model.train_engine().p_drop = 0.0                       # deterministic comparison (dropout is tested on its own)
    data_loader = _batches(2)
    seen = {}

    # ------------------------- main_bigvul.py:296-342, loop body verbatim -------------------------
    model.train()
    optimizer.zero_grad()
    criterion = torch.nn.CrossEntropyLoss()
    for idx, (g, img_embedding, func_text_embedding, target) in enumerate(data_loader):
        cuda = next(model.parameters()).device
        g = g.to(cuda)
        img_embedding = img_embedding.cuda(non_blocking=True)
        func_text_embedding = func_text_embedding.cuda(non_blocking=True)
        targets = target.cuda(non_blocking=True)
        with torch.cuda.amp.autocast(enabled=False):
            outputs = model(g, img_embedding, func_text_embedding)
            probs = F.softmax(outputs, dim=1)
        loss = criterion(outputs, targets)
        loss = loss / config.TRAIN.ACCUMULATION_STEPS
        is_second_order = hasattr(optimizer, 'is_second_order') and optimizer.is_second_order
        if (idx + 1) % config.TRAIN.ACCUMULATION_STEPS == 0:     # (test hook: keep the accumulated gradients)
            pass
        grad_norm = loss_scaler(loss, optimizer, clip_grad=config.TRAIN.CLIP_GRAD,
                                parameters=model.parameters(), create_graph=is_second_order,
                                update_grad=(idx + 1) % config.TRAIN.ACCUMULATION_STEPS == 0)
        if (idx + 1) % config.TRAIN.ACCUMULATION_STEPS == 0:
            seen["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
            optimizer.zero_grad()
        loss_scale_value = loss_scaler.state_dict()["scale"]
        torch.cuda.synchronize()
        seen.setdefault("loss", []).append(loss.item())
        assert probs.shape == outputs.shape and outputs.requires_grad
    # -----------------------------------------------------------------------------------------------
    assert grad_norm is not None and torch.isfinite(grad_norm) and loss_scale_value == 65536.0
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # the dead h_func branch has no gradient (find_unused_parameters=True in the reference, main_bigvul.py:162-164)
    dead = {n.split(".")[0] for n, p in model.named_parameters() if p.grad is None and n not in seen["grads"]}
    assert dead == {"fconly", "ln_text", "hbn", "hln", "hfc"}

    # fast path on a fresh copy of the same model: sum over the two micro-batches of grad(CE / 2), same BN statistics
    ref_model = cases.make_fusion().to(DEV)
    ref_model.load_state_dict(sd0)
    tr = train.FusionTrainer(ref_model, dropout=0.0, world_size=1)
    acc = {n: torch.zeros_like(v) for n, v in tr.named_grads().items()}
    losses = []
    for (g, img, txt, tgt) in data_loader:
        loss, _ = tr.forward_backward(g.to(DEV), img.to(DEV), txt.to(DEV), tgt.to(DEV))
        losses.append(float(loss) / 2)
        for n, v in tr.named_grads().items():
            acc[n] += v / 2
    assert max(abs(a - b) / abs(b) for a, b in zip(seen["loss"], losses)) < 1e-5
    flat_a = torch.cat([seen["grads"][n].reshape(-1) for n in tr.names])
    flat_b = torch.cat([acc[n].reshape(-1) for n in tr.names])
    # clip_grad_norm_(5.0) was applied in place to the accumulated gradients before they were recorded
    nrm = float(flat_b.double().norm())
    flat_b = flat_b * min(1.0, 5.0 / (nrm + 1e-6))
    rel = float((flat_a - flat_b).norm() / flat_b.norm())
    assert rel < 2e-2, rel            # same kernels; float atomics in the GAT / bias reductions and the 65536x loss scale through bf16 cotangents
    assert abs(float(grad_norm) - nrm) / nrm < 2e-2

    # AdamW took the step on every trained parameter, left the dead branch alone, and BatchNorm statistics moved
    moved = [n for n in tr.names if not torch.equal(after[n], sd0[n])]
    assert len(moved) == len(tr.names)
    assert all(torch.equal(after[k], sd0[k]) for k in after if k.split(".")[0] in dead)
    assert not torch.equal(after["swinbn.running_mean"], sd0["swinbn.running_mean"])
    assert int(after["swinbn.num_batches_tracked"]) == int(sd0["swinbn.num_batches_tracked"]) + 2
    # first AdamW step from zero moments: |delta| = lr (1 + wd |w|...) ~ lr for every element with a non-zero gradient
    w0, w1 = sd0["hidden.3.weight"], after["hidden.3.weight"]
    assert float((w1 - w0).abs().max()) < 1.2e-3

    # the next train-mode forward and the eval-mode forward both see the updated weights (bf16 operands refreshed)
    g, img, txt, _ = data_loader[0]
    model.eval()
    out_new = model(g.to(DEV), img.to(DEV), txt.to(DEV))
    ref_model.eval()
    out_old = ref_model(g.to(DEV), img.to(DEV), txt.to(DEV))
    assert not out_new.requires_grad and not torch.equal(out_new, out_old)
    model.train()
    out_t1 = model(g.to(DEV), img.to(DEV), txt.to(DEV))
    ref_model.load_state_dict(after)
    tr2 = train.FusionTrainer(ref_model, dropout=0.0, world_size=1)
    out_t2, _ = tr2.forward_train(g.to(DEV), img.to(DEV), txt.to(DEV))
    assert torch.allclose(out_t1.detach(), out_t2, rtol=1e-4, atol=1e-5)


def test_input_embedding_gradients_match_the_autograd_oracle():
    """d loss / d img_embedding and d func_text_embedding (what trainable encoders receive) against fp32 autograd
    through the oracle's train-mode restatement."""
    from oracle import fusion_train as otrain
    model = cases.make_fusion().to(DEV)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    model.train()
    model.train_engine().p_drop = 0.0
    (g, img, txt, tgt), = _batches(1)
    img_d, txt_d = img.to(DEV).requires_grad_(True), txt.to(DEV).requires_grad_(True)
    out = model(g.to(DEV), img_d, txt_d)
    loss = torch.nn.functional.cross_entropy(out, tgt.to(DEV))
    loss.backward()
    assert img_d.grad is not None and txt_d.grad is not None
    # oracle: fp32 autograd with the embeddings as leaves
    img_o, txt_o = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    otrain.loss_and_grads(sd, cases.to_host_batch(g), img_o, txt_o, tgt, emulate_bf16=True)
    for a, b in ((img_d.grad, img_o.grad), (txt_d.grad, txt_o.grad)):
        rel = float((a.cpu() - b).norm() / b.norm())
        assert rel < 5e-2, rel


def test_train_mode_forward_without_grad_and_eval_mode_has_no_graph():
    model = cases.make_fusion().to(DEV)
    (g, img, txt, _), = _batches(1)
    model.eval()
    assert not model(g.to(DEV), img.to(DEV), txt.to(DEV)).requires_grad
    model.train()
    with torch.no_grad():
        assert not model(g.to(DEV), img.to(DEV), txt.to(DEV)).requires_grad
    assert model(g.to(DEV), img.to(DEV), txt.to(DEV)).requires_grad
    with pytest.raises(RuntimeError):
        model(g.to(DEV), img, txt.to(DEV))                 # CPU tensor: no fallback
