"""GPU: SwinV2 encoder backward (BASELINE.json configs[4], primary reading: the image encoder trains as in
/root/reference/mvuld/main.py:251-300) -- kernel-level parity against fp32 autograd, then the whole encoder against the
backward oracle pinned to the reference module (tests/golden/swin_train.pt)."""
import math

import pytest
import torch

from mvuld_b200 import _lib
from oracle import swin as oswin
from tests.conftest import record_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _window_major(x_tok, B, H, W, ws, shift, nH):
    """token-major [B*H*W, nH*32] -> window-major head-major [B*nW, nH, ws*ws, 32] (cyclic shift applied)."""
    x = x_tok.view(B, H, W, nH * 32)
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    w = oswin._partition(x, ws)                                   # [B*nW, N, C]
    return w.view(w.shape[0], ws * ws, nH, 32).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("Hres,ws,shift,nH,B,fixed", [(28, 28, 0, 4, 2, 1), (56, 28, 14, 2, 1, 1), (56, 28, 14, 2, 1, 0),
                                                     (28, 14, 7, 4, 1, 0), (14, 14, 0, 8, 2, 0), (14, 7, 3, 4, 2, 0)])
def test_window_attention_backward_matches_autograd(Hres, ws, shift, nH, B, fixed):
    """mvuld_swin_window_attention_train + mvuld_swin_attention_bwd_prep + mvuld_swin_attention_bwd against fp32
    autograd of softmax(q^ k^T + bias + mask) v on the SAME q^, k^, v (swin_transformer_v2.py:155-176)."""
    g = gen(900 + Hres + ws + shift + nH)
    H = W = Hres
    C = nH * 32
    N = ws * ws
    M = B * H * W
    nW = (H // ws) * (W // ws)
    n_bh = B * nW * nH
    side = 2 * ws - 1
    scale = 8.0 + 4.0 * torch.rand(nH, generator=g)                               # exp(logit_scale): constant-reference range
    q = torch.nn.functional.normalize(torch.randn(n_bh, N, 32, generator=g), dim=-1)
    q = (q.view(B * nW, nH, N, 32) * (scale * LOG2E).view(1, nH, 1, 1)).reshape(n_bh, N, 32).to(torch.float16)
    k = torch.nn.functional.normalize(torch.randn(n_bh, N, 32, generator=g), dim=-1).to(torch.float16)
    v = torch.randn(n_bh, N, 32, generator=g).to(torch.bfloat16)
    tab_nat = 16 * torch.sigmoid(torch.randn(nH, side * side, generator=g))       # [nH, T] natural units
    tab_rev = (tab_nat.view(nH, side, side).flip(2) * LOG2E).reshape(nH, side * side).contiguous()
    tab_max = tab_rev.max(1).values.contiguous()
    dO_tok = (torch.randn(M, C, generator=g) * 0.5).to(torch.bfloat16)

    # ---- fp32 autograd reference on the same operands ----
    qf = q.float().view(B * nW, nH, N, 32).requires_grad_(True)
    kf = k.float().view(B * nW, nH, N, 32).requires_grad_(True)
    vf = v.float().view(B * nW, nH, N, 32).requires_grad_(True)
    idx = oswin.relative_position_index(ws).view(-1)
    bias = tab_nat[:, idx].view(nH, N, N).clone().requires_grad_(True)           # natural units, a leaf per (h, q, k)
    s_nat = (qf @ kf.transpose(-2, -1)) / LOG2E + bias.unsqueeze(0)
    if shift:
        mask = oswin.shifted_window_mask(H, W, ws, shift)                         # [nW, N, N]
        s_nat = (s_nat.view(B, nW, nH, N, N) + mask[None, :, None]).view(B * nW, nH, N, N)
    pr = s_nat.softmax(-1)
    out = pr @ vf
    dOw_ref = _window_major(dO_tok.float(), B, H, W, ws, shift, nH)
    (out * dOw_ref).sum().backward()
    G = None                                                                      # dL / d natural logits, per window
    # kernel conventions: dq = G k^, dk = G^T q^ (q^ as stored): autograd's d/dq^ = dq / log2e etc.
    dq_ref = qf.grad * LOG2E
    dk_ref = kf.grad * LOG2E
    dv_ref = vf.grad
    dbias_ref = bias.grad                                                         # [nH, N, N] = sum over windows of G

    # ---- CUDA path ----
    d = lambda t: t.to(DEV).contiguous()
    qd, kd, vd = d(q), d(k), d(v)
    out_tok = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(n_bh, N, device=DEV)
    qn = d(scale * LOG2E)
    _lib.call("mvuld_swin_window_attention_train", qd, kd, vd, d(tab_rev), d(tab_max), qn, out_tok, lse, fixed, B, H, W,
              C, nH, ws, shift)
    torch.cuda.synchronize()
    out_ref_tok = oswin._reverse(out.detach().permute(0, 2, 1, 3).reshape(B * nW, N, C), ws, H, W)
    if shift:
        out_ref_tok = torch.roll(out_ref_tok, (shift, shift), (1, 2))
    assert rel_err(out_tok, out_ref_tok.reshape(M, C)) < 1e-2
    lse_ref = torch.logsumexp(s_nat.detach(), -1) * LOG2E                         # log2 units
    assert float((lse.cpu().view_as(lse_ref) - lse_ref).abs().max()) < 2e-2

    dOw = torch.empty(n_bh, N, 32, device=DEV, dtype=torch.bfloat16)
    ld = torch.empty(n_bh, N, 2, device=DEV)
    qb = torch.empty(n_bh, N, 32, device=DEV, dtype=torch.bfloat16)
    kb = torch.empty(n_bh, N, 32, device=DEV, dtype=torch.bfloat16)
    _lib.call("mvuld_swin_attention_bwd_prep", d(dO_tok), out_tok, lse, qd, kd, dOw, ld, qb, kb, B, H, W, C, nH, ws, shift)
    torch.cuda.synchronize()
    assert torch.equal(dOw.cpu().float().view_as(dOw_ref), dOw_ref)              # a pure gather
    npad = (N + 7) // 8 * 8
    dq = torch.zeros(n_bh, N, 32, device=DEV)
    dk = torch.zeros(n_bh, N, 32, device=DEV)
    dv = torch.zeros(n_bh, N, 32, device=DEV)
    gt = torch.zeros(n_bh, N, npad, device=DEV, dtype=torch.bfloat16)
    _lib.call("mvuld_swin_attention_bwd", qd, qb, kd, kb, vd, dOw, ld, d(tab_rev), dq, dk, dv, gt, npad, B, H, W, nH, ws,
              shift)
    torch.cuda.synchronize()
    errs = dict(dv=rel_err(dv.view_as(dv_ref), dv_ref), dq=rel_err(dq.view_as(dq_ref), dq_ref),
                dk=rel_err(dk.view_as(dk_ref), dk_ref))
    gsum = gt.float().view(B * nW, nH, N, npad)[..., :N].sum(0).transpose(-1, -2).cpu()     # [nH, q, k]
    errs["dbias"] = rel_err(gsum, dbias_ref)
    for name, e in errs.items():
        record_parity(f"swin_attention_bwd[H{Hres} ws{ws} shift{shift} fixed{fixed}] {name} rel-L2", e, 1e-2)
    assert max(errs.values()) < 1e-2, errs


# --------------------------------------------------------------------------------------------------------
# the whole encoder: forward_train + backward_train against the backward oracle (fp32 autograd through the restated
# forward, pinned to autograd through the UNMODIFIED reference module by tests/golden/swin_train.pt)
# --------------------------------------------------------------------------------------------------------
def _encoder_case(name, B, seed_x, cot=None):
    import os
    from mvuld_b200 import swin_train, synth
    from tests import cases
    model = cases.make_swin(name)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    geo = cases.swin_geometry(name)
    x = synth.images(B, cases.SWIN_CASES[name]["img_size"], seed=seed_x)
    tr = swin_train.SwinTrainer(model.to(DEV), world_size=1)
    feat, ctx = tr.forward_train(x.to(DEV))
    if cot is None:
        cot = torch.randn(feat.shape, generator=gen(seed_x + 1))
    tr.flat_g.zero_()
    tr.backward_train(ctx, cot.to(DEV))
    torch.cuda.synchronize()
    feats_ref, gref, _ = oswin.features_and_grads(sd, geo, x, cot)
    grads = {k: v.detach().cpu().clone() for k, v in tr.named_grads().items()}
    return tr, feat.cpu(), feats_ref, grads, gref


def _check_grads(tag, grads, gref, worst_tol, flat_tol, floor=1e-3):
    """Flat relative L2 error over all parameters, and per tensor |g - ref| <= worst_tol * max(|ref|, floor * |ref_flat|).
    The floor matters for a handful of tiny gradients that are sums with near-total cancellation (logit_scale, q_bias and
    the cpb_mlp of late stages, whose keys / queries are almost collinear on these synthetic weights: |ref| ~ 1e-1 next
    to a flat norm of 5e3): bf16 G / k^ operands leave an ABSOLUTE error of ~4e-2 there, i.e. 1e-5 of the gradient."""
    names = [n for n in gref if n in grads]
    assert set(n for n in grads if not n.startswith("head.")) == set(names), "every trained parameter needs a reference gradient"
    a = torch.cat([grads[n].reshape(-1).float() for n in names])
    b = torch.cat([gref[n].reshape(-1).float() for n in names])
    flat, flat_norm = rel_err(a, b), float(b.norm())
    errs = {}
    for n in names:
        ref = gref[n].float()
        errs[n] = float((grads[n].reshape(ref.shape).float() - ref).norm()) / max(float(ref.norm()), floor * flat_norm)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    dense = ("qkv.weight", "proj.weight", "fc1.weight", "fc2.weight", "reduction.weight")       # the GEMM weights
    big = {n: rel_err(grads[n].reshape(gref[n].shape), gref[n]) for n in names if n.endswith(dense)}
    worst_big = max(big.items(), key=lambda kv: kv[1])
    record_parity(f"swin_train[{tag}] flat gradient rel-L2", flat, flat_tol)
    record_parity(f"swin_train[{tag}] worst tensor, error / max(|ref|, 1e-3 |flat|) ({worst[0][0]})", worst[0][1], worst_tol)
    record_parity(f"swin_train[{tag}] worst weight matrix rel-L2 ({worst_big[0]})", worst_big[1], worst_tol)
    assert flat < flat_tol, (flat, worst)
    assert worst[0][1] < worst_tol, worst
    assert worst_big[1] < worst_tol, worst_big
    return errs


def test_swin_encoder_backward_matches_reference_autograd_golden():
    """small_ws7 (three stages, shifted windows, two patch mergings), batch 2, the golden file's image and cotangent:
    gradients of EVERY parameter vs the oracle, and the oracle's own pin (norm + samples of the reference module's
    autograd gradients) re-checked on the same run."""
    import os
    from tests import cases
    from tests.conftest import GOLDEN
    gold = torch.load(os.path.join(GOLDEN, "swin_train.pt"), weights_only=False)
    tr, feat, feats_ref, grads, gref = _encoder_case("small_ws7", 2, cases.SEED + 21, gold["cotangent"])
    assert rel_err(feats_ref, gold["features"]) < 1e-5                          # same inputs as the golden run
    assert rel_err(feat, feats_ref) < 1e-2
    for k, g in gold["grads"].items():                                          # the oracle against the reference module
        f = gref[k].reshape(-1)
        assert abs(float(f.double().norm()) - g["norm"]) <= 1e-3 * max(g["norm"], 1e-6), k
    errs = _check_grads("small_ws7", grads, gref, worst_tol=5e-2, flat_tol=1e-2)
    # against the reference module's own numbers: norms of the CUDA gradients
    flat_norm = math.sqrt(sum(g["norm"] ** 2 for g in gold["grads"].values()))
    for k, g in gold["grads"].items():                                         # (same floor as _check_grads)
        tol = 5e-2 * max(g["norm"], 1e-3 * flat_norm)
        assert abs(float(grads[k].double().norm()) - g["norm"]) <= tol, (k, float(grads[k].norm()), g["norm"])
        stride = max(1, g["numel"] // 16)
        mine = grads[k].reshape(-1)[::stride][:16].float()
        assert float((mine - g["strided"]).norm()) <= 5e-2 * max(float(g["strided"].norm()), 1e-3 * flat_norm / math.sqrt(max(1, g["numel"] / 16))) + 1e-6, k


def test_swin_encoder_backward_ws14_matches_oracle():
    """mid_ws14: 14x14 windows (two key / query tiles per window in the attention backward), shift 7, one merging."""
    from tests import cases
    tr, feat, feats_ref, grads, gref = _encoder_case("mid_ws14", 2, cases.SEED + 31)
    assert rel_err(feat, feats_ref) < 1e-2
    _check_grads("mid_ws14", grads, gref, worst_tol=5e-2, flat_tol=1e-2)


def test_swin_trainer_step_updates_and_is_reproducible():
    """step(): cross-entropy through the head, clip + AdamW on the flat buffer, bit-identical when repeated."""
    from mvuld_b200 import swin_train, synth
    from tests import cases
    outs = []
    for _ in range(2):
        model = cases.make_swin("small_ws7").to(DEV)
        tr = swin_train.SwinTrainer(model, lr=2e-5, world_size=1)
        x = synth.images(3, 112, seed=cases.SEED + 41).to(DEV)
        y = torch.tensor([0, 1, 1], device=DEV)
        p0 = tr.flat_p.clone()
        losses = []
        for _s in range(2):
            loss, logits = tr.step(x, y)
            losses.append(float(loss))
        assert all(math.isfinite(v) for v in losses) and float(tr.grad_norm()) > 0
        assert not torch.equal(p0, tr.flat_p)
        outs.append((losses, tr.flat_p.cpu().clone()))
        # the eval-mode module sees the updated weights
        f = model.eval()(x)
        assert torch.isfinite(f).all()
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    assert outs[0][0][-1] < outs[0][0][0], outs[0][0]                          # the repeated batch is fitted better


# --------------------------------------------------------------------------------------------------------
# BASELINE.json's full size: SwinV2-B 448 / window 28 (the four-group attention forward, the 7 x 7-unit backward)
# --------------------------------------------------------------------------------------------------------
def test_swin_encoder_backward_full_size_matches_oracle():
    """SwinV2-B 448 px / window 28, one image: every gradient against the fp32 oracle (about a minute of CPU autograd)."""
    from tests import cases
    tr, feat, feats_ref, grads, gref = _encoder_case("full", 1, cases.SEED + 51)
    assert rel_err(feat, feats_ref) < 1e-2
    assert any(b["fixed"] for b in tr.blocks)                                  # the constant-reference forward kernel ran
    _check_grads("full 448/w28", grads, gref, worst_tol=5e-2, flat_tol=1e-2)


def test_joint_trainer_chains_fusion_and_image_encoder_backward():
    """MVulDTrainer (configs[4], primary reading): the fusion backward's input gradients drive the SwinV2 and RoBERTa
    backward passes; one clip over the three parameter sets (232 M parameters).  Checked: every parameter set
    receives a gradient, the joint norm is the norm over the three flat buffers, repeated steps on one batch lower the
    loss (each encoder's own gradient parity is pinned in test_gpu_swin_train / test_gpu_roberta_train)."""
    import mvuld_b200 as mv
    from mvuld_b200 import synth
    from mvuld_b200.joint_train import MVulDTrainer
    from tests import cases
    torch.manual_seed(cases.SEED)
    model = mv.MVulD(mv.default_config()).eval()
    synth.randomize_for_parity(model, seed=777)
    model = model.to(DEV)
    B = 2
    tr = MVulDTrainer(model, lr=2e-5, dropout=0.0, world_size=1)
    assert tr.num_parameters > 225e6                                           # SwinV2-B 87 M + RoBERTa-base 125 M + fusion 19 M
    img = synth.images(B, 448, seed=cases.SEED + 61).to(DEV)
    ids = synth.token_ids(B, 512, seed=cases.SEED + 62).to(DEV)
    g = synth.cpg_batch(B, seed=cases.SEED + 63)
    g.ndata.pop("_FUNC_EMB", None)
    g = g.to(DEV)
    y = torch.tensor([0, 1], device=DEV)
    losses = []
    for _ in range(3):
        g._csr = g._ocsr = None
        loss, logits = tr.step(g, img, ids, y)
        losses.append(float(loss))
        assert torch.isfinite(logits).all()
    gs, gf = tr.swin.flat_g.double().pow(2).sum(), tr.fusion.flat_g.double().pow(2).sum()
    gt = tr.text.flat_g.double().pow(2).sum()
    assert float(gs) > 0 and float(gf) > 0 and float(gt) > 0
    assert abs(float(tr.grad_norm()) - float((gs + gf + gt).sqrt())) <= 1e-4 * float((gs + gf + gt).sqrt())
    assert losses[-1] < losses[0], losses
