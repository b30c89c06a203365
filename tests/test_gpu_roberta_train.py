"""GPU: RoBERTa (UniXcoder) encoder backward -- the text half of BASELINE.json configs[4] in its primary reading --
against the backward oracle (fp32 autograd through the restated encoder, pinned to autograd through the installed HF
RobertaModel by tests/golden/roberta_train.pt)."""
import math
import os

import pytest
import torch

from mvuld_b200 import _lib, roberta_train, synth
from oracle import roberta as oroberta
from tests import cases
from tests.conftest import GOLDEN, record_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_seq_attention_backward_matches_autograd():
    """mvuld_seq_attention_train + _bwd_prep + _bwd + mvuld_seq_qkv_bwd against fp32 autograd of
    softmax(q k^T / sqrt(hd) + key mask) v on the same q, k, v (ragged lengths, one sequence of full length)."""
    g = torch.Generator().manual_seed(4711)
    B, L, nH, hd = 3, 512, 2, 64
    H = nH * hd
    lens = torch.tensor([512, 300, 77], dtype=torch.int32)
    qr = torch.randn(B, nH, L, hd, generator=g)                       # x Wq + b
    k = torch.randn(B, nH, L, hd, generator=g).to(torch.bfloat16)
    v = torch.randn(B, nH, L, hd, generator=g).to(torch.bfloat16)
    qs = (qr * (LOG2E / math.sqrt(hd))).to(torch.bfloat16)            # as mvuld_heads_qkv stores it
    valid = (torch.arange(L)[None, :] < lens[:, None])
    dO_tok = (torch.randn(B * L, H, generator=g) * 0.5 * valid.reshape(-1, 1)).to(torch.bfloat16)   # padded rows: no gradient
    # reference
    qf = (qs.float() * (math.sqrt(hd) / LOG2E)).requires_grad_(True)  # the un-scaled projection, from the stored values
    kf, vf = k.float().requires_grad_(True), v.float().requires_grad_(True)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(hd) + (~valid)[:, None, None, :] * -10000.0
    out = s.softmax(-1) @ vf                                          # [B, nH, L, hd]
    dOh_ref = dO_tok.float().view(B, L, nH, hd).permute(0, 2, 1, 3)
    (out * dOh_ref).sum().backward()
    d = lambda t: t.to(DEV).contiguous()
    qd, kd, vd, ld_ = d(qs), d(k), d(v), d(lens)
    att = torch.zeros(B * L, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B * nH, L, device=DEV)
    _lib.call("mvuld_seq_attention_train", qd, kd, vd, ld_, att, lse, B, L, nH, hd)
    torch.cuda.synchronize()
    out_tok = out.detach().permute(0, 2, 1, 3).reshape(B * L, H)
    assert rel_err(att[valid.reshape(-1)], out_tok[valid.reshape(-1)]) < 1e-2
    dOh, ldp = torch.empty(B * nH, L, hd, device=DEV, dtype=torch.bfloat16), torch.empty(B * nH, L, 2, device=DEV)
    _lib.call("mvuld_seq_attention_bwd_prep", d(dO_tok), att, lse, dOh, ldp, B, L, nH)
    dq, dk, dv = (torch.zeros(B * nH, L, hd, device=DEV) for _ in range(3))
    _lib.call("mvuld_seq_attention_bwd", qd, kd, vd, dOh, ldp, ld_, dq, dk, dv, B, L, nH, hd)
    dqkv = torch.empty(B * L, 3 * H, device=DEV, dtype=torch.bfloat16)
    _lib.call("mvuld_seq_qkv_bwd", dq, dk, dv, dqkv, B, L, nH, hd)
    torch.cuda.synchronize()
    tok = lambda t: t.permute(0, 2, 1, 3).reshape(B * L, H)
    ref = torch.cat([tok(qf.grad), tok(kf.grad), tok(vf.grad)], 1)
    m = valid.reshape(-1)
    errs = {n: rel_err(dqkv[m][:, i * H:(i + 1) * H], ref[m][:, i * H:(i + 1) * H]) for i, n in enumerate("qkv")}
    for n, e in errs.items():
        record_parity(f"seq_attention_bwd d{n} rel-L2", e, 1e-2)
    assert max(errs.values()) < 1e-2, errs
    assert float(dqkv[~m].float().abs().max()) == 0.0                # padded rows get exactly nothing


def _case(full: bool, B: int, seed: int, cot=None, ids=None):
    model = cases.make_roberta(full)
    enc = model.encoder
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = enc.config
    if ids is None:
        ids = synth.token_ids(B, 512, cfg.vocab_size, seed=seed)
    tr = roberta_train.RobertaTrainer(enc.to(DEV), world_size=1)
    sent, ctx = tr.forward_train(ids.to(DEV))
    if cot is None:
        cot = torch.randn(sent.shape, generator=torch.Generator().manual_seed(seed + 1))
    tr.flat_g.zero_()
    tr.backward_train(ctx, cot.to(DEV))
    torch.cuda.synchronize()
    sent_ref, gref = oroberta.sentence_and_grads(sd, cases.roberta_geometry(cfg), ids, cot)
    grads = {"encoder." + k: v.detach().cpu().clone() for k, v in tr.named_grads().items()}
    return tr, sent.cpu(), sent_ref, grads, gref


def _check(tag, grads, gref, worst_tol=5e-2, flat_tol=1e-2, floor=1e-3):
    assert set(grads) == set(gref), (set(grads) ^ set(gref))
    names = list(gref)
    a = torch.cat([grads[n].reshape(-1).float() for n in names])
    b = torch.cat([gref[n].reshape(-1).float() for n in names])
    flat, flat_norm = rel_err(a, b), float(b.norm())
    errs = {n: float((grads[n].reshape(gref[n].shape).float() - gref[n].float()).norm()) /
            max(float(gref[n].norm()), floor * flat_norm) for n in names}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    dense = [n for n in names if n.endswith("weight") and gref[n].dim() == 2 and "embeddings" not in n]
    wd = max(((rel_err(grads[n], gref[n]), n) for n in dense))
    record_parity(f"roberta_train[{tag}] flat gradient rel-L2", flat, flat_tol)
    record_parity(f"roberta_train[{tag}] worst tensor, error / max(|ref|, 1e-3 |flat|) ({worst[0][0]})", worst[0][1], worst_tol)
    record_parity(f"roberta_train[{tag}] worst weight matrix rel-L2 ({wd[1]})", wd[0], worst_tol)
    assert flat < flat_tol, (flat, worst)
    assert worst[0][1] < worst_tol, worst
    assert wd[0] < worst_tol, wd


def test_roberta_encoder_backward_matches_hf_autograd_golden():
    """2-layer, 2-head config, the golden file's ids and cotangent: every gradient vs the oracle, and the oracle's pin
    (norms of HF RobertaModel's autograd gradients) re-checked, then the CUDA gradient norms against HF's."""
    gold = torch.load(os.path.join(GOLDEN, "roberta_train.pt"), weights_only=False)
    cfg = cases.roberta_small_config()
    ids = synth.token_ids(cases.ROBERTA_BATCH, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED + 31)
    tr, sent, sent_ref, grads, gref = _case(False, cases.ROBERTA_BATCH, cases.SEED + 31, gold["cotangent"], ids)
    assert rel_err(sent_ref, gold["sent"]) < 1e-5
    assert rel_err(sent, sent_ref) < 1e-2
    flat_norm = math.sqrt(sum(g["norm"] ** 2 for g in gold["grads"].values()))
    for k, g in gold["grads"].items():
        assert abs(float(gref[k].double().norm()) - g["norm"]) <= 1e-3 * max(g["norm"], 1e-6), k
        assert abs(float(grads[k].double().norm()) - g["norm"]) <= 5e-2 * max(g["norm"], 1e-3 * flat_norm), k
    _check("2-layer", grads, gref)


def test_roberta_base_backward_full_size_matches_oracle():
    """RoBERTa-base (12 layers, 768 hidden, 12 heads), two sequences of 512 token slots."""
    tr, sent, sent_ref, grads, gref = _case(True, 2, cases.SEED + 71)
    assert rel_err(sent, sent_ref) < 1e-2
    _check("roberta-base", grads, gref)
