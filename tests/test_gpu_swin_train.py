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
