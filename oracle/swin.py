"""Oracle (test infrastructure): functional fp32 restatement of SwinV2.

Follows /root/reference/mvuld/models/swin_transformer_v2.py.  State-dict keys
are the reference's (``layers.{i}.blocks.{j}.attn.qkv.weight`` ...).  Buffers
(``relative_coords_table``, ``relative_position_index``, ``attn_mask``) are
re-derived from geometry here, never read from the state dict, exactly as the
reference's pretrained loader does (mvuld/utils_multi.py:40-53).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class SwinGeometry:
    img_size: int = 448
    patch_size: int = 4
    in_chans: int = 3
    embed_dim: int = 128
    depths: Tuple[int, ...] = (2, 2, 18, 2)
    num_heads: Tuple[int, ...] = (4, 8, 16, 32)
    window_size: int = 28
    mlp_ratio: float = 4.0
    pretrained_window_sizes: Tuple[int, ...] = (12, 12, 12, 6)
    num_classes: int = 2

    def stage(self, i: int):
        """(H, W, C, heads, window, shift_for_odd_block) of stage i.

        swin_transformer_v2.py:228-231: when the resolution is <= window the
        window is clamped to it and the shift is dropped.
        """
        res = self.img_size // self.patch_size // (2 ** i)
        c = self.embed_dim * (2 ** i)
        ws = self.window_size
        shift = ws // 2
        if res <= ws:
            ws, shift = res, 0
        return res, res, c, self.num_heads[i], ws, shift


def relative_coords_table(ws: int, pretrained_ws: int) -> torch.Tensor:
    """swin_transformer_v2.py:98-111 -> [ (2ws-1)^2, 2 ] fp32 log-spaced coordinates."""
    r = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    hh = r[:, None].expand(2 * ws - 1, 2 * ws - 1)
    ww = r[None, :].expand(2 * ws - 1, 2 * ws - 1)
    t = torch.stack([hh, ww], dim=-1).contiguous()
    denom = (pretrained_ws - 1) if pretrained_ws > 0 else (ws - 1)
    t = t / denom
    t = t * 8
    t = torch.sign(t) * torch.log2(torch.abs(t) + 1.0) / math.log2(8)
    return t.reshape(-1, 2)


def relative_position_index(ws: int) -> torch.Tensor:
    """swin_transformer_v2.py:116-125 -> int64 [ws*ws, ws*ws].

    idx(i, j) = (h_i - h_j + ws - 1) * (2 ws - 1) + (w_i - w_j + ws - 1)
    """
    n = torch.arange(ws * ws)
    h, w = n // ws, n % ws
    dh = h[:, None] - h[None, :] + ws - 1
    dw = w[:, None] - w[None, :] + ws - 1
    return dh * (2 * ws - 1) + dw


def shifted_region_ids(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """swin_transformer_v2.py:247-262: 9-region id of every token of the
    *shifted* image, window-partitioned -> int64 [nW, ws*ws]."""
    def band(n):
        idx = torch.arange(n)
        return (idx >= n - ws).long() + (idx >= n - shift).long()
    reg = band(H)[:, None] * 3 + band(W)[None, :]
    reg = reg.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    return reg


def shifted_window_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """swin_transformer_v2.py:263-264 -> fp32 [nW, N, N] in {0, -100}."""
    reg = shifted_region_ids(H, W, ws, shift)
    diff = reg[:, None, :] != reg[:, :, None]
    return torch.where(diff, torch.tensor(-100.0), torch.tensor(0.0))


def cpb_bias_table(sd: Dict[str, torch.Tensor], prefix: str, ws: int, pretrained_ws: int) -> torch.Tensor:
    """swin_transformer_v2.py:159,163 -> 16*sigmoid(cpb_mlp(table)) as [(2ws-1)^2, nH]."""
    t = relative_coords_table(ws, pretrained_ws)
    hdn = F.relu(F.linear(t, sd[prefix + "cpb_mlp.0.weight"], sd[prefix + "cpb_mlp.0.bias"]))
    tab = F.linear(hdn, sd[prefix + "cpb_mlp.2.weight"])
    return 16 * torch.sigmoid(tab)


def window_attention(sd, prefix, x, ws, num_heads, pretrained_ws, mask):
    """swin_transformer_v2.py:140-179. x: [B_, N, C]."""
    B_, N, C = x.shape
    qkv_bias = torch.cat([sd[prefix + "q_bias"], torch.zeros_like(sd[prefix + "v_bias"]), sd[prefix + "v_bias"]])
    qkv = F.linear(x, sd[prefix + "qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B_, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    scale = torch.clamp(sd[prefix + "logit_scale"], max=math.log(1.0 / 0.01)).exp()
    attn = attn * scale
    bias = cpb_bias_table(sd, prefix, ws, pretrained_ws)            # [T, nH]
    idx = relative_position_index(ws).view(-1)
    bias = bias[idx].view(N, N, num_heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def _partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _reverse(win, ws, H, W):
    B = win.shape[0] // ((H // ws) * (W // ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def swin_block(sd, prefix, x, H, W, C, heads, ws, shift, pretrained_ws, mlp_ratio):
    """swin_transformer_v2.py:270-306 (eval mode: DropPath is identity)."""
    B, L, _ = x.shape
    shortcut = x
    x = x.view(B, H, W, C)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        mask = shifted_window_mask(H, W, ws, shift)
    else:
        mask = None
    xw = _partition(x, ws)
    aw = window_attention(sd, prefix + "attn.", xw, ws, heads, pretrained_ws, mask)
    x = _reverse(aw, ws, H, W)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x.reshape(B, L, C)
    x = shortcut + F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    y = F.linear(x, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])
    x = x + F.layer_norm(y, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])
    return x


def patch_merging(sd, prefix, x, H, W, C):
    """swin_transformer_v2.py:343-364: gather order (0,0),(1,0),(0,1),(1,1)."""
    B = x.shape[0]
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.view(B, -1, 4 * C)
    x = F.linear(x, sd[prefix + "reduction.weight"])
    return F.layer_norm(x, (2 * C,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


def patch_embed(sd, x, geo: SwinGeometry):
    """swin_transformer_v2.py:485-493."""
    y = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=geo.patch_size)
    y = y.flatten(2).transpose(1, 2)
    return F.layer_norm(y, (geo.embed_dim,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"])


def _forward_features(sd, geo: SwinGeometry, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """Differentiable body of forward_features (sd already fp32)."""
    x = patch_embed(sd, x.float(), geo)
    if taps is not None:
        taps["patch_embed"] = x.clone()
    for i, depth in enumerate(geo.depths):
        H, W, C, heads, ws, shift = geo.stage(i)
        for j in range(depth):
            x = swin_block(sd, f"layers.{i}.blocks.{j}.", x, H, W, C, heads, ws,
                           shift if (j % 2 == 1) else 0, geo.pretrained_window_sizes[i], geo.mlp_ratio)
            if taps is not None:
                taps[f"layers.{i}.blocks.{j}"] = x.clone()
        if i < len(geo.depths) - 1:
            x = patch_merging(sd, f"layers.{i}.downsample.", x, H, W, C)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), sd["norm.weight"], sd["norm.bias"])
    return x.mean(dim=1)


@torch.no_grad()
def forward_features(sd: Dict[str, torch.Tensor], geo: SwinGeometry, x: torch.Tensor,
                     taps: Optional[dict] = None) -> torch.Tensor:
    """swin_transformer_v2.py:623-635 -> [B, num_features]."""
    sd = {k: v.float() for k, v in sd.items() if torch.is_floating_point(v)}
    return _forward_features(sd, geo, x, taps)


def features_and_grads(sd: Dict[str, torch.Tensor], geo: SwinGeometry, x: torch.Tensor, cotangent: torch.Tensor):
    """Backward oracle of the image branch (the "whole path trainable" reading of configs[4]):
    fp32 autograd through the restated forward (DropPath / dropout rates 0, as in the pinned test cases) ->
    (features [B, num_features], {parameter name: d <features, cotangent> / d parameter}, d / d image).
    Pinned against autograd through the reference module (tests/golden/swin_train.pt)."""
    leaf = {k: v.detach().float().clone().requires_grad_(True) for k, v in sd.items() if torch.is_floating_point(v)}
    xin = x.detach().float().clone().requires_grad_(True)
    feats = _forward_features(leaf, geo, xin)
    (feats * cotangent.float()).sum().backward()
    grads = {k: v.grad.detach() for k, v in leaf.items() if v.grad is not None}
    return feats.detach(), grads, xin.grad.detach()


@torch.no_grad()
def forward(sd, geo: SwinGeometry, x):
    """swin_transformer_v2.py:640-643."""
    f = forward_features(sd, geo, x)
    return F.linear(f, sd["head.weight"].float(), sd["head.bias"].float())
