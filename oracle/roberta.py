"""Oracle (test infrastructure): functional fp32 restatement of the UniXcoder encode path.

Reference call sites: /root/reference/mvuld/models/unixcoder.py:33-38 (get_xcode_vec),
:91-95 (get_repr).  The encoder arithmetic is HF ``transformers==4.18.0``
``RobertaModel`` (third-party, pinned at environment.yml:289, source absent from
/root/reference); this file restates its published algorithm:

* embeddings = word[ids] + position[cumsum(ids!=pad)*(ids!=pad) + pad] + token_type[0], LayerNorm(eps)
* 12 post-LN layers: softmax(QK^T/sqrt(hd) + (1 - mask3d) * -10000) V, dense+residual+LN,
  dense(3072)+GELU(erf)+dense+residual+LN
* the reference passes the 3-D mask ``mask[:,None,:] * mask[:,:,None]`` (unixcoder.py:36) so pad-query
  rows see every key masked by the same constant (= plain softmax over all keys); those rows are
  excluded by the masked mean (unixcoder.py:37) and never reach a valid row.

State-dict keys are HF's, optionally prefixed (``encoder.`` inside ``MyUniXcoder``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass
class RobertaGeometry:
    vocab_size: int = 51416
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 1026
    type_vocab_size: int = 10
    pad_token_id: int = 1
    layer_norm_eps: float = 1e-5
    max_source_length: int = 512


def position_ids(ids: torch.Tensor, pad: int) -> torch.Tensor:
    """HF create_position_ids_from_input_ids: int64, pad rows -> pad."""
    m = ids.ne(pad).long()
    return torch.cumsum(m, dim=1) * m + pad


@torch.no_grad()
def encode(sd: Dict[str, torch.Tensor], geo: RobertaGeometry, ids: torch.Tensor, prefix: str = "",
           key_mask_only: bool = False):
    """-> (token_embeddings [B,L,H], sentence_embeddings [B,H]); unixcoder.py:33-38."""
    return _encode(sd, geo, ids, prefix, key_mask_only)


def sentence_and_grads(sd: Dict[str, torch.Tensor], geo: RobertaGeometry, ids: torch.Tensor, cotangent: torch.Tensor,
                       prefix: str = "encoder."):
    """Backward oracle of the text branch (the "whole path trainable" reading of configs[4]): fp32
    autograd through the restated encoder + masked mean -> (sentence vectors [B, H], {parameter name: gradient of
    <vectors, cotangent>}).  Dropout 0.  Pinned against autograd through the installed HF RobertaModel
    (tests/golden/roberta_train.pt); the pooler takes no part in get_repr and gets no gradient."""
    leaf = {k: v.detach().float().clone().requires_grad_(True) for k, v in sd.items()
            if k.startswith(prefix) and torch.is_floating_point(v)}
    sent = _encode(leaf, geo, ids, prefix, False)[1]
    (sent * cotangent.float()).sum().backward()
    return sent.detach(), {k: v.grad.detach() for k, v in leaf.items() if v.grad is not None}


def _encode(sd, geo, ids, prefix, key_mask_only):
    g = lambda k: sd[prefix + k].float()
    B, L = ids.shape
    Hd, nH = geo.hidden_size, geo.num_attention_heads
    hd = Hd // nH
    m = ids.ne(geo.pad_token_id)
    x = g("embeddings.word_embeddings.weight")[ids] \
        + g("embeddings.position_embeddings.weight")[position_ids(ids, geo.pad_token_id)] \
        + g("embeddings.token_type_embeddings.weight")[0]
    x = F.layer_norm(x, (Hd,), g("embeddings.LayerNorm.weight"), g("embeddings.LayerNorm.bias"), geo.layer_norm_eps)
    mf = m.float()
    if key_mask_only:
        ext = (1.0 - mf)[:, None, None, :] * -10000.0
    else:
        ext = (1.0 - mf[:, None, :] * mf[:, :, None])[:, None] * -10000.0
    for i in range(geo.num_hidden_layers):
        p = f"encoder.layer.{i}."
        q = F.linear(x, g(p + "attention.self.query.weight"), g(p + "attention.self.query.bias"))
        k = F.linear(x, g(p + "attention.self.key.weight"), g(p + "attention.self.key.bias"))
        v = F.linear(x, g(p + "attention.self.value.weight"), g(p + "attention.self.value.bias"))
        sh = lambda t: t.view(B, L, nH, hd).permute(0, 2, 1, 3)
        s = sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(hd) + ext
        ctx = (s.softmax(-1) @ sh(v)).permute(0, 2, 1, 3).reshape(B, L, Hd)
        a = F.linear(ctx, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias"))
        x = F.layer_norm(a + x, (Hd,), g(p + "attention.output.LayerNorm.weight"),
                         g(p + "attention.output.LayerNorm.bias"), geo.layer_norm_eps)
        h = F.gelu(F.linear(x, g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")))
        o = F.linear(h, g(p + "output.dense.weight"), g(p + "output.dense.bias"))
        x = F.layer_norm(o + x, (Hd,), g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"),
                         geo.layer_norm_eps)
    sent = (x * mf[..., None]).sum(1) / mf.sum(-1)[..., None]
    return x, sent


@torch.no_grad()
def get_repr(sd, geo: RobertaGeometry, input_ids: torch.Tensor, prefix: str = "encoder."):
    """unixcoder.py:91-95 -> vec [B, 768]."""
    ids = input_ids.view(-1, geo.max_source_length)
    return encode(sd, geo, ids, prefix)[1]
