"""UniXcoder text branch (RoBERTa-base encoder + masked mean), B200-native.

Interface mirror of /root/reference/mvuld/models/unixcoder.py:20-95 (``MyUniXcoder``): constructor
``(encoder, config, tokenizer, tokenize)``, ``get_xcode_vec`` / ``get_repr`` / ``forward``.  ``RobertaEncoder`` is a
parameter container with HF ``RobertaModel`` state-dict keys (``embeddings.word_embeddings.weight``,
``encoder.layer.{i}.attention.self.query.weight`` ...), so a fine-tuned UniXcoder checkpoint loads unchanged.  The
arithmetic follows HF transformers 4.18 RobertaModel as the reference calls it (3-D mask ``m[:,None,:]*m[:,:,None]``,
unixcoder.py:36): valid queries see valid keys only; pad-query rows are excluded by the masked mean, so the kernels
treat the mask as a per-sequence key length (pads must form a suffix, which is what ``tokenize`` produces, :150).

Per layer: heads_qkv GEMM (bias, 1/sqrt(hd)*log2e folded into q, head-major scatter) -> seq_attention (tcgen05) ->
dense GEMM -> LN(x + .) -> dense GEMM + GELU -> dense GEMM -> LN(x + .).  Eval mode only.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib

LOG2E = 1.4426950408889634


def roberta_base_config(**over):
    """microsoft/unixcoder-base-nine geometry (RoBERTa-base); SURVEY.md section 8(c)."""
    cfg = dict(vocab_size=51416, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
               intermediate_size=3072, max_position_embeddings=1026, type_vocab_size=10, pad_token_id=1,
               layer_norm_eps=1e-5, initializer_range=0.02)
    cfg.update(over)
    return SimpleNamespace(**cfg)


class _SelfAttention(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)


class _SelfOutput(nn.Module):
    def __init__(self, h_in, h, eps):
        super().__init__()
        self.dense = nn.Linear(h_in, h)
        self.LayerNorm = nn.LayerNorm(h, eps=eps)


class _Attention(nn.Module):
    def __init__(self, h, eps):
        super().__init__()
        setattr(self, "self", _SelfAttention(h))
        self.output = _SelfOutput(h, h, eps)


class _Intermediate(nn.Module):
    def __init__(self, h, inter):
        super().__init__()
        self.dense = nn.Linear(h, inter)


class _Layer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg.hidden_size, cfg.layer_norm_eps)
        self.intermediate = _Intermediate(cfg.hidden_size, cfg.intermediate_size)
        self.output = _SelfOutput(cfg.intermediate_size, cfg.hidden_size, cfg.layer_norm_eps)


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size,
                                                padding_idx=cfg.pad_token_id)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])


class _Pooler(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.dense = nn.Linear(h, h)


class RobertaEncoder(nn.Module):
    """HF ``RobertaModel`` stand-in: same parameter names, fused B200 forward."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config or roberta_base_config()
        cfg = self.config
        if cfg.hidden_size // cfg.num_attention_heads != 64:
            raise NotImplementedError("mvuld_b200 RobertaEncoder: head_dim 64 only")
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.pooler = _Pooler(cfg.hidden_size)       # present in HF checkpoints; unused by the MVulD path
        std = getattr(cfg, "initializer_range", 0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=std)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=std)
                if m.padding_idx is not None:
                    with torch.no_grad():
                        m.weight[m.padding_idx].zero_()
        self._plan = None

    def invalidate(self):
        self._plan = None

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def prepare(self):
        dev = self.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mvuld_b200 RobertaEncoder runs on CUDA only (no CPU fallback)")
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        e = self.embeddings
        plan = dict(dev=dev, word=f32(e.word_embeddings.weight), pos=f32(e.position_embeddings.weight),
                    type0=f32(e.token_type_embeddings.weight[0]), eg=f32(e.LayerNorm.weight),
                    eb=f32(e.LayerNorm.bias), layers=[], ws={})
        for lyr in self.encoder.layer:
            sa = getattr(lyr.attention, "self")
            plan["layers"].append(dict(
                wqkv=b16(torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0)),
                bqkv=f32(torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], 0)),
                wo=b16(lyr.attention.output.dense.weight), bo=f32(lyr.attention.output.dense.bias),
                g1=f32(lyr.attention.output.LayerNorm.weight), b1=f32(lyr.attention.output.LayerNorm.bias),
                wi=b16(lyr.intermediate.dense.weight), bi=f32(lyr.intermediate.dense.bias),
                wo2=b16(lyr.output.dense.weight), bo2=f32(lyr.output.dense.bias),
                g2=f32(lyr.output.LayerNorm.weight), b2=f32(lyr.output.LayerNorm.bias)))
        self._plan = plan
        return self

    def _workspace(self, B, L):
        p = self._plan
        key = (B, L)
        if key in p["ws"]:
            return p["ws"][key]
        dev, H, I = p["dev"], self.config.hidden_size, self.config.intermediate_size
        M = B * L
        e = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
        ws = dict(x32=e((M, H), torch.float32), xb=e((M, H), torch.bfloat16), q=e((M, H), torch.bfloat16),
                  k=e((M, H), torch.bfloat16), v=e((M, H), torch.bfloat16), ctx=e((M, H), torch.bfloat16),
                  y=e((M, H), torch.bfloat16), h=e((M, I), torch.bfloat16), pos=e((M,), torch.int32),
                  len=e((B,), torch.int32), ok=e((1,), torch.int32), sent=e((B, H), torch.float32))
        p["ws"] = {key: ws}
        return ws

    @torch.no_grad()
    def encode(self, source_ids: torch.Tensor, check_suffix_padding: bool = False):
        """-> (token_embeddings fp32 [B, L, H], sentence_embeddings fp32 [B, H]); unixcoder.py:33-38."""
        if self.training:
            raise RuntimeError("mvuld_b200 RobertaEncoder implements the eval-mode forward: call model.eval()")
        if not source_ids.is_cuda:
            raise RuntimeError("mvuld_b200 RobertaEncoder takes CUDA tensors (no CPU fallback)")
        if self._plan is None:
            self.prepare()
        p, cfg = self._plan, self.config
        ids = source_ids.to(torch.int64).contiguous()
        B, L = ids.shape
        if L > 512 or L % 8 != 0:
            raise ValueError("sequence length must be a multiple of 8 and <= 512")
        H, nH = cfg.hidden_size, cfg.num_attention_heads
        M = B * L
        w = self._workspace(B, L)
        w["ok"].fill_(1)
        _lib.call("mvuld_seq_positions", ids, B, L, int(cfg.pad_token_id), w["pos"], w["len"], w["ok"])
        _lib.call("mvuld_roberta_embed", ids, w["pos"], p["word"], p["pos"], p["type0"], p["eg"], p["eb"], w["x32"],
                  w["xb"], M, H, float(cfg.layer_norm_eps))
        qmul = LOG2E / math.sqrt(H // nH)
        eps = float(cfg.layer_norm_eps)
        for lp in p["layers"]:
            _lib.call("mvuld_heads_qkv", w["xb"], lp["wqkv"], lp["bqkv"], w["q"], w["k"], w["v"], B, L, H, nH, qmul)
            _lib.call("mvuld_seq_attention", w["q"], w["k"], w["v"], w["len"], w["ctx"], B, L, nH, H // nH)
            _lib.gemm(w["ctx"], lp["wo"], bias=lp["bo"], out_bf16=w["y"])
            _lib.call("mvuld_ln_rows", w["y"], w["x32"], lp["g1"], lp["b1"], w["x32"], w["xb"], M, H, eps, 2)
            _lib.gemm(w["xb"], lp["wi"], bias=lp["bi"], act=_lib.ACT_GELU, out_bf16=w["h"])
            _lib.gemm(w["h"], lp["wo2"], bias=lp["bo2"], out_bf16=w["y"])
            _lib.call("mvuld_ln_rows", w["y"], w["x32"], lp["g2"], lp["b2"], w["x32"], w["xb"], M, H, eps, 2)
        _lib.call("mvuld_masked_mean", w["x32"], w["len"], w["sent"], B, L, H)
        if check_suffix_padding and int(w["ok"].item()) != 1:
            raise ValueError("pad tokens must form a suffix of every sequence (unixcoder.py:150 tokenisation)")
        return w["x32"].view(B, L, H), w["sent"].clone()

    def forward(self, input_ids, attention_mask=None):
        """HF-style call ``encoder(ids, attention_mask=...)[0]``; the mask is re-derived from the pad id."""
        return (self.encode(input_ids)[0],)


class MyUniXcoder(nn.Module):
    """Mirror of unixcoder.py:20-95."""

    def __init__(self, encoder, config, tokenizer=None, tokenize=None):
        super().__init__()
        self.encoder = encoder
        self.config = config
        self.tokenizer = tokenizer
        self.tokenize = tokenize
        self.classifier = nn.Linear(config.hidden_size, 2)
        self.max_source_length = 512

    def get_xcode_vec(self, source_ids):
        """unixcoder.py:33-38."""
        return self.encoder.encode(source_ids)

    def get_repr(self, input_ids, labels=None):
        """unixcoder.py:91-95."""
        source_ids = input_ids.view(-1, self.max_source_length)
        _, vec = self.get_xcode_vec(source_ids)
        return vec, labels

    @torch.no_grad()
    def forward(self, source_ids=None, labels=None):
        """unixcoder.py:40-54: softmax(classifier(vec)); with labels also the cross-entropy loss."""
        source_ids = source_ids.view(-1, self.max_source_length)
        _, vec = self.get_xcode_vec(source_ids)
        logits = torch.empty(vec.shape[0], 2, device=vec.device, dtype=torch.float32)
        _lib.call("mvuld_linear_small", vec, self.classifier.weight.detach().float().contiguous(),
                  self.classifier.bias.detach().float().contiguous(), logits, None, vec.shape[0], 2, vec.shape[1])
        prob = torch.softmax(logits, dim=-1)
        if labels is not None:
            return nn.functional.cross_entropy(logits, labels), prob
        return prob


def build_MyUniXcoder(config=None):
    """Random-init stand-in for mvuld/data/bigvul_dataset.py:87-99 (pretrained weights need the network)."""
    cfg = config or roberta_base_config()
    return MyUniXcoder(RobertaEncoder(cfg), cfg)
