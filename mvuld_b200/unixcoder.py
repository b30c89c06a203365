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
    def encode(self, source_ids: torch.Tensor, check_suffix_padding: bool = False, clone_tokens: bool = True):
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
        # the token embeddings live in the cached per-(B, L) workspace: callers get their own copy unless they ask for
        # the view (internal fast paths that only use the sentence vector)
        tok = w["x32"].view(B, L, H)
        return (tok.clone() if clone_tokens else tok), w["sent"].clone()

    def forward(self, input_ids, attention_mask=None):
        """HF-style call ``encoder(ids, attention_mask=...)[0]``; the mask is re-derived from the pad id."""
        return (self.encode(input_ids)[0],)

    def pack(self, line_ids, rows_per_pass: int = 64, policy: str = "auto") -> "PackedLines":
        """``pack_host`` + copy to the encoder's device."""
        if self._plan is None:
            self.prepare()
        return self.pack_host(line_ids, rows_per_pass, policy).to(self._plan["dev"], non_blocking=True)

    def pack_host(self, line_ids, rows_per_pass: int = 64, policy: str = "auto") -> "PackedLines":
        """Host side of ``encode_lines`` (data-loader work, like the tokenizer's padding it replaces): next-fit packing
        (``pack_lines``) and the per-token arrays the kernels need (ids, position ids restarting per sequence, the key
        range [lo, hi) of each token's sequence), vectorised with numpy, as PINNED host tensors grouped in passes of
        ``rows_per_pass`` rows of 512 tokens."""
        import numpy as np
        pad, L = int(self.config.pad_token_id), 512
        flat, lens = _lines_to_flat(line_ids, pad)
        n = int(lens.size)
        row_of, off_of, n_rows = pack_lines(lens.tolist(), L, policy)
        row_of, off_of = np.asarray(row_of, dtype=np.int64), np.asarray(off_of, dtype=np.int64)
        first = np.concatenate([[0], np.cumsum(lens)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        line = np.repeat(np.arange(n), lens)                               # line of every token
        t = np.arange(flat.size) - np.repeat(first, lens)                  # index inside its line
        dest = row_of[line] * L + off_of[line] + t
        ids = np.full(n_rows * L, pad, dtype=np.int64)
        pos = np.full(n_rows * L, pad, dtype=np.int32)                     # padding_idx of the position table
        lo = np.zeros(n_rows * L, dtype=np.int32)                          # tail padding: the kernel points it at key 0
        hi = np.ones(n_rows * L, dtype=np.int32)
        ids[dest] = flat
        pos[dest] = t + pad + 1                                            # HF create_position_ids_from_input_ids, per line
        lo[dest] = off_of[line]
        hi[dest] = off_of[line] + lens[line]
        used = np.zeros(n_rows, dtype=np.int32)
        np.maximum.at(used, row_of, (off_of + lens).astype(np.int32))
        # key tiles (128 keys) each 128-row query tile must visit: the sequences of a row lie back to back, so the union
        # of the key ranges of a tile's valid rows is [lo of its first valid token, hi of its last valid token)
        QT = 128
        nqt = L // QT
        first_tok = np.arange(n_rows * nqt, dtype=np.int64) * QT                           # global token index
        row_used = np.repeat(used.astype(np.int64), nqt)
        t_in_row = np.tile(np.arange(nqt, dtype=np.int64) * QT, n_rows)
        last_tok = first_tok + np.clip(row_used - t_in_row, 1, QT) - 1
        has = row_used > t_in_row
        tile_lo = np.where(has, lo[first_tok] // QT, 0).astype(np.int32)
        tile_hi = np.where(has, (hi[last_tok] + QT - 1) // QT, 1).astype(np.int32)
        pin = torch.cuda.is_available()
        h2d = lambda x: (torch.from_numpy(np.ascontiguousarray(x)).pin_memory() if pin
                         else torch.from_numpy(np.ascontiguousarray(x)))
        passes = []
        for r0 in range(0, n_rows, rows_per_pass):
            r1 = min(n_rows, r0 + rows_per_pass)
            sel = np.nonzero((row_of >= r0) & (row_of < r1))[0]                # the sequences of this pass
            sl = slice(r0 * L, r1 * L)
            passes.append(dict(R=r1 - r0, n=int(sel.size), ids=h2d(ids[sl].reshape(r1 - r0, L)), pos=h2d(pos[sl]),
                               lo=h2d(lo[sl]), hi=h2d(hi[sl]), used=h2d(used[r0:r1]),
                               tlo=h2d(tile_lo[r0 * nqt:r1 * nqt]), thi=h2d(tile_hi[r0 * nqt:r1 * nqt]),
                               start=h2d(((row_of[sel] - r0) * L + off_of[sel]).astype(np.int32)),
                               len=h2d(lens[sel].astype(np.int32)), dst=h2d(sel.astype(np.int32))))
        return PackedLines(n, n_rows, int(flat.size), passes)

    @torch.no_grad()
    def encode_packed(self, packed: "PackedLines") -> torch.Tensor:
        """Device side of ``encode_lines``: the encoder over packed rows with block-diagonal attention, then the mean
        over each line's tokens (unixcoder.py:37) -> fp32 [n_lines, H]."""
        if self.training:
            raise RuntimeError("mvuld_b200 RobertaEncoder implements the eval-mode forward: call model.eval()")
        if self._plan is None:
            self.prepare()
        p, cfg = self._plan, self.config
        H, nH, L = cfg.hidden_size, cfg.num_attention_heads, 512
        qmul = LOG2E / math.sqrt(H // nH)
        eps = float(cfg.layer_norm_eps)
        out = torch.empty(packed.n_lines, H, device=p["dev"], dtype=torch.float32)
        if packed.passes and not packed.passes[0]["ids"].is_cuda:
            raise RuntimeError("encode_packed takes a device-resident PackedLines: call .to(device) (no CPU fallback)")
        for ps in packed.passes:
            R = ps["R"]
            M = R * L
            w = self._workspace(R, L)
            _lib.call("mvuld_roberta_embed", ps["ids"], ps["pos"], p["word"], p["pos"], p["type0"], p["eg"], p["eb"],
                      w["x32"], w["xb"], M, H, eps)
            for lp in p["layers"]:
                _lib.call("mvuld_heads_qkv", w["xb"], lp["wqkv"], lp["bqkv"], w["q"], w["k"], w["v"], R, L, H, nH, qmul)
                _lib.call("mvuld_seq_attention_packed", w["q"], w["k"], w["v"], ps["used"], ps["lo"], ps["hi"], ps["tlo"],
                          ps["thi"], w["ctx"], R, L, nH, H // nH)
                _lib.gemm(w["ctx"], lp["wo"], bias=lp["bo"], out_bf16=w["y"])
                _lib.call("mvuld_ln_rows", w["y"], w["x32"], lp["g1"], lp["b1"], w["x32"], w["xb"], M, H, eps, 2)
                _lib.gemm(w["xb"], lp["wi"], bias=lp["bi"], act=_lib.ACT_GELU, out_bf16=w["h"])
                _lib.gemm(w["h"], lp["wo2"], bias=lp["bo2"], out_bf16=w["y"])
                _lib.call("mvuld_ln_rows", w["y"], w["x32"], lp["g2"], lp["b2"], w["x32"], w["xb"], M, H, eps, 2)
            _lib.call("mvuld_seq_segment_mean", w["x32"], ps["start"], ps["len"], ps["dst"], out, ps["n"], H)
        return out

    def encode_lines(self, line_ids, rows_per_pass: int = 64) -> torch.Tensor:
        """Sentence vectors of MANY SHORT sequences (the per-node line encoding of mvuld/data/data_list.py:292-299 via
        unixcoder.py:56-68, where the reference pads every line to 512 tokens): -> fp32 [n_lines, H] on the GPU.

        ``line_ids``: HOST int64 tensor [n, L <= 512] padded with ``pad_token_id`` (what ``tokenize(..., padding=True)``
        returns), or a list of token-id lists.  Lines are packed back to back into rows of 512 tokens (``pack``) and run
        with block-diagonal attention and per-line position ids (``encode_packed``), which is arithmetically what the
        reference computes for the valid tokens of each padded line; pad tokens are never the key of a valid query."""
        return self.encode_packed(self.pack(line_ids, rows_per_pass))


class PackedLines:
    """Sequences packed into rows of 512 tokens (see ``RobertaEncoder.pack_host``); host (pinned) or device resident."""

    def __init__(self, n_lines, n_rows, n_tokens, passes):
        self.n_lines, self.n_rows, self.n_tokens, self.passes = n_lines, n_rows, n_tokens, passes

    def to(self, device, non_blocking: bool = False) -> "PackedLines":
        mv = lambda v: v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        return PackedLines(self.n_lines, self.n_rows, self.n_tokens, [{k: mv(v) for k, v in ps.items()} for ps in self.passes])

    def tensors(self):
        for ps in self.passes:
            for v in ps.values():
                if isinstance(v, torch.Tensor):
                    yield v

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    @property
    def fill(self) -> float:
        """Fraction of the packed rows' token slots that hold real tokens."""
        return self.n_tokens / max(1, self.n_rows * 512)


def _lines_to_flat(line_ids, pad: int):
    """-> (all valid tokens back to back, int64; length per line, int64).  Pad tokens must form a suffix."""
    import numpy as np
    if isinstance(line_ids, torch.Tensor):
        if line_ids.is_cuda:
            raise RuntimeError("encode_lines packs on the host: pass the tokenizer's (CPU) ids")
        a = line_ids.to(torch.int64).numpy()
        if a.ndim != 2 or a.shape[1] > 512:
            raise ValueError("line ids must be [n_lines, L <= 512]")
        valid = a != pad
        lens = valid.sum(1).astype(np.int64)
        if not np.array_equal(valid, np.arange(a.shape[1])[None, :] < lens[:, None]):
            raise ValueError("pad tokens must form a suffix of every line (unixcoder.py:150 tokenisation)")
        flat = a[valid]
    else:
        rows = [np.asarray(r, dtype=np.int64) for r in line_ids]
        rows = [r[r != pad] for r in rows]
        lens = np.asarray([r.size for r in rows], dtype=np.int64)
        flat = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    if lens.size and (lens.min() < 1 or lens.max() > 512):
        raise ValueError("every line needs between 1 and 512 tokens (the tokenizer always emits <s> <encoder-only> "
                         "</s> ... </s>)")
    return flat, lens


def _lines_to_rows(line_ids, pad: int):
    """-> (list of 1-D int64 numpy rows without padding, list of lengths); helper kept for tests."""
    import numpy as np
    flat, lens = _lines_to_flat(line_ids, pad)
    ends = np.cumsum(lens)
    return [flat[e - l:e] for e, l in zip(ends, lens)], lens.tolist()


def pack_lines(lengths, L: int = 512, policy: str = "next_fit"):
    """Pack sequences of the given lengths into rows of ``L`` tokens -> (row per sequence, token offset per sequence,
    number of rows).

    ``next_fit``: input order, a new row whenever the next sequence does not fit (fill ~0.98 for code lines of ~20
    tokens, O(n)).  ``first_fit_decreasing``: longest first into the first row with room -- for whole functions
    (hundreds of tokens) next-fit leaves ~20 % of the slots empty, FFD a few per cent.  ``auto`` picks FFD when the mean
    length exceeds L / 8 and there are at most 8192 sequences."""
    import numpy as np
    lengths = [int(x) for x in lengths]
    if any(ln > L for ln in lengths):
        raise ValueError(f"a sequence does not fit a row of {L} tokens")
    n = len(lengths)
    if policy == "auto":
        policy = "first_fit_decreasing" if n and n <= 8192 and sum(lengths) / n > L / 8 else "next_fit"
    if policy == "next_fit":
        row_of, off_of = [], []
        row, used = 0, 0
        for ln in lengths:
            if used + ln > L:
                row, used = row + 1, 0
            row_of.append(row)
            off_of.append(used)
            used += ln
        return row_of, off_of, (row + 1 if n else 0)
    if policy != "first_fit_decreasing":
        raise ValueError(f"unknown packing policy {policy!r}")
    order = sorted(range(n), key=lambda i: -lengths[i])                 # stable: ties keep input order
    free = np.full(n, L, dtype=np.int64)                                 # at most n rows
    n_rows = 0
    row_of, off_of = [0] * n, [0] * n
    for i in order:
        ln = lengths[i]
        fits = free[:n_rows] >= ln
        r = int(np.argmax(fits)) if fits.any() else n_rows
        if r == n_rows:
            n_rows += 1
        row_of[i] = r
        off_of[i] = int(L - free[r])
        free[r] -= ln
    return row_of, off_of, n_rows


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

    def myEncode_ids(self, line_ids) -> torch.Tensor:
        """unixcoder.py:56-68 (``myEncode``) from token ids: one sentence vector per code line, [n_lines, 768].  The
        reference runs every line padded to 512 tokens through the encoder; here lines are packed (``encode_lines``)."""
        return self.encoder.encode_lines(line_ids)

    def myEncode(self, sents: list) -> torch.Tensor:
        """unixcoder.py:56-68: needs the tokenizer callable the reference passes to the constructor."""
        if self.tokenize is None:
            raise RuntimeError("MyUniXcoder.myEncode needs the `tokenize` callable (unixcoder.py:137-151); offline use "
                               "myEncode_ids with token ids")
        ids = [self.tokenize([' '.join(s.split())], max_length=512, padding=True)[0] for s in sents]
        return self.myEncode_ids(torch.tensor(ids, dtype=torch.long))

    def get_repr(self, input_ids, labels=None):
        """unixcoder.py:91-95.  ``input_ids``: the reference's ``[B, 512]`` ids (CUDA: run as padded rows; CPU: packed on
        the host first), or a ``PackedLines`` from ``encoder.pack_host(ids)`` made at data-loading time -- the padding
        of short functions (a Big-Vul function averages well under half of 512 tokens) then costs nothing."""
        if isinstance(input_ids, PackedLines):
            return self.encoder.encode_packed(input_ids), labels
        if not input_ids.is_cuda:
            return self.encoder.encode_lines(input_ids.view(-1, self.max_source_length)), labels
        source_ids = input_ids.view(-1, self.max_source_length)
        _, vec = self.encoder.encode(source_ids, clone_tokens=False)
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
