"""ctypes binding of ``libmvuld_b200.so`` (the C ABI declared in ``include/mvuld_b200.h``).

PyTorch is used only for device memory and the current CUDA stream: every wrapper passes raw device pointers,
plain ints and the stream handle.  There is no CPU or PyTorch fallback: if the library is missing or a call
fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVULD_LIB") or os.path.join(_HERE, "libmvuld_b200.so")   # MVULD_LIB: an experimental build of the same library

_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argument ctypes (the stream argument is included); mirrors include/mvuld_b200.h
SIGNATURES = {
    "mvuld_gemm_bf16": [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _I, _P],
    "mvuld_gemm_ln_bf16": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P],
    "mvuld_gemm_ln_wide_bf16": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P],
    "mvuld_mlp_ln_bf16": [_P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _P],
    "mvuld_swin_qkv": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_heads_qkv": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "mvuld_cpb_table": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "mvuld_swin_window_attention": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_swin_window_attention_fixed": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_swin_window_attention_train": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_swin_attention_bwd_prep": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_swin_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_seq_attention_train": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_seq_attention_bwd_prep": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_seq_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_seq_qkv_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_roberta_embed_train": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "mvuld_masked_mean_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "mvuld_embed_grad_rows": [_P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_swin_bias_grad": [_P, _I, _I, _I, _I, _P, _P, _P],
    "mvuld_swin_bias_grad_splits": [_I, _I, _I],
    "mvuld_cpb_mlp_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "mvuld_swin_qkv_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_swin_qkv_bwd_blocks": [_I, _I, _I, _I],
    "mvuld_gelu_fwd": [_P, _P, _LL, _P],
    "mvuld_patch_merge_scatter": [_P, _P, _I, _I, _I, _I, _P],
    "mvuld_patch_im2col": [_P, _P, _I, _I, _I, _P],
    "mvuld_swin_qkv_train": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mvuld_seq_attention": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_ln_rows": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "mvuld_patch_embed": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "mvuld_patch_merge_gather": [_P, _P, _I, _I, _I, _I, _P],
    "mvuld_ln_meanpool": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "mvuld_seq_positions": [_P, _I, _I, _I, _P, _P, _P, _P],
    "mvuld_roberta_embed": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "mvuld_masked_mean": [_P, _P, _P, _I, _I, _I, _P],
    "mvuld_csr_from_coo": [_P, _P, _I, _I, _P, C.POINTER(C.c_size_t), _P, _P, _P, _P, _P],
    "mvuld_gather_etype": [_P, _P, _I, _I, _P, _P, _P],
    "mvuld_ggnn_gather_sum": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_gemm_gru": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "mvuld_ggnn_init": [_P, _P, _P, _I, _LL, _I, _I, _P],
    "mvuld_segment_sum": [_P, _P, _P, _I, _I, _P],
    "mvuld_gat_scores": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_gat_aggregate": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P],
    "mvuld_unbatch_pad_bn": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_pos_branch": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvuld_f32_to_bf16": [_P, _P, _LL, _P],
    "mvuld_collate_edges": [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _LL, _P, _P],
    "mvuld_seq_attention_packed": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_seq_segment_mean": [_P, _P, _P, _P, _P, _I, _I, _P],
    "mvuld_rs_gcn_affinity_f32": [_P, _P, _P, _I, _I, _I, _P],
    "mvuld_split3_bf16": [_P, _I, _P, _I, _I, _I, _P],
    "mvuld_fusion_head_mode": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvuld_linear_small": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_transpose_bf16": [_P, _I, _P, _I, _I, _I, _P],
    "mvuld_transpose_bf16_batched": [_P, _P, _I, _I, _P],
    "mvuld_gemm_dw": [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    "mvuld_gemm_dw_workspace": [_I, _I, _I],
    "mvuld_colsum": [_P, _I, _I, _P, _P, _I, _I, _P],
    "mvuld_colsum_slabs": [_I, _I],
    "mvuld_ln_rows_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "mvuld_ln_rows_bwd_blocks": [_I],
    "mvuld_gelu_bwd": [_P, _P, _P, _LL, _P],
    "mvuld_gelu_bwd_colsum": [_P, _P, _P, _P, _P, _I, _I, _P],
    "mvuld_elu_bwd": [_P, _P, _P, _LL, _I, C.c_ulonglong, _F, _P],
    "mvuld_dropout_bf16": [_P, _P, _LL, C.c_ulonglong, _F, _P],
    "mvuld_bn_cols_fwd": [_P, _P, _P, _F, _P, _I, _P, _I, _P, _P, _P, _P, _P, _F, _I, _I, _P],
    "mvuld_bn_cols_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "mvuld_elu_bwd_rows": [_P, _I, _P, _I, _P, _I, _I, _I, _P],
    "mvuld_pos_slot_stats": [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _F, _I, _I, _P],
    "mvuld_pos_branch_bwd": [_P] * 13 + [_I, _I, _I, _I, _I, _P],
    "mvuld_bn_slot_fwd": [_P, _P, _P, _F, _P, _P, _P, _P, _P, _F, _I, _I, _I, _P],
    "mvuld_bn_slot_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_unbatch_pad_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "mvuld_gat_bwd": [_P] * 18 + [_I, _I, _I, _F, _P],
    "mvuld_rs_gcn_affinity_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "mvuld_l2norm_mean_fwd": [_P, _P, _I, _P, _I, _I, _I, _P],
    "mvuld_l2norm_mean_bwd": [_P, _P, _P, _I, _P, _P, _I, _I, _I, _P],
    "mvuld_ce_loss": [_P, _P, _P, _P, _I, _I, _F, _P],
    "mvuld_linear_small_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_sumsq_f32": [_P, _LL, _P, _P, _P],
    "mvuld_adamw": [_P, _P, _P, _P, _LL, _P, _P, _I, _P, _F, _F, _F, _F, _F, _I, _P],
    "mvuld_node_linear4": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvuld_unbatch_pad_bn_elu": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_gru_sequence_workspace": [_I, _I],
    "mvuld_gru_sequence": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvuld_gate_fusion": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
}

_SYNC_EACH = bool(os.environ.get("MVULD_SYNC_EACH"))     # debug: synchronise after every call and name the one that faulted
_lib = None
launch_count = 0     # kernels launched through this binding (bench.py reports it as gpu_launches)
_LAUNCHES_PER_CALL = {"mvuld_cpb_table": 2, "mvuld_csr_from_coo": 5, "mvuld_gat_bwd": 3, "mvuld_sumsq_f32": 2,
                      "mvuld_ln_rows_bwd": 2, "mvuld_gelu_bwd_colsum": 2, "mvuld_pos_branch_bwd": 2, "mvuld_swin_bias_grad": 2,
                      "mvuld_swin_qkv_bwd": 2, "mvuld_colsum": 2, "mvuld_gemm_dw": 2}


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the MVulD hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.mvuld_last_error.restype = C.c_char_p
    lib.mvuld_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_longlong if name.endswith("_workspace") else C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


class _Raw:
    """A tensor passed by base pointer only (row stride given separately, so it need not be contiguous)."""

    def __init__(self, t: torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("mvuld_b200 kernels take CUDA tensors only (no CPU fallback)")
        self.ptr = C.c_void_p(t.data_ptr())


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mvuld_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("mvuld_b200 kernels take contiguous tensors")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name: str, *args):
    """Invoke ``name`` with tensors (-> device pointers), ints and floats; appends the current stream."""
    global launch_count
    lib = load()
    conv = []
    for a in args:
        if isinstance(a, _Raw):
            conv.append(a.ptr)
        elif isinstance(a, torch.Tensor) or a is None:
            conv.append(_ptr(a))
        else:
            conv.append(a)
    rc = getattr(lib, name)(*conv, _stream())
    if rc != 0:
        msg = lib.mvuld_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)
    if _SYNC_EACH:
        _sync_check(name)
    return rc


def gemm_dw(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, n_out: Optional[int] = None):
    """out [n_out, k_in] fp32 = dy[:, :n_out]^T @ x  (dy [M, *] and x [M, k_in] bf16 row-major: no transposed copies)."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dy.stride(1) == 1 and x.stride(1) == 1
    assert out.dtype == torch.float32 and out.stride(1) == 1
    M, k_in = x.shape
    n_out = dy.shape[1] if n_out is None else n_out
    assert dy.shape[0] == M and tuple(out.shape) == (n_out, k_in)
    need = load().mvuld_gemm_dw_workspace(M, n_out, k_in)
    part = torch.empty(need, device=out.device, dtype=torch.float32) if need else None
    call("mvuld_gemm_dw", _Raw(dy), dy.stride(0), _Raw(x), x.stride(0), _Raw(out), out.stride(0), part, M, n_out, k_in)


def colsum(x, is_bf16: int, ldx: int, out: torch.Tensor, R: int, C: int):
    """out[c] += sum_r x[r, c] (x: a tensor or a ``_Raw`` base pointer with row stride ``ldx``); fixed summation order."""
    slabs = load().mvuld_colsum_slabs(int(R), int(C))
    part = torch.empty(slabs * C, device=out.device, dtype=torch.float32) if slabs > 1 else None
    call("mvuld_colsum", x, is_bf16, ldx, out, part, R, C)


def gelu_bwd_colsum(pre: torch.Tensor, dh: torch.Tensor, dpre: torch.Tensor, dbias: torch.Tensor):
    """dpre = dh * GELU'(pre) (bf16 [M, C]) and dbias += column sums of dpre, one pass (fixed summation order)."""
    M, Cc = pre.shape
    assert pre.is_contiguous() and dh.is_contiguous() and dpre.is_contiguous() and dh.shape == pre.shape == dpre.shape
    slabs = load().mvuld_colsum_slabs(int(M), int(Cc))
    part = torch.empty(slabs * Cc, device=pre.device, dtype=torch.float32) if slabs > 1 else None
    call("mvuld_gelu_bwd_colsum", pre, dh, dpre, dbias, part, M, Cc)


class TransposeTable:
    """bf16 [R, C] -> [C, Rp] (Rp = R rounded up to 8, zero filled) for a fixed list of (source, destination) pairs in ONE
    launch (``mvuld_transpose_bf16_batched``).  The table lives on the device; the tensors must stay where they are."""

    def __init__(self, pairs):
        desc, ends, tot = [], [], 0
        for x, out in pairs:
            R, Cc = x.shape
            assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and x.stride(1) == 1 and out.is_contiguous()
            assert out.shape[0] == Cc and out.shape[1] >= R
            ldo = out.shape[1]
            desc += [x.data_ptr(), out.data_ptr(), R, Cc, x.stride(0), ldo]
            tot += ((ldo + 31) // 32) * ((Cc + 31) // 32)
            ends.append(tot)
        dev = pairs[0][0].device
        self.keep = pairs
        self.desc = torch.tensor(desc, dtype=torch.int64, device=dev)
        self.ends = torch.tensor(ends, dtype=torch.int32, device=dev)
        self.n, self.total = len(pairs), tot

    def run(self):
        call("mvuld_transpose_bf16_batched", self.desc, self.ends, self.n, self.total)


def ln_rows_bwd_partials(M: int, C: int, device) -> torch.Tensor:
    """Workspace of ``mvuld_ln_rows_bwd`` for M rows of C columns (fixed-order dgamma / dbeta reduction)."""
    return torch.empty(load().mvuld_ln_rows_bwd_blocks(int(M)) * 3 * C, device=device, dtype=torch.float32)


def _sync_check(name: str):
    try:
        torch.cuda.synchronize()
    except Exception as exc:                                  # noqa: BLE001
        raise RuntimeError(f"{name}: device fault surfaced after this call: {exc}") from exc


# ---------------------------------------------------------------------------------------------------------
# thin typed wrappers
# ---------------------------------------------------------------------------------------------------------
ACT_NONE, ACT_GELU, ACT_ELU = 0, 1, 2


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, act=ACT_NONE, res=None, out_bf16=None, out_f32=None,
         ldc=None):
    """out = act(a @ w.T + bias) + res.  a [M,K] bf16, w [N,K] bf16 (row strides may exceed K)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    lda, ldw = a.stride(0), w.stride(0)
    assert a.stride(1) == 1 and w.stride(1) == 1
    out = out_bf16 if out_bf16 is not None else out_f32
    if ldc is None:
        ldc = out.stride(0)
    if out_bf16 is not None and out_f32 is not None:
        assert out_bf16.stride(0) == out_f32.stride(0)
    ldr = res.stride(0) if res is not None else 0
    global launch_count
    lib = load()
    rc = lib.mvuld_gemm_bf16(C.c_void_p(a.data_ptr()), lda, C.c_void_p(w.data_ptr()), ldw, M, N, K, _ptr(bias), act,
                             C.c_void_p(res.data_ptr()) if res is not None else None, ldr,
                             C.c_void_p(out_bf16.data_ptr()) if out_bf16 is not None else None,
                             C.c_void_p(out_f32.data_ptr()) if out_f32 is not None else None, ldc, _stream())
    if rc != 0:
        raise RuntimeError(f"mvuld_gemm_bf16 failed (code {rc}): {lib.mvuld_last_error().decode()}")
    launch_count += 1
    if _SYNC_EACH:
        _sync_check(f"mvuld_gemm_bf16[M={M},N={N},K={K}]")


def gemm_ln(a: torch.Tensor, w: torch.Tensor, gamma, beta, eps: float, bias=None, shortcut=None, x32=None, xb=None):
    """x = shortcut + LayerNorm(a @ w.T + bias) * gamma + beta  (N = w.shape[0] in {128, 256, 512})."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    for t in (shortcut, x32, xb):
        assert t is None or (t.is_contiguous() and t.shape == (M, N))
    call("mvuld_gemm_ln_bf16", _Raw(a), a.stride(0), _Raw(w), w.stride(0), M, N, K, bias, gamma, beta, float(eps),
         shortcut, x32, xb)


def mlp_ln(x: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor, b2, gamma, beta, eps: float, shortcut=None, x32=None,
           xb=None):
    """x = shortcut + LayerNorm(fc2(GELU(fc1(x) + b1)) + b2) * gamma + beta in one launch (C = x.shape[1] in {128, 256};
    csrc/mlp_ln.cu).  ``xb`` may be ``x`` itself."""
    M, Cc = x.shape
    assert x.dtype == torch.bfloat16 and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    assert x.is_contiguous() and w1.is_contiguous() and w2.is_contiguous()
    assert w1.shape == (4 * Cc, Cc) and w2.shape == (Cc, 4 * Cc)
    for t in (shortcut, x32, xb):
        assert t is None or (t.is_contiguous() and t.shape == (M, Cc))
    call("mvuld_mlp_ln_bf16", _Raw(x), _Raw(w1), b1, _Raw(w2), b2, gamma, beta, float(eps), shortcut, x32, xb, M, Cc)


def gemm_ln_wide(a: torch.Tensor, w: torch.Tensor, gamma, beta, eps: float, bias=None, shortcut=None, x32=None, xb=None):
    """N = w.shape[0] in {512, 768, 1024}: x = shortcut + LayerNorm(a @ w.T + bias) * gamma + beta in one cluster launch
    (csrc/gemm_ln_cluster.cu)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    for t in (shortcut, x32, xb):
        assert t is None or (t.is_contiguous() and t.shape == (M, N))
    call("mvuld_gemm_ln_wide_bf16", _Raw(a), a.stride(0), _Raw(w), w.stride(0), M, N, K, bias, gamma, beta, float(eps),
         shortcut, x32, xb)


def csr_from_coo(src: torch.Tensor, dst: torch.Tensor, num_nodes: int):
    """In-edge CSR sorted by (dst, edge id).  Returns (indptr int32 [N+1], idx_src int32 [E], eids int32 [E])."""
    assert src.dtype == torch.int64 and dst.dtype == torch.int64 and src.is_cuda
    E = src.numel()
    dev = src.device
    need = C.c_size_t(0)
    lib = load()
    rc = lib.mvuld_csr_from_coo(None, None, E, num_nodes, None, C.byref(need), None, None, None, None, _stream())
    if rc != 0:
        raise RuntimeError(f"mvuld_csr_from_coo(size query) failed: {lib.mvuld_last_error().decode()}")
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    indptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    idx_src = torch.empty(E, dtype=torch.int32, device=dev)
    eids = torch.empty(E, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.mvuld_csr_from_coo(_ptr(src), _ptr(dst), E, num_nodes, _ptr(ws), C.byref(need), _ptr(indptr),
                                _ptr(idx_src), _ptr(eids), _ptr(status), _stream())
    if rc != 0:
        raise RuntimeError(f"mvuld_csr_from_coo failed (code {rc}): {lib.mvuld_last_error().decode()}")
    global launch_count
    launch_count += 5
    return indptr, idx_src, eids, status
