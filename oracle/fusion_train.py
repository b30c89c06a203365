"""Oracle (test infrastructure): the fusion model's TRAINING step in plain fp32 PyTorch with autograd.

Restates /root/reference/mvuld/models/GraphModel.py:150-211 in train mode as main_bigvul.py:294-342 drives it:
BatchNorm layers on batch statistics (biased variance, eps 1e-5), CrossEntropyLoss (mean over the batch),
dropout optional (the parity tests run it with p = 0; the reference uses 0.2 for feat_drop / mlpdropout /
hdropout).  The dead ``h_func`` branch (GraphModel.py:172,177) is not evaluated: ``fconly``, ``ln_text``, ``hbn``,
``hln``, ``hfc`` receive no gradient, which is why the reference wraps the model with
``find_unused_parameters=True`` (main_bigvul.py:162-164).  DGL semantics as in oracle.dgl_ops (parity unpinned).

``emulate_bf16=True`` rounds (straight-through for autograd) exactly the tensors the B200 path holds in bf16 -- GEMM
weights outside the Rs_GCN blocks, BatchNorm outputs that feed a GEMM, GATConv projections / outputs and the node-MLP
activations -- while every accumulation stays fp32.  It exists because the train-mode model is ill-conditioned at small
batch: the fp32 restatement itself moves its logits by ~5 % and its gradient vector by ~50 % when ONLY the weights are
rounded to bf16 (BatchNorm on batch statistics divides nearly batch-constant features by their tiny deviation, eight
times in the Rs_GCN chain and once more in final_fc_bn), so a bf16 implementation can only be checked tightly against
a reference that rounds at the same places.

Only ``tests/`` and the ``cpu_baseline`` leg of ``bench.py`` may import this module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import dgl_ops

TRAINABLE_PREFIXES = ("gat.", "gat2.", "fc.", "hidden.", "Rs_GCN_", "bn_text.", "fc_text.", "bn_gat.", "fc_gat.",
                      "bn_bbox.", "fc_bbox.", "swinbn.", "swinfc.", "final_fc.", "final_fc_bn.")


def is_trained(name: str) -> bool:
    """Parameters that receive a gradient in the reference forward (buffers excluded)."""
    if name.endswith(("running_mean", "running_var", "num_batches_tracked")):
        return False
    return name.startswith(TRAINABLE_PREFIXES)


def _bn_train(p, prefix, x, dim, eps=1e-5):
    """BatchNorm1d in training mode over every axis but ``dim`` (biased variance, as F.batch_norm)."""
    axes = [a for a in range(x.dim()) if a != dim]
    mean = x.mean(axes, keepdim=True)
    var = x.var(axes, unbiased=False, keepdim=True)
    shape = [1] * x.dim()
    shape[dim] = -1
    return (x - mean) / torch.sqrt(var + eps) * p[prefix + "weight"].view(shape) + p[prefix + "bias"].view(shape)


def _ste_bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 in the forward pass, identity in the backward pass."""
    return t + (t.detach().bfloat16().float() - t.detach())


def _gat(p, prefix, src_t, dst_t, x, H, out_feats, slope=0.2, q=lambda t: t):
    N = x.shape[0]
    z = q(F.linear(x, q(p[prefix + "fc.weight"]))).view(N, H, out_feats)
    el = (z * p[prefix + "attn_l"]).sum(-1)
    er = (z * p[prefix + "attn_r"]).sum(-1)
    e = F.leaky_relu(el[src_t] + er[dst_t], slope)
    emax = torch.full((N, H), -float("inf")).scatter_reduce(0, dst_t[:, None].expand_as(e), e.detach(), "amax")
    pexp = torch.exp(e - emax[dst_t])
    den = torch.zeros(N, H).index_add(0, dst_t, pexp)
    alpha = pexp / den[dst_t]
    out = torch.zeros(N, H, out_feats).index_add(0, dst_t, alpha[..., None] * z[src_t])
    return q(out + p[prefix + "bias"].view(1, H, out_feats))


def _rs_gcn(p, prefix, v):
    conv = lambda name, t: F.conv1d(t, p[prefix + name + ".weight"], p[prefix + name + ".bias"])
    g_v = conv("g", v).permute(0, 2, 1)
    theta = conv("theta", v).permute(0, 2, 1)
    phi = conv("phi", v)
    R = theta @ phi
    R = R / R.size(-1)
    y = (R @ g_v).permute(0, 2, 1).contiguous()
    return _bn_train(p, prefix + "W.1.", conv("W.0", y), 1) + v


def _graph_to_gcn_in(p, batch, max_node, q, taps=None):
    """GraphModel.py:163-189: GATConv x2 -> node MLP -> unbatch / pad -> slot BatchNorms -> fc_gat | fc_bbox -> concat.
    Returns the token-major [B, max_node, 512] input of the Rs_GCN chain."""
    lin = lambda name, t: F.linear(t, p[name + ".weight"], p[name + ".bias"])
    linq = lambda name, t: F.linear(q(t), q(p[name + ".weight"]), p[name + ".bias"])     # bf16 operands, fp32 accumulate
    src_t = torch.as_tensor(batch.src, dtype=torch.long)
    dst_t = torch.as_tensor(batch.dst, dtype=torch.long)
    h = q(batch.ndata["_UNIX_NODE_EMB"].float())
    pos = batch.ndata["pos_emb"].float()
    h = _gat(p, "gat.", src_t, dst_t, h, 4, 512, q=q).reshape(h.shape[0], -1)
    h = _gat(p, "gat2.", src_t, dst_t, h, 4, 512, q=q).reshape(h.shape[0], -1)
    h = q(F.elu(linq("fc", h)))
    for i in range(8):
        h = q(F.elu(linq(f"hidden.{i}", h)))
    if taps is not None:
        taps["node_mlp"] = h.detach().clone()
    h_i = dgl_ops.unbatch_pad(h, batch.batch_num_nodes, max_node)
    pos_i = dgl_ops.unbatch_pad(pos, batch.batch_num_nodes, max_node)
    h_i = F.elu(linq("fc_gat", _bn_train(p, "bn_gat.", h_i, 1)))
    pos_i = F.elu(lin("fc_bbox", _bn_train(p, "bn_bbox.", pos_i, 1)))
    return torch.cat([h_i, pos_i], 2)


def loss_and_grads(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", img: torch.Tensor, txt: torch.Tensor,
                   labels: torch.Tensor, max_node: int = 100, taps: dict | None = None,
                   emulate_bf16: bool = False, gcn_in: torch.Tensor | None = None):
    """One forward + backward of main_bigvul.py:328-339 (dropout off) -> (loss, logits, {name: grad}).

    ``taps`` receives intermediate activations (token-major).  ``gcn_in`` [B, max_node, 512] REPLACES the graph branch's
    output as the input of the Rs_GCN chain (a relay point: the chain amplifies a 0.4 % input perturbation to ~5 % at
    the logits, so the chain and the head are checked from the implementation's own ``gcn_in``); the gradient with
    respect to it is returned under the key ``"__gcn_in__"`` and upstream parameters get no gradient."""
    p = {k: v.detach().clone().float().requires_grad_(is_trained(k)) for k, v in sd.items()}
    q = _ste_bf16 if emulate_bf16 else (lambda t: t)
    lin = lambda name, t: F.linear(t, p[name + ".weight"], p[name + ".bias"])
    linq = lambda name, t: F.linear(q(t), q(p[name + ".weight"]), p[name + ".bias"])
    x = F.elu(linq("swinfc", _bn_train(p, "swinbn.", img.float(), 1)))
    t = F.elu(linq("fc_text", _bn_train(p, "bn_text.", txt.float(), 1)))
    if gcn_in is None:
        zin = _graph_to_gcn_in(p, batch, max_node, q, taps)
    else:
        zin = gcn_in.detach().clone().float().requires_grad_(True)
    if taps is not None:
        taps["gcn_in"] = zin.detach().clone()
    z = zin.permute(0, 2, 1)
    for k in range(1, 9):
        z = _rs_gcn(p, f"Rs_GCN_{k}.", z)
        if taps is not None:
            taps[f"gcn_{k}"] = z.detach().permute(0, 2, 1).clone()
    zt = z.permute(0, 2, 1)
    zt = zt / torch.pow(zt, 2).sum(dim=1, keepdim=True).sqrt()
    feats = torch.cat([x, zt.mean(dim=1), t], 1)
    if taps is not None:
        taps["feats"] = feats.detach().clone()
    logits = lin("final_fc", _bn_train(p, "final_fc_bn.", feats, 1))
    loss = F.cross_entropy(logits, labels.long())
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items() if v.requires_grad and v.grad is not None}
    if gcn_in is not None:
        grads["__gcn_in__"] = zin.grad.detach()
    return loss.detach(), logits.detach(), grads


def graph_branch_grads(sd: Dict[str, torch.Tensor], batch: "dgl_ops.HostBatch", cotangent: torch.Tensor,
                       max_node: int = 100, emulate_bf16: bool = False) -> Dict[str, torch.Tensor]:
    """Gradients of the graph branch's parameters (GATConv x2, node MLP, slot BatchNorms, fc_gat, fc_bbox) for a given
    cotangent d loss / d gcn_in [B, max_node, 512]: the backward half of the relay described in ``loss_and_grads``."""
    p = {k: v.detach().clone().float().requires_grad_(is_trained(k)) for k, v in sd.items()}
    q = _ste_bf16 if emulate_bf16 else (lambda t: t)
    zin = _graph_to_gcn_in(p, batch, max_node, q)
    zin.backward(cotangent.float())
    return {k: v.grad.detach() for k, v in p.items() if v.requires_grad and v.grad is not None}
