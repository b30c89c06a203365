"""Generate tests/golden/*.pt by running the REFERENCE code in this container.

Run here (``/root/reference`` is not present on the GPU box):  ``python tools/make_golden.py``

* SwinV2: ``/root/reference/mvuld/models/swin_transformer_v2.py`` imported unmodified through a 3-symbol
  ``timm.models.layers`` shim (DropPath = identity in eval, to_2tuple, trunc_normal_); weights come from the
  product classes under a fixed seed and are loaded with ``strict=True`` -- which also pins state-dict / buffer
  compatibility.
* Rs_GCN: ``/root/reference/mvuld/models/Rs_GCN.py`` imported unmodified.
* RoBERTa: HF ``transformers`` (5.5 here; the reference pins 4.18) ``RobertaModel`` in encoder mode with the 2-D key
  mask; valid-token rows and the pooled vector are what the reference's masked mean consumes.
* DGL-dependent ops have no runnable reference (dgl is not installable offline): their golden files are produced by
  ``oracle.dgl_ops`` itself and only guard against drift (parity unpinned, see oracle/__init__.py).

Only small tensors are stored: inputs and weights are regenerated from seeds by tests through the same functions.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

import mvuld_b200 as mv                      # noqa: E402
from mvuld_b200 import synth                 # noqa: E402
from tests import cases                      # noqa: E402


def _shim_timm():
    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training, "shim DropPath is eval-only"
            return x

    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@torch.no_grad()
def golden_swin():
    _shim_timm()
    ref = _load(os.path.join(REF, "mvuld/models/swin_transformer_v2.py"), "ref_swin_v2")
    out = {}
    for name, kw in cases.SWIN_CASES.items():
        model = cases.make_swin(name)
        ref_model = ref.SwinTransformerV2(**kw).eval()
        missing = ref_model.load_state_dict(model.state_dict(), strict=True)
        x = synth.images(cases.SWIN_BATCH[name], kw["img_size"], seed=cases.SEED)
        feats = ref_model.forward_features(x)
        logits = ref_model(x)
        out[name] = dict(features=feats.clone(), logits=logits.clone())
        print(name, "features", tuple(feats.shape), float(feats.abs().mean()), missing)
    # integer artefacts of the full-size geometry (too large to store: keep digests)
    wa = ref.WindowAttention(128, (28, 28), 4, pretrained_window_size=(12, 12))
    out["rpi28_sha"] = sha(wa.relative_position_index)
    out["rpi28_sum"] = int(wa.relative_position_index.sum())
    out["coords28"] = wa.relative_coords_table.clone()
    blk = ref.SwinTransformerBlock(128, (112, 112), 4, window_size=28, shift_size=14)
    out["mask112_sha"] = sha(blk.attn_mask)
    out["mask112_nonzero"] = int((blk.attn_mask != 0).sum())
    wa7 = ref.WindowAttention(128, (7, 7), 4, pretrained_window_size=(6, 6))
    out["rpi7"] = wa7.relative_position_index.clone()
    blk7 = ref.SwinTransformerBlock(128, (28, 28), 4, window_size=7, shift_size=3)
    out["mask28_ws7"] = blk7.attn_mask.clone()
    torch.save(out, os.path.join(OUT, "swin.pt"))
    for k in ("timm", "timm.models", "timm.models.layers"):      # the shim must not leak into transformers' probes
        sys.modules.pop(k, None)


def golden_swin_train():
    """Autograd through the UNMODIFIED reference SwinV2 module (small_ws7 case, DropPath / dropout inactive): pins the
    backward oracle oracle.swin.features_and_grads that the encoder-backward kernels will be checked against.  Per
    parameter: gradient norm, the first 16 entries and 16 strided samples; d/d image as norm + samples."""
    _shim_timm()
    ref = _load(os.path.join(REF, "mvuld/models/swin_transformer_v2.py"), "ref_swin_v2_train")
    name = "small_ws7"
    kw = cases.SWIN_CASES[name]
    model = cases.make_swin(name)
    ref_model = ref.SwinTransformerV2(**kw).eval()
    ref_model.load_state_dict(model.state_dict(), strict=True)
    x = synth.images(2, kw["img_size"], seed=cases.SEED + 21).requires_grad_(True)
    feats = ref_model.forward_features(x)
    cot = torch.randn(feats.shape, generator=torch.Generator().manual_seed(cases.SEED + 22))
    (feats * cot).sum().backward()

    def pack(g):
        f = g.detach().reshape(-1)
        stride = max(1, f.numel() // 16)
        return dict(norm=float(f.double().norm()), head=f[:16].clone(), strided=f[::stride][:16].clone(), numel=f.numel())

    out = dict(cotangent=cot, features=feats.detach().clone(), dx=pack(x.grad),
               grads={k: pack(p.grad) for k, p in ref_model.named_parameters() if p.grad is not None})
    torch.save(out, os.path.join(OUT, "swin_train.pt"))
    print("swin_train", len(out["grads"]), "parameter gradients; |dx|", out["dx"]["norm"])
    for k in ("timm", "timm.models", "timm.models.layers"):
        sys.modules.pop(k, None)


@torch.no_grad()
def golden_rs_gcn():
    ref = _load(os.path.join(REF, "mvuld/models/Rs_GCN.py"), "ref_rs_gcn")
    mine = cases.make_rs_gcn()
    m = ref.Rs_GCN(in_channels=512, inter_channels=512).eval()
    m.load_state_dict(mine.state_dict(), strict=True)
    v = cases.rs_gcn_input()
    v_star, R = m(v)
    torch.save(dict(v_star=v_star.clone(), R=R.clone()), os.path.join(OUT, "rs_gcn.pt"))
    print("rs_gcn", float(v_star.abs().mean()), float((v_star - v).abs().mean()))


def golden_rs_gcn_train():
    """Reference Rs_GCN.py in TRAIN mode (BatchNorm on batch statistics) with autograd: pins the train-mode oracle
    (oracle/fusion_train.py::_rs_gcn) that the CUDA training step is checked against.  Slices only (small file)."""
    ref = _load(os.path.join(REF, "mvuld/models/Rs_GCN.py"), "ref_rs_gcn_train")
    mine = cases.make_rs_gcn()
    m = ref.Rs_GCN(in_channels=512, inter_channels=512).train()
    m.load_state_dict(mine.state_dict(), strict=True)
    v = cases.rs_gcn_input().requires_grad_(True)
    v_star, R = m(v)
    gsel = torch.Generator().manual_seed(cases.SEED + 1)
    dout = torch.randn(v_star.shape, generator=gsel)
    (v_star * dout).sum().backward()
    torch.save(dict(v_star=v_star.detach()[:, :16].clone(), dv=v.grad[:, :16].clone(),
                    d_theta=m.theta.weight.grad[:16].clone(), d_phi_bias=m.phi.bias.grad.clone(),
                    d_g=m.g.weight.grad[:16].clone(), d_W0=m.W[0].weight.grad[:16].clone(),
                    d_bn_weight=m.W[1].weight.grad.clone(), d_bn_bias=m.W[1].bias.grad.clone(),
                    running_mean=m.W[1].running_mean.clone(), running_var=m.W[1].running_var.clone()),
               os.path.join(OUT, "rs_gcn_train.pt"))
    print("rs_gcn_train", float(v_star.abs().mean()), float(v.grad.abs().mean()))


@torch.no_grad()
def golden_roberta():
    from transformers import RobertaConfig, RobertaModel
    cfg = cases.roberta_small_config()
    hf_cfg = RobertaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           intermediate_size=cfg.intermediate_size,
                           max_position_embeddings=cfg.max_position_embeddings, type_vocab_size=cfg.type_vocab_size,
                           pad_token_id=cfg.pad_token_id, layer_norm_eps=cfg.layer_norm_eps,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, is_decoder=False)
    hf = RobertaModel(hf_cfg, add_pooling_layer=True).eval()
    mine = cases.make_roberta()
    sd = {k[len("encoder."):]: v for k, v in mine.state_dict().items() if k.startswith("encoder.")}
    res = hf.load_state_dict(sd, strict=False)
    assert not res.missing_keys or all("position_ids" in k or "token_type_ids" in k for k in res.missing_keys), res
    assert not res.unexpected_keys, res
    ids = synth.token_ids(cases.ROBERTA_BATCH, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED)
    mask = ids.ne(cfg.pad_token_id)
    tok = hf(ids, attention_mask=mask.long())[0]
    sent = (tok * mask.unsqueeze(-1)).sum(1) / mask.sum(-1).unsqueeze(-1)
    torch.save(dict(sent=sent.clone(), tok_valid_checksum=float((tok * mask.unsqueeze(-1)).abs().sum()),
                    tok=tok.clone(), mask=mask.clone()), os.path.join(OUT, "roberta.pt"))
    print("roberta sent", tuple(sent.shape), float(sent.abs().mean()))


def golden_roberta_train():
    """Autograd through the installed HF RobertaModel (encoder mode, key mask; dropout 0) of <masked-mean sentence
    vectors, cotangent>: pins oracle.roberta.sentence_and_grads.  Per parameter: norm + 32 samples."""
    from transformers import RobertaConfig, RobertaModel
    cfg = cases.roberta_small_config()
    hf_cfg = RobertaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           intermediate_size=cfg.intermediate_size,
                           max_position_embeddings=cfg.max_position_embeddings, type_vocab_size=cfg.type_vocab_size,
                           pad_token_id=cfg.pad_token_id, layer_norm_eps=cfg.layer_norm_eps,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, is_decoder=False)
    hf = RobertaModel(hf_cfg, add_pooling_layer=True).eval()
    mine = cases.make_roberta()
    sd = {k[len("encoder."):]: v for k, v in mine.state_dict().items() if k.startswith("encoder.")}
    hf.load_state_dict(sd, strict=False)
    ids = synth.token_ids(cases.ROBERTA_BATCH, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED + 31)
    mask = ids.ne(cfg.pad_token_id)
    tok = hf(ids, attention_mask=mask.long())[0]
    sent = (tok * mask.unsqueeze(-1)).sum(1) / mask.sum(-1).unsqueeze(-1)
    cot = torch.randn(sent.shape, generator=torch.Generator().manual_seed(cases.SEED + 32))
    (sent * cot).sum().backward()

    def pack(g):
        f = g.detach().reshape(-1)
        stride = max(1, f.numel() // 16)
        return dict(norm=float(f.double().norm()), head=f[:16].clone(), strided=f[::stride][:16].clone(), numel=f.numel())

    out = dict(cotangent=cot, sent=sent.detach().clone(),
               grads={"encoder." + k: pack(p.grad) for k, p in hf.named_parameters() if p.grad is not None})
    torch.save(out, os.path.join(OUT, "roberta_train.pt"))
    print("roberta_train", len(out["grads"]), "parameter gradients")


@torch.no_grad()
def golden_graph():
    """Oracle-generated (drift guard only): integer artefacts + fusion / GGNN outputs on small seeded batches."""
    from oracle import dgl_ops, fusion
    from tests.cases import to_host_batch
    g = synth.cpg_batch(cases.FUSION_BATCH, seed=cases.SEED)
    hb = to_host_batch(g)
    indptr, indices, eids = dgl_ops.in_csr(hb.src, hb.dst, hb.num_nodes)
    pmap = dgl_ops.pad_truncate_map(hb.batch_num_nodes, 100)
    model = cases.make_fusion()
    ge = torch.Generator().manual_seed(cases.SEED)
    img = torch.randn(cases.FUSION_BATCH, 1024, generator=ge)
    txt = torch.randn(cases.FUSION_BATCH, 768, generator=ge)
    logits = fusion.fusion_forward(model.state_dict(), hb, img, txt)
    gg = synth.ggnn_batch(cases.GGNN_BATCH, seed=cases.SEED, n_etypes=cases.GGNN_T)
    hg = to_host_batch(gg)
    gm = cases.make_ggnn()
    prob, logit, ssum, h = fusion.ggnn_sum_forward(gm.state_dict(), hg, cases.GGNN_D, cases.GGNN_STEPS, cases.GGNN_T)
    torch.save(dict(bnn=torch.from_numpy(hb.batch_num_nodes), indptr_sha=sha(torch.from_numpy(indptr)),
                    indices_sha=sha(torch.from_numpy(indices)), eids_sha=sha(torch.from_numpy(eids)),
                    pad_map_sha=sha(torch.from_numpy(pmap)), fusion_logits=logits.clone(),
                    ggnn_prob=prob.clone(), ggnn_logit=logit.clone(), ggnn_sum=ssum.clone(),
                    ggnn_h_checksum=float(h.abs().sum())), os.path.join(OUT, "graph.pt"))
    print("fusion logits", logits)
    print("ggnn prob", prob[:4])



# --------------------------------------------------------------------------------------------------------
# The reference's OWN fusion classes (GraphModel.py, new_model.py, myModels.py) run on CPU.  Their third-party
# graph library is absent (dgl 0.8.1), so `dgl` is stubbed with the oracle's restatement of the
# three calls the classes make (GATConv.forward, dgl.unbatch, dgl.mean_nodes); everything else -- the class
# constructors, the Python unbatch / pad loops, the Rs_GCN module, the order and wiring of every layer -- is the
# reference's code, unmodified, loaded from /root/reference.  Pins oracle.fusion.{fusion_forward, gat_variant_forward,
# ablation_forward, variant2_forward, gating_forward} and the mirror classes' state-dict keys (strict load).
# --------------------------------------------------------------------------------------------------------
def _stub_dgl():
    import contextlib
    from oracle import dgl_ops

    class _Sub:
        def __init__(self, ndata):
            self.ndata = ndata

        def number_of_nodes(self):
            return next(iter(self.ndata.values())).shape[0]

    class RefGraph:
        """What the reference forwards touch of a batched DGLGraph: ndata, local_scope, unbatch, mean_nodes."""

        def __init__(self, hb):
            self.hb = hb
            self.ndata = dict(hb.ndata)

        @contextlib.contextmanager
        def local_scope(self):
            saved = dict(self.ndata)
            try:
                yield
            finally:
                self.ndata = saved

    def unbatch(g):
        off = dgl_ops.node_offsets(g.hb.batch_num_nodes)
        return [_Sub({k: v[off[i]:off[i + 1]] for k, v in g.ndata.items()}) for i in range(len(off) - 1)]

    def mean_nodes(g, key):
        return dgl_ops.mean_nodes(g.ndata[key], g.hb.batch_num_nodes)

    class GATConv(nn.Module):
        def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2):
            super().__init__()
            self.h, self.o, self.slope = num_heads, out_feats, negative_slope
            self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
            self.attn_l = nn.Parameter(torch.zeros(1, num_heads, out_feats))
            self.attn_r = nn.Parameter(torch.zeros(1, num_heads, out_feats))
            self.bias = nn.Parameter(torch.zeros(num_heads * out_feats))

        def forward(self, g, feat):
            assert not self.training
            return dgl_ops.gat_conv(self.state_dict(), "", g.hb.src, g.hb.dst, feat, self.h, self.o, self.slope)

    dgl = types.ModuleType("dgl")
    dgl.unbatch, dgl.mean_nodes = unbatch, mean_nodes
    dnn = types.ModuleType("dgl.nn")
    dpt = types.ModuleType("dgl.nn.pytorch")
    dpt.GATConv, dpt.GraphConv, dpt.GatedGraphConv = GATConv, None, None
    dgl.nn, dnn.pytorch = dnn, dpt
    utils = types.ModuleType("utils")
    for n in ("load_checkpoint", "auto_resume_helper", "reduce_tensor", "resume_bestf1_helper", "save_bestf1_checkpoint"):
        setattr(utils, n, None)
    sys.modules.update({"dgl": dgl, "dgl.nn": dnn, "dgl.nn.pytorch": dpt, "utils": utils})
    return RefGraph


def _load_ref_models():
    """GraphModel / new_model / myModels of the reference as submodules of a synthetic package whose unrelated
    siblings (`.build`, `.fusion`, the other Swin files) are empty stubs."""
    RefGraph = _stub_dgl()
    _shim_timm()
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "mvuld", "models")]
    sys.modules["refmodels"] = pkg
    for stub in ("build", "fusion", "swin_transformer_v2"):
        sys.modules[f"refmodels.{stub}"] = types.ModuleType(f"refmodels.{stub}")
    mods = {}
    for name in ("GraphModel", "new_model", "myModels"):
        mods[name] = importlib.import_module(f"refmodels.{name}")
    return RefGraph, mods


@torch.no_grad()
def golden_fusion_classes():
    from tests.cases import to_host_batch, fusion_class_cases, fusion_class_inputs, fusion_class_model, FUSION_CLASS_BATCH
    RefGraph, mods = _load_ref_models()
    g, img, txt = fusion_class_inputs()
    hb = to_host_batch(g)
    out = {}
    for key, modname, clsname, mirror in fusion_class_cases():
        m = fusion_class_model(key, mirror)
        ref = getattr(mods[modname], clsname)(mv.default_config()).eval()
        ref.load_state_dict(m.state_dict(), strict=True)                # pins the mirror's key set and shapes
        rg = RefGraph(hb)
        rg.ndata["_ALL_NODE_EMB"] = torch.zeros(hb.num_nodes, 800)     # read and never used by myModels.py:357
        logits = ref(rg, img.clone(), txt.clone())
        assert logits.shape == (FUSION_CLASS_BATCH, 2) and torch.isfinite(logits).all(), key
        out[key] = dict(logits=logits.clone(), keys=sorted(m.state_dict().keys()))
        print(f"{key:36s}", logits[0].tolist())
    torch.save(out, os.path.join(OUT, "fusion_classes.pt"))

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if "rs_gcn_train" in sys.argv[1:]:
        golden_rs_gcn_train()
        sys.exit(0)
    if "fusion_classes" in sys.argv[1:]:
        golden_fusion_classes()
        sys.exit(0)
    if "swin_train" in sys.argv[1:] or "roberta_train" in sys.argv[1:]:
        if "swin_train" in sys.argv[1:]:
            golden_swin_train()
        if "roberta_train" in sys.argv[1:]:
            golden_roberta_train()
        sys.exit(0)
    golden_rs_gcn_train()
    golden_swin_train()
    golden_roberta_train()
    golden_swin()
    golden_rs_gcn()
    golden_roberta()
    golden_graph()
    print("golden files written to", OUT)
