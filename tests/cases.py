"""Seeded test cases shared by the golden generator (tools/make_golden.py) and the tests."""
from __future__ import annotations

import numpy as np
import torch

import mvuld_b200 as mv
from mvuld_b200 import synth

SEED = 12345

# small geometries keep the CPU oracle fast; "full" is the BASELINE SwinV2-B 448 / window 28 model (2 images)
SWIN_CASES = {
    "small_ws7": dict(img_size=112, patch_size=4, in_chans=3, num_classes=2, embed_dim=128, depths=[2, 2, 2],
                      num_heads=[4, 8, 16], window_size=7, pretrained_window_sizes=[6, 6, 6], drop_path_rate=0.2),
    "mid_ws14": dict(img_size=224, patch_size=4, in_chans=3, num_classes=2, embed_dim=128, depths=[2, 2],
                     num_heads=[4, 8], window_size=14, pretrained_window_sizes=[12, 12], drop_path_rate=0.0),
    "full": dict(img_size=448, patch_size=4, in_chans=3, num_classes=2, embed_dim=128, depths=[2, 2, 18, 2],
                 num_heads=[4, 8, 16, 32], window_size=28, pretrained_window_sizes=[12, 12, 12, 6],
                 drop_path_rate=0.2),
}
SWIN_BATCH = {"small_ws7": 3, "mid_ws14": 2, "full": 2}


@torch.no_grad()
def round_matrices_to_bf16(model):
    """Make every weight MATRIX of ``model`` bf16-representable (values a bf16 checkpoint would hold), in place: the
    B200 path packs matrices as bf16 GEMM operands, so with such weights the oracle and the path compute with
    identical numbers and a parity test measures arithmetic, not weight quantisation.  Vectors (biases, norms) stay.
    Every model-level parity case uses such weights (north_star: "identical random-init weights"); the effect of
    quantising fp32 weights to bf16 is measured separately (tests/test_gpu_fullsize.py)."""
    for p in model.parameters():
        if p.dim() >= 2:
            p.copy_(p.to(torch.bfloat16).to(p.dtype))
    return model


def make_swin(name: str) -> mv.SwinTransformerV2:
    torch.manual_seed(SEED)
    m = mv.SwinTransformerV2(**SWIN_CASES[name]).eval()
    return round_matrices_to_bf16(synth.randomize_for_parity(m, seed=SEED))


def swin_geometry(name: str):
    from oracle.swin import SwinGeometry
    kw = SWIN_CASES[name]
    return SwinGeometry(img_size=kw["img_size"], embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]),
                        num_heads=tuple(kw["num_heads"]), window_size=kw["window_size"],
                        pretrained_window_sizes=tuple(kw["pretrained_window_sizes"]), num_classes=kw["num_classes"])


ROBERTA_BATCH, ROBERTA_L = 3, 512


def roberta_small_config():
    return mv.roberta_base_config(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                  intermediate_size=256)


def make_roberta(full: bool = False) -> mv.MyUniXcoder:
    torch.manual_seed(SEED)
    m = mv.build_MyUniXcoder(None if full else roberta_small_config()).eval()
    return round_matrices_to_bf16(synth.randomize_for_parity(m, seed=SEED))


def roberta_geometry(cfg):
    from oracle.roberta import RobertaGeometry
    return RobertaGeometry(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           intermediate_size=cfg.intermediate_size,
                           max_position_embeddings=cfg.max_position_embeddings, type_vocab_size=cfg.type_vocab_size,
                           pad_token_id=cfg.pad_token_id, layer_norm_eps=cfg.layer_norm_eps)


def make_rs_gcn() -> mv.Rs_GCN:
    torch.manual_seed(SEED)
    m = mv.Rs_GCN(512, 512).eval()
    return synth.randomize_for_parity(m, seed=SEED)


def rs_gcn_input() -> torch.Tensor:
    g = torch.Generator().manual_seed(SEED)
    return torch.randn(2, 512, 100, generator=g)


FUSION_BATCH = 4


def make_fusion() -> mv.Multi_DefectModel_new_GCN:
    torch.manual_seed(SEED)
    m = mv.Multi_DefectModel_new_GCN(mv.default_config()).eval()
    return round_matrices_to_bf16(synth.randomize_for_parity(m, seed=SEED))


GGNN_BATCH, GGNN_D, GGNN_T, GGNN_STEPS, GGNN_IN = 6, 200, 4, 6, 132


def make_ggnn() -> mv.GGNNSum:
    torch.manual_seed(SEED)
    m = mv.GGNNSum(GGNN_IN, GGNN_D, max_edge_types=GGNN_T, num_steps=GGNN_STEPS).eval()
    return round_matrices_to_bf16(synth.randomize_for_parity(m, seed=SEED))


def to_host_batch(g):
    """mvuld_b200.graph.Graph (CPU) -> oracle.dgl_ops.HostBatch."""
    from oracle.dgl_ops import HostBatch
    src, dst = g.edges()
    return HostBatch(src.cpu().numpy().astype(np.int64), dst.cpu().numpy().astype(np.int64), g.num_nodes(),
                     g.batch_num_nodes().cpu().numpy().astype(np.int64),
                     g.batch_num_edges().cpu().numpy().astype(np.int64),
                     {k: v.cpu() for k, v in g.ndata.items()}, {k: v.cpu() for k, v in g.edata.items()})


# ---- every fusion class against the reference's own class code (tests/golden/fusion_classes.pt) ----
def fusion_class_cases():
    """(golden key, reference module, class name, mirror class) of every fusion class."""
    import mvuld_b200 as mv
    from mvuld_b200 import my_models
    out = [("Multi_DefectModel_new_GCN", "GraphModel", "Multi_DefectModel_new_GCN", mv.Multi_DefectModel_new_GCN),
           ("Multi_DefectModel", "GraphModel", "Multi_DefectModel", mv.Multi_DefectModel),
           ("myModels.Multi_DefectModel", "myModels", "Multi_DefectModel", my_models.Multi_DefectModel)]
    for n, c in {**mv.ABLATIONS, **mv.GRID_VARIANTS}.items():
        out.append((n, "new_model" if n in ("Multi_DefectModel_noFunc", "Multi_DefectModel_noGlobalImage") else "GraphModel",
                    n, c))
    return out


FUSION_CLASS_BATCH = 5


def fusion_class_inputs():
    from mvuld_b200 import synth
    g = synth.cpg_batch(FUSION_CLASS_BATCH, seed=SEED + 13)       # graph sizes on both sides of the 100-slot truncation
    gen = torch.Generator().manual_seed(3)
    img, txt = torch.randn(FUSION_CLASS_BATCH, 1024, generator=gen), torch.randn(FUSION_CLASS_BATCH, 768, generator=gen)
    return g, img, txt


def fusion_class_model(key, cls):
    import mvuld_b200 as mv
    from mvuld_b200 import synth
    torch.manual_seed(SEED)
    m = cls(mv.default_config()).eval()
    round_matrices_to_bf16(synth.randomize_for_parity(m, seed=SEED))
    return m
