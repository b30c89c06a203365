"""CPU: host-side mirror of the reference interface, graph container semantics, C-ABI surface, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import mvuld_b200 as mv
from mvuld_b200 import _lib, synth
from mvuld_b200 import graph as G
from oracle import dgl_ops
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_mirrors_reference_defaults_and_yaml():
    cfg = mv.default_config()
    assert cfg.MODEL.TYPE == "swinv2" and cfg.DATA.IMG_SIZE == 448
    assert cfg.MODEL.SWINV2.DEPTHS == [2, 2, 18, 2] and cfg.MODEL.SWINV2.NUM_HEADS == [4, 8, 16, 32]
    assert cfg.MODEL.SWINV2.WINDOW_SIZE == 28 and cfg.MODEL.SWINV2.PRETRAINED_WINDOW_SIZES == [12, 12, 12, 6]
    assert cfg.MODEL.DROP_PATH_RATE == 0.2 and cfg.MODEL.NUM_CLASSES == 2
    assert cfg.TRAIN.BASE_LR == 5e-5 and cfg.TRAIN.WEIGHT_DECAY == 0.005 and cfg.TRAIN.CLIP_GRAD == 5.0
    assert cfg.OUTPUT.endswith(os.path.join(cfg.MODEL.NAME, "default"))
    with pytest.raises(AttributeError):
        cfg.MODEL.NUM_CLASSES = 3                       # frozen, like yacs
    cfg.defrost()
    cfg.MODEL.NUM_CLASSES = 3
    cfg.freeze()
    assert cfg.clone().MODEL.NUM_CLASSES == 3 and "NUM_CLASSES: 3" in cfg.dump()

    class Args:
        cfg = mv.config.DEFAULT_YAML
        opts = ["MODEL.SWINV2.WINDOW_SIZE", "14", "TRAIN.BASE_LR", "1e-4"]
        batch_size = 4
        local_rank = 1
        eval = True
    c2 = mv.get_config(Args())
    assert c2.MODEL.SWINV2.WINDOW_SIZE == 14 and c2.TRAIN.BASE_LR == 1e-4 and c2.DATA.BATCH_SIZE == 4
    assert c2.LOCAL_RANK == 1 and c2.EVAL_MODE is True
    Args.opts = ["MODEL.NOPE", "1"]
    with pytest.raises(KeyError):
        mv.get_config(Args())


def test_build_model_boundary():
    cfg = mv.default_config()
    m = mv.build_model(cfg)
    assert isinstance(m, mv.SwinTransformerV2)
    assert sum(p.numel() for p in m.parameters()) == 86895866
    assert m.no_weight_decay() == {"absolute_pos_embed"}
    assert m.no_weight_decay_keywords() == {"cpb_mlp", "logit_scale", "relative_position_bias_table"}
    cfg.defrost()
    cfg.MODEL.TYPE = "resnet"
    with pytest.raises(NotImplementedError, match="Unkown model"):
        mv.build_model(cfg)
    keys = set(m.state_dict().keys())
    for k in ("patch_embed.proj.weight", "layers.0.blocks.1.attn_mask", "layers.0.blocks.0.attn.relative_position_index",
              "layers.2.blocks.17.attn.cpb_mlp.2.weight", "layers.1.downsample.reduction.weight", "norm.weight",
              "head.bias", "layers.3.blocks.0.attn.logit_scale", "layers.0.blocks.0.attn.q_bias"):
        assert k in keys, k
    assert "layers.2.blocks.1.attn_mask" not in keys                     # res <= window: no shift, no mask
    # reference init: res-post-norm LayerNorms are zero (swin_transformer_v2.py:447-452)
    assert float(m.layers[0].blocks[0].norm1.weight.abs().sum()) == 0.0


def test_swin_buffers_match_reference_golden(golden):
    g = golden["swin"]
    m = cases.make_swin("small_ws7")
    assert torch.equal(m.layers[0].blocks[0].attn.relative_position_index, g["rpi7"])
    assert torch.equal(m.layers[0].blocks[1].attn_mask, g["mask28_ws7"])
    from mvuld_b200.swin_transformer_v2 import _relative_coords_table, _relative_position_index
    assert torch.allclose(_relative_coords_table(28, 12), g["coords28"], atol=1e-6)
    assert int(_relative_position_index(28).sum()) == g["rpi28_sum"]


def test_model_sizes_and_state_dict_keys():
    f = mv.Multi_DefectModel_new_GCN(mv.default_config())
    assert sum(p.numel() for p in f.parameters()) == 19178002
    keys = set(f.state_dict().keys())
    for k in ("gat.fc.weight", "gat.attn_l", "gat2.bias", "hidden.7.weight", "Rs_GCN_8.W.1.running_var",
              "Rs_GCN_1.theta.weight", "bn_gat.running_mean", "fc_bbox.weight", "final_fc_bn.weight", "hln.weight",
              "fconly.bias", "ln_text.weight"):
        assert k in keys, k
    assert f.gat.attn_l.shape == (1, 4, 512) and f.Rs_GCN_1.g.weight.shape == (512, 512, 1)
    assert float(f.Rs_GCN_3.W[1].weight.abs().sum()) == 0.0               # Rs_GCN.py:33-34 zero init
    u = mv.build_MyUniXcoder()
    assert sum(p.numel() for p in u.encoder.parameters()) == 125929728
    assert "encoder.encoder.layer.11.attention.self.query.weight" in u.state_dict()
    gg = mv.GGNNSum(132, 200, max_edge_types=3, num_steps=8)
    assert set(gg.state_dict().keys()) >= {"ggnn.linears.2.weight", "ggnn.gru.weight_ih", "ggnn.gru.bias_hh",
                                           "classifier.weight"}


def test_graph_container_follows_dgl_semantics():
    g = G.graph((torch.tensor([0, 1, 1, 2]), torch.tensor([1, 2, 2, 2])))
    g.edata["_ETYPE"] = torch.tensor([3, 1, 1, 0])
    g.ndata["x"] = torch.arange(3.0).view(3, 1)
    g2 = G.add_self_loop(g)
    assert g2.edges()[0].tolist() == [0, 1, 1, 2, 0, 1, 2] and g2.edata["_ETYPE"].tolist() == [3, 1, 1, 0, 0, 0, 0]
    h = G.add_self_loop(G.graph((torch.tensor([0]), torch.tensor([1]))))
    h.edata["_ETYPE"] = torch.tensor([2, 0, 0])
    h.ndata["x"] = torch.zeros(2, 1)
    b = G.batch([g2, h])
    assert b.num_nodes() == 5 and b.batch_size == 2 and b.batch_num_nodes().tolist() == [3, 2]
    assert b.edges()[0].tolist()[7:] == [3, 3, 4] and b.edges()[1].tolist()[7:] == [4, 3, 4]
    # identical to the oracle's restatement on a synthetic batch
    gb = synth.cpg_batch(5, seed=4)
    hb = cases.to_host_batch(gb)
    off = dgl_ops.node_offsets(hb.batch_num_nodes)
    assert off[-1] == gb.num_nodes() and int(gb.batch_num_edges().sum()) == gb.num_edges()
    # every node has an in-edge (self loop) and self loops come last in each graph's edge list
    src, dst = gb.edges()
    assert set(dst.tolist()) == set(range(gb.num_nodes()))
    n0, e0 = int(hb.batch_num_nodes[0]), int(hb.batch_num_edges[0])
    assert src[e0 - n0:e0].tolist() == list(range(n0)) and gb.edata["_ETYPE"][e0 - n0:e0].sum() == 0
    assert set(gb.edata["_ETYPE"].tolist()) == {0, 1, 3}


def test_synthetic_inputs_are_deterministic_and_shaped():
    a, b = synth.images(2, 64, seed=5), synth.images(2, 64, seed=5)
    assert torch.equal(a, b) and a.shape == (2, 3, 64, 64)
    ids = synth.token_ids(4, 512, seed=5)
    assert ids.shape == (4, 512) and (ids[:, 0] == 0).all() and (ids[:, 1] == 6).all()
    lens = ids.ne(1).sum(1)
    for r in range(4):
        assert (ids[r, lens[r]:] == 1).all() and ids[r, lens[r] - 1] == 2      # suffix padding, closing [SEP]
    g = synth.ggnn_batch(16, seed=5)
    assert g.ndata["_WORD2VEC"].shape[1] == 132 and int(g.edata["_ETYPE"].max()) <= 3
    assert g.num_edges() == 5 * g.num_nodes()
    hb = cases.to_host_batch(g)
    off = dgl_ops.node_offsets(hb.batch_num_nodes)
    gid_src = np.searchsorted(off, hb.src, side="right") - 1
    gid_dst = np.searchsorted(off, hb.dst, side="right") - 1
    assert np.array_equal(gid_src, gid_dst)                                  # no cross-graph edges


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    header = open(os.path.join(ROOT, "include", "mvuld_b200.h")).read()
    declared = set(re.findall(r"\b(mvuld_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mvuld_b200.h but not exported"
    assert declared - {"mvuld_last_error"} == set(_lib.SIGNATURES.keys())
    lib.mvuld_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mvuld_last_error(), bytes)
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if sass:                                                                 # cuobjdump present: Blackwell-native code
        assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_product_path_fails_loudly_without_gpu():
    m = cases.make_swin("small_ws7")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            m.forward_features(torch.zeros(1, 3, 112, 112))
        f = cases.make_fusion()
        g = synth.cpg_batch(1, seed=1)
        with pytest.raises(RuntimeError, match="CUDA"):
            f(g, torch.zeros(1, 1024), torch.zeros(1, 768))
    import glob
    mods = sorted(glob.glob(os.path.join(ROOT, "mvuld_b200", "*.py")))
    assert len(mods) >= 14
    for mod in mods:                                     # every product module, including train / prefetch / checkpoint
        src = open(mod).read()
        assert "import oracle" not in src and "from oracle" not in src, f"{mod} must not use the oracle"
    f = cases.make_fusion()
    if not torch.cuda.is_available():
        from mvuld_b200 import train
        with pytest.raises(RuntimeError, match="CUDA"):
            train.FusionTrainer(f)                       # training has no CPU path either


def test_pack_lines_next_fit_keeps_order_and_capacity():
    from mvuld_b200.unixcoder import pack_lines, _lines_to_rows
    lens = [200, 200, 200, 512, 1, 511, 2, 100]
    row, off, n = pack_lines(lens, 512)
    assert row == [0, 0, 1, 2, 3, 3, 4, 4] and off == [0, 200, 0, 0, 0, 1, 0, 2] and n == 5
    used = {}
    for r, o, l in zip(row, off, lens):
        assert o == used.get(r, 0)                      # back to back, input order
        used[r] = o + l
        assert used[r] <= 512
    assert pack_lines([], 512) == ([], [], 0)
    # first-fit-decreasing: whole functions (hundreds of tokens) pack tighter than in input order
    lens2 = [260, 260, 250, 250, 400, 100, 12]
    r2, o2, n2 = pack_lines(lens2, 512, "first_fit_decreasing")
    assert n2 == 3 < pack_lines(lens2, 512, "next_fit")[2]
    spans = {}
    for r, o, l in zip(r2, o2, lens2):
        spans.setdefault(r, []).append((o, o + l))
    for r, sp in spans.items():
        sp.sort()
        assert sp[0][0] == 0 and sp[-1][1] <= 512 and all(a[1] == b[0] for a, b in zip(sp, sp[1:]))   # back to back
    assert pack_lines(lens2, 512, "auto")[2] == n2 and pack_lines([20] * 100, 512, "auto") == pack_lines([20] * 100, 512)
    with pytest.raises(ValueError):
        pack_lines([513], 512)
    ids = synth.line_token_ids(5, vocab=1000, seed=3)
    rows, ln = _lines_to_rows(ids, 1)
    assert all(len(r) == l and r[0] == 0 and r[1] == 6 and r[2] == 2 and r[-1] == 2 for r, l in zip(rows, ln))
    bad = ids.clone()
    bad[0, 2] = 1                                       # a pad token in the middle of a line
    with pytest.raises(ValueError):
        _lines_to_rows(bad, 1)


def test_load_pretrained_strips_derived_buffers_and_reinits_head(golden):
    """utils_multi.py:35-122 on a reference-layout SwinV2 checkpoint: derived buffers never overwrite the rebuilt ones,
    everything else loads, a head of another width is re-initialised to zero."""
    from mvuld_b200 import checkpoint
    kw = dict(cases.SWIN_CASES["small_ws7"])
    src = cases.make_swin("small_ws7")
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    for k in sd:
        if "relative_position_index" in k:
            sd[k] = torch.full_like(sd[k], 123)            # poison: must NOT be loaded
    dst = mv.SwinTransformerV2(**kw).eval()
    msg = checkpoint.load_pretrained(dst, {"model": sd})
    assert not msg.unexpected_keys
    assert all(any(d in k for d in ("relative_position_index", "relative_coords_table", "attn_mask"))
               for k in msg.missing_keys)
    got = dst.state_dict()
    for k, v in src.state_dict().items():
        assert torch.equal(got[k], v), k                   # parameters copied, buffers equal the re-derived ones
    kw5 = dict(kw, num_classes=5)
    dst5 = mv.SwinTransformerV2(**kw5).eval()
    checkpoint.load_pretrained(dst5, sd)
    assert float(dst5.head.weight.abs().sum()) == 0.0 and float(dst5.head.bias.abs().sum()) == 0.0
    assert torch.equal(dst5.state_dict()["layers.0.blocks.0.attn.qkv.weight"], sd["layers.0.blocks.0.attn.qkv.weight"])


def _reference_style_ablation(m, name, g, img, txt):
    """The reference forward of an ablation class written with the class's own torch modules in eval mode and the
    reference's Python unbatch loop (GraphModel.py:30-54) -- an independent path to check oracle.fusion.ablation_forward."""
    import torch.nn.functional as F
    x = F.elu(m.swinfc(m.swinbn(img)))
    t = F.elu(m.fc_text(m.bn_text(txt)))
    if name.endswith("noGraph"):
        return m.final_fc(m.final_fc_bn(torch.cat((x, t), dim=1)))
    h = F.elu(m.fconly(g.ndata["_UNIX_NODE_EMB"]))
    if name.endswith("NOGAT2"):
        for layer in m.hidden:
            h = F.elu(layer(h))
    sizes = [int(v) for v in g.batch_num_nodes()]
    if name.endswith("000"):
        hf = torch.stack([c.mean(0) for c in torch.split(h, sizes)])
        hf = F.elu(m.hfc(m.hbn(hf)))
    else:
        def slots(feat):
            out = []
            for c in torch.split(feat, sizes):
                c = c[:100]
                out.append(torch.cat([c, torch.zeros(100 - c.shape[0], c.shape[1])], 0))
            return torch.stack(out)
        z = F.elu(m.fc_gat(m.bn_gat(slots(h))))
        if hasattr(m, "fc_bbox"):
            z = torch.cat((z, F.elu(m.fc_bbox(m.bn_bbox(slots(g.ndata["pos_emb"]))))), dim=2)
        if hasattr(m, "Rs_GCN_1"):
            from oracle.fusion import rs_gcn, l2norm_dim1
            sd = m.state_dict()
            z = z.permute(0, 2, 1)
            for k in range(1, 9):
                z, _ = rs_gcn(sd, f"Rs_GCN_{k}.", z)           # pinned against the reference Rs_GCN.py (golden)
            z = l2norm_dim1(z.permute(0, 2, 1))
        hf = z.mean(dim=1)
    return m.final_fc(m.final_fc_bn(torch.cat((x, hf, t), dim=1)))


@pytest.mark.parametrize("name,width", [("Multi_DefectModel_noFunc", 1024), ("Multi_DefectModel_noGlobalImage", 512)])
def test_rq2_head_variants_surface_and_oracle(name, width):
    """new_model.py:81-319: the live graph branch with one modality dropped from the head."""
    import torch.nn.functional as F
    from oracle import fusion as ofusion
    torch.manual_seed(cases.SEED)
    m = mv.ABLATIONS[name](mv.default_config()).eval()
    synth.randomize_for_parity(m, seed=cases.SEED)
    live = set(cases.make_fusion().state_dict())
    assert set(m.state_dict()) == live                      # same modules as the live model, only the head width differs
    assert m.final_fc.weight.shape == (2, width) and m.final_fc_bn.weight.shape == (width,)
    g = synth.cpg_batch(3, seed=cases.SEED + 6)
    gen = torch.Generator().manual_seed(6)
    img, txt = torch.randn(3, 1024, generator=gen), torch.randn(3, 768, generator=gen)
    sd = m.state_dict()
    taps = {}
    got = ofusion.ablation_forward(name, sd, cases.to_host_batch(g), img, txt)
    ofusion.fusion_forward(sd, cases.to_host_batch(g), img, txt, taps=taps, head=name.split("_")[-1])
    z = taps["gcn_out"].permute(0, 2, 1)
    z = (z / z.pow(2).sum(1, keepdim=True).sqrt()).mean(1)
    with torch.no_grad():
        x, t = F.elu(m.swinfc(m.swinbn(img))), F.elu(m.fc_text(m.bn_text(txt)))
        feats = torch.cat((x, z), 1) if name.endswith("noFunc") else t * z
        ref = m.final_fc(m.final_fc_bn(feats))
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got, ref)
    with pytest.raises(RuntimeError):
        m(g, img, txt)


@pytest.mark.parametrize("name", sorted(n for n in mv.ABLATIONS if n not in ("Multi_DefectModel_noFunc",
                                                                             "Multi_DefectModel_noGlobalImage")))
def test_ablation_variants_surface_and_oracle(name):
    from oracle import fusion as ofusion
    torch.manual_seed(cases.SEED)
    m = mv.ABLATIONS[name](mv.default_config()).eval()
    synth.randomize_for_parity(m, seed=cases.SEED)
    keys = set(m.state_dict())
    want = {"fconly.weight", "hidden.7.bias", "bn_text.running_var", "ln_text.weight", "fc_text.weight", "swinbn.weight",
            "swinfc.bias", "hbn.running_mean", "hln.bias", "hfc.weight", "final_fc.weight", "final_fc_bn.running_mean"}
    spec = ofusion.VARIANT_SPECS[name]
    if spec["readout"] == "slots":
        want |= {"bn_gat.running_mean", "fc_gat.weight"}
        assert m.fc_gat.weight.shape == ((480, 512) if spec["pos"] else (512, 512))
        if spec["pos"]:
            want |= {"bn_bbox.weight", "fc_bbox.weight"}
        if spec["gcn"]:
            want |= {"Rs_GCN_1.theta.weight", "Rs_GCN_8.W.1.running_var"}
    assert want <= keys, want - keys
    assert ("Rs_GCN_1.g.weight" in keys) == bool(spec.get("gcn"))
    assert m.final_fc.weight.shape == (2, 1024 if spec["nodes"] is None else 1536)
    g = synth.cpg_batch(4, seed=cases.SEED + 5)
    gen = torch.Generator().manual_seed(5)
    img, txt = torch.randn(4, 1024, generator=gen), torch.randn(4, 768, generator=gen)
    with torch.no_grad():
        ref = _reference_style_ablation(m, name, g, img, txt)
    got = ofusion.ablation_forward(name, m.state_dict(), cases.to_host_batch(g), img, txt)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got, ref)
    with pytest.raises(RuntimeError):                     # the product path has no CPU fallback
        m(g, img, txt)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle port, no GPU): one JSON line with the base contract's keys, the
    reference arm's `impl`, `cpu_baseline` and a zero-copy `e2e`; under torchrun only rank 0 prints."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "ggnn", "--steps", "1",
           "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "graphs/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]
