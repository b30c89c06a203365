"""CPU: the oracle restatements against golden vectors produced by the reference itself (tools/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import dgl_ops, fusion, roberta, swin
from mvuld_b200 import synth
from tests import cases


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["small_ws7", "mid_ws14"])
def test_swin_oracle_matches_reference(golden, name):
    model = cases.make_swin(name)
    x = synth.images(cases.SWIN_BATCH[name], cases.SWIN_CASES[name]["img_size"], seed=cases.SEED)
    geo = cases.swin_geometry(name)
    feats = swin.forward_features(model.state_dict(), geo, x)
    ref = golden["swin"][name]["features"]
    assert torch.allclose(feats, ref, rtol=1e-4, atol=2e-5), float((feats - ref).abs().max())
    logits = swin.forward(model.state_dict(), geo, x)
    assert torch.allclose(logits, golden["swin"][name]["logits"], rtol=1e-4, atol=2e-5)


def test_swin_backward_oracle_matches_reference_autograd(golden):
    """oracle.swin.features_and_grads (the checker of the encoder-backward kernels to come, SURVEY.md 8d row 4 primary)
    against autograd through the unmodified reference module (tools/make_golden.py swin_train): every parameter
    gradient's norm and 32 sampled entries, and the image gradient."""
    g = golden["swin_train"]
    name = "small_ws7"
    model = cases.make_swin(name)
    x = synth.images(2, cases.SWIN_CASES[name]["img_size"], seed=cases.SEED + 21)
    feats, grads, dx = swin.features_and_grads(model.state_dict(), cases.swin_geometry(name), x, g["cotangent"])
    assert torch.allclose(feats, g["features"], rtol=1e-4, atol=2e-5)
    assert set(grads) == set(g["grads"]), set(grads) ^ set(g["grads"])

    def check(mine, ref, what):
        f = mine.reshape(-1)
        assert f.numel() == ref["numel"], what
        scale = max(ref["norm"] / max(ref["numel"], 1) ** 0.5, 1e-12)              # rms of the reference gradient
        assert abs(float(f.double().norm()) - ref["norm"]) <= 2e-4 * ref["norm"] + 1e-9, (what, float(f.norm()), ref["norm"])
        stride = max(1, f.numel() // 16)
        assert float((f[:16] - ref["head"]).abs().max()) <= 2e-3 * scale + 1e-7, what
        assert float((f[::stride][:16] - ref["strided"]).abs().max()) <= 2e-3 * scale + 1e-7, what

    for k, r in g["grads"].items():
        check(grads[k], r, k)
    check(dx, g["dx"], "d/d image")
    # the pieces the backward kernels will have to produce are all there and non-trivial
    for k in ("layers.0.blocks.0.attn.logit_scale", "layers.0.blocks.1.attn.cpb_mlp.2.weight", "layers.1.blocks.0.attn.q_bias",
              "layers.0.downsample.reduction.weight", "patch_embed.proj.weight"):
        assert g["grads"][k]["norm"] > 0, k


def test_roberta_backward_oracle_matches_hf_autograd(golden):
    """oracle.roberta.sentence_and_grads against autograd through the installed HF RobertaModel (make_golden.py
    roberta_train): every encoder parameter that reaches the masked-mean sentence vector."""
    g = golden["roberta_train"]
    m = cases.make_roberta()
    cfg = m.config
    ids = synth.token_ids(cases.ROBERTA_BATCH, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED + 31)
    sent, grads = roberta.sentence_and_grads(m.state_dict(), cases.roberta_geometry(cfg), ids, g["cotangent"])
    assert torch.allclose(sent, g["sent"], rtol=1e-4, atol=1e-5)
    assert set(grads) == set(g["grads"]), set(grads) ^ set(g["grads"])
    for k, ref in g["grads"].items():
        f = grads[k].reshape(-1)
        assert f.numel() == ref["numel"], k
        scale = max(ref["norm"] / max(ref["numel"], 1) ** 0.5, 1e-12)
        assert abs(float(f.double().norm()) - ref["norm"]) <= 2e-4 * ref["norm"] + 1e-9, (k, float(f.norm()), ref["norm"])
        stride = max(1, f.numel() // 16)
        assert float((f[:16] - ref["head"]).abs().max()) <= 2e-3 * scale + 1e-7, k
        assert float((f[::stride][:16] - ref["strided"]).abs().max()) <= 2e-3 * scale + 1e-7, k
    assert not any("pooler" in k for k in grads)


def test_swin_integer_artefacts_match_reference(golden):
    g = golden["swin"]
    rpi = swin.relative_position_index(28)
    assert rpi.dtype == torch.int64 and sha(rpi) == g["rpi28_sha"] and int(rpi.sum()) == g["rpi28_sum"]
    assert torch.equal(swin.relative_position_index(7), g["rpi7"])
    assert torch.allclose(swin.relative_coords_table(28, 12).view(1, 55, 55, 2), g["coords28"], atol=1e-6)
    mask = swin.shifted_window_mask(112, 112, 28, 14)
    assert sha(mask) == g["mask112_sha"] and int((mask != 0).sum()) == g["mask112_nonzero"]
    assert torch.equal(swin.shifted_window_mask(28, 28, 7, 3), g["mask28_ws7"])


def test_rs_gcn_oracle_matches_reference(golden):
    m = cases.make_rs_gcn()
    v_star, R = fusion.rs_gcn(m.state_dict(), "", cases.rs_gcn_input())
    assert torch.allclose(v_star, golden["rs_gcn"]["v_star"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(R, golden["rs_gcn"]["R"], rtol=1e-4, atol=1e-5)


def test_roberta_oracle_matches_hf(golden):
    m = cases.make_roberta()
    cfg = m.config
    ids = synth.token_ids(cases.ROBERTA_BATCH, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED)
    tok, sent = roberta.encode(m.state_dict(), cases.roberta_geometry(cfg), ids, prefix="encoder.")
    g = golden["roberta"]
    assert torch.allclose(sent, g["sent"], rtol=1e-4, atol=1e-5), float((sent - g["sent"]).abs().max())
    mask = g["mask"].unsqueeze(-1)
    assert torch.allclose(tok * mask, g["tok"] * mask, rtol=1e-3, atol=1e-4)     # valid-token rows only
    # the reference's 3-D mask and a plain key mask give the same pooled vector (pad-query rows are excluded)
    _, sent_k = roberta.encode(m.state_dict(), cases.roberta_geometry(cfg), ids, prefix="encoder.", key_mask_only=True)
    assert torch.allclose(sent, sent_k, rtol=1e-5, atol=1e-6)


def test_graph_oracle_drift_guard(golden):
    g = synth.cpg_batch(cases.FUSION_BATCH, seed=cases.SEED)
    hb = cases.to_host_batch(g)
    gg = golden["graph"]
    assert torch.equal(torch.from_numpy(hb.batch_num_nodes), gg["bnn"])
    indptr, indices, eids = dgl_ops.in_csr(hb.src, hb.dst, hb.num_nodes)
    assert sha(torch.from_numpy(indptr)) == gg["indptr_sha"]
    assert sha(torch.from_numpy(indices)) == gg["indices_sha"]
    assert sha(torch.from_numpy(eids)) == gg["eids_sha"]
    assert sha(torch.from_numpy(dgl_ops.pad_truncate_map(hb.batch_num_nodes, 100))) == gg["pad_map_sha"]
    model = cases.make_fusion()
    ge = torch.Generator().manual_seed(cases.SEED)
    img = torch.randn(cases.FUSION_BATCH, 1024, generator=ge)
    txt = torch.randn(cases.FUSION_BATCH, 768, generator=ge)
    logits = fusion.fusion_forward(model.state_dict(), hb, img, txt)
    assert torch.allclose(logits, gg["fusion_logits"], rtol=1e-4, atol=1e-5)


def test_dgl_semantics_small_cases():
    # add_self_loop: appended after existing edges, edata zero-filled, multi-edges / existing loops kept
    g = dgl_ops.graph([0, 1, 1, 2], [1, 2, 2, 2])
    g.edata["_ETYPE"] = torch.tensor([3, 1, 1, 0])
    g2 = dgl_ops.add_self_loop(g)
    assert g2.src.tolist() == [0, 1, 1, 2, 0, 1, 2] and g2.dst.tolist() == [1, 2, 2, 2, 0, 1, 2]
    assert g2.edata["_ETYPE"].tolist() == [3, 1, 1, 0, 0, 0, 0]
    # batch: node ids shifted, edges concatenated graph by graph
    h = dgl_ops.graph([0], [1])
    h.edata["_ETYPE"] = torch.tensor([2])
    b = dgl_ops.batch([g2, dgl_ops.add_self_loop(h)])
    assert b.batch_num_nodes.tolist() == [3, 2] and b.batch_num_edges.tolist() == [7, 3]
    assert b.src.tolist()[7:] == [3, 3, 4] and b.dst.tolist()[7:] == [4, 3, 4]
    # in-edge CSR sorted by (dst, eid)
    indptr, idx, eids = dgl_ops.in_csr(g2.src, g2.dst, 3)
    assert indptr.tolist() == [0, 1, 3, 7] and eids.tolist() == [4, 0, 5, 1, 2, 3, 6]
    # pad / truncate map: short graph padded with -1, long graph truncated
    pm = dgl_ops.pad_truncate_map(np.array([2, 5, 0]), 3)
    assert pm.tolist() == [[0, 1, -1], [2, 3, 4], [-1, -1, -1]]
    # segment sum / mean incl. an empty graph
    feat = torch.arange(14, dtype=torch.float32).view(7, 2)
    assert dgl_ops.segment_sum(feat, np.array([2, 5, 0])).tolist() == [[2., 4.], [40., 45.], [0., 0.]]
    assert dgl_ops.mean_nodes(feat, np.array([2, 5, 0]))[2].tolist() == [0., 0.]
    # GATConv raises on a zero-in-degree node
    sd = {"fc.weight": torch.randn(8, 2), "attn_l": torch.randn(1, 2, 4), "attn_r": torch.randn(1, 2, 4),
          "bias": torch.zeros(8)}
    with pytest.raises(RuntimeError):
        dgl_ops.gat_conv(sd, "", [0], [1], torch.randn(2, 2), 2, 4)
    # GatedGraphConv asserts the edge-type range
    with pytest.raises(AssertionError):
        dgl_ops.gated_graph_conv({}, "", [0], [1], [5], torch.randn(2, 2), 4, 1, 3)


def test_gat_softmax_is_per_destination():
    torch.manual_seed(0)
    sd = {"fc.weight": torch.randn(8, 3), "attn_l": torch.randn(1, 2, 4), "attn_r": torch.randn(1, 2, 4),
          "bias": torch.zeros(8)}
    x = torch.randn(3, 3)
    # node 2 has in-edges from 0, 1 and itself; a single in-edge node gets exactly z[src]
    out = dgl_ops.gat_conv(sd, "", [0, 1, 2, 0, 1], [2, 2, 2, 0, 1], x, 2, 4)
    z = (x @ sd["fc.weight"].T).view(3, 2, 4)
    assert torch.allclose(out[0], z[0], atol=1e-6) and torch.allclose(out[1], z[1], atol=1e-6)


FUSION_KEYS = ["Multi_DefectModel_new_GCN", "Multi_DefectModel", "myModels.Multi_DefectModel", "Multi_DefectModel_noGraph",
               "Multi_DefectModel_000", "Multi_DefectModel_001", "Multi_DefectModel_100", "Multi_DefectModel_NOGAT2",
               "Multi_DefectModel_noFunc", "Multi_DefectModel_noGlobalImage", "Multi_DefectModel_110",
               "Multi_DefectModel_GATPOS", "Multi_DefectModel_011", "Multi_DefectModel_NOGAT", "Multi_DefectModel_NOGAT3",
               "Multi_DefectModel_NOGAT4"]


def test_fusion_class_golden_covers_every_class(golden):
    assert sorted(golden["fusion_classes"]) == sorted(FUSION_KEYS) == sorted(k for k, *_ in cases.fusion_class_cases())


@pytest.mark.parametrize("key", FUSION_KEYS)
def test_fusion_class_oracle_matches_reference_class(golden, key):
    """oracle.fusion.class_forward against the logits of the reference's OWN class (GraphModel.py / new_model.py /
    myModels.py loaded unmodified from /root/reference by tools/make_golden.py fusion_classes, `dgl` stubbed with the
    restated GATConv / unbatch / mean_nodes); the mirror class's state-dict keys are the ones the reference class
    loaded with strict=True."""
    ref = golden["fusion_classes"][key]
    mirror = {k: c for k, _, _, c in cases.fusion_class_cases()}[key]
    m = cases.fusion_class_model(key, mirror)
    assert sorted(m.state_dict().keys()) == ref["keys"]
    g, img, txt = cases.fusion_class_inputs()
    got = fusion.class_forward(key, m.state_dict(), cases.to_host_batch(g), img, txt)
    assert torch.allclose(got, ref["logits"], rtol=1e-4, atol=1e-4), (got, ref["logits"])


def test_gru_restatement_matches_torch_gru():
    """oracle.fusion.gru_last_state (checker of mvuld_gru_sequence) against torch.nn.GRU."""
    torch.manual_seed(3)
    gru = torch.nn.GRU(64, 64, 1, batch_first=True)
    x = torch.randn(3, 37, 64)
    with torch.no_grad():
        _, hn = gru(x)
    got = fusion.gru_last_state({"g." + k: v for k, v in gru.state_dict().items()}, "g.", x)
    assert torch.allclose(got, hn[0], rtol=1e-5, atol=1e-5)


# ---- second, independently written DGL restatement (oracle/dgl_ops_alt.py) against the first ----
def _small_multigraph(seed):
    """A small graph with parallel edges, self loops appended last and every node with an in-edge."""
    g = torch.Generator().manual_seed(seed)
    n = 23
    src = torch.randint(0, n, (70,), generator=g)
    dst = torch.randint(0, n, (70,), generator=g)
    src = torch.cat([src, src[:9], torch.arange(n)])          # 9 duplicated (parallel) edges + the self loops
    dst = torch.cat([dst, dst[:9], torch.arange(n)])
    return src.numpy().astype(np.int64), dst.numpy().astype(np.int64), n


def test_second_dgl_restatement_gatconv_dense_adjacency():
    from oracle import dgl_ops_alt
    src, dst, n = _small_multigraph(1)
    torch.manual_seed(2)
    H, Fo, Fi = 4, 24, 40
    sd = {"fc.weight": torch.randn(H * Fo, Fi) * 0.2, "attn_l": torch.randn(1, H, Fo), "attn_r": torch.randn(1, H, Fo),
          "bias": torch.randn(H * Fo)}
    x = torch.randn(n, Fi)
    a = dgl_ops.gat_conv(sd, "", src, dst, x, H, Fo)
    b = dgl_ops_alt.gat_conv_dense(sd, "", src, dst, x, H, Fo)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), float((a - b).abs().max())
    with pytest.raises(RuntimeError):                                      # a node without in-edges: both raise
        dgl_ops_alt.gat_conv_dense(sd, "", src[:20], dst[:20], x, H, Fo)
    with pytest.raises(RuntimeError):
        dgl_ops.gat_conv(sd, "", src[:20], dst[:20], x, H, Fo)


def test_second_dgl_restatement_gated_graph_conv_edge_loop():
    from oracle import dgl_ops_alt
    src, dst, n = _small_multigraph(3)
    torch.manual_seed(4)
    D, T, steps, Fi = 16, 3, 4, 10
    et = torch.randint(0, T, (len(src),))
    sd = {f"linears.{t}.weight": torch.randn(D, D) * 0.3 for t in range(T)}
    sd.update({f"linears.{t}.bias": torch.randn(D) * 0.1 for t in range(T)})
    cell = torch.nn.GRUCell(D, D)
    sd.update({"gru.weight_ih": cell.weight_ih.detach(), "gru.weight_hh": cell.weight_hh.detach(),
               "gru.bias_ih": cell.bias_ih.detach(), "gru.bias_hh": cell.bias_hh.detach()})
    x = torch.randn(n, Fi)
    a = dgl_ops.gated_graph_conv(sd, "", src, dst, et, x, D, steps, T)
    b = dgl_ops_alt.gated_graph_conv_loop(sd, "", src, dst, et, x, D, steps, T)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())


def test_second_dgl_restatement_index_artefacts_bit_exact():
    from oracle import dgl_ops_alt
    rng = np.random.default_rng(5)
    raw = []
    for n in (1, 7, 130, 2):                                                # one-node graph, one above the 100-slot cut
        e = int(rng.integers(0, 3 * n + 1))
        raw.append((rng.integers(0, n, e), rng.integers(0, n, e), n, torch.from_numpy(rng.integers(0, 4, e))))
    graphs = []
    for s, d, n, et in raw:
        hg = dgl_ops.graph(s, d, n)
        hg.edata["_ETYPE"] = et
        hg.ndata["f"] = torch.from_numpy(rng.standard_normal((n, 6)).astype(np.float32))
        graphs.append(dgl_ops.add_self_loop(hg))
    hb = dgl_ops.batch(graphs)
    S, D, T, bnn, bne = dgl_ops_alt.batch_with_self_loops_loop(raw)
    assert np.array_equal(hb.src, S) and np.array_equal(hb.dst, D) and np.array_equal(hb.edata["_ETYPE"].numpy(), T)
    assert np.array_equal(hb.batch_num_nodes, bnn) and np.array_equal(hb.batch_num_edges, bne)
    ip, idx, eid = dgl_ops.in_csr(hb.src, hb.dst, hb.num_nodes)
    ip2, idx2, eid2 = dgl_ops_alt.in_edges_loop(S, D, hb.num_nodes)
    assert np.array_equal(ip, ip2) and np.array_equal(idx, idx2) and np.array_equal(eid, eid2)
    a = dgl_ops.unbatch_pad(hb.ndata["f"], hb.batch_num_nodes, 100)
    assert torch.equal(a, dgl_ops_alt.unbatch_pad_loop(hb.ndata["f"], bnn, 100))
