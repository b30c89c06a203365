"""GPU: model-level parity of the B200 path against the CPU oracle and the reference-generated golden vectors.

Tolerances are north_star's: integer artefacts bit-exact; bf16 logits within 1e-2 relative of the fp32 reference;
identical argmax.  Intermediate feature vectors are compared by relative L2 error.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mvuld_b200 as mv                     # noqa: E402
from mvuld_b200 import synth                # noqa: E402
from oracle import fusion as ofusion, roberta as oroberta, swin as oswin   # noqa: E402
from tests import cases                     # noqa: E402

DEV = "cuda"


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def logits_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))


def logits_close(a, b, tol=1e-2):
    return logits_err(a, b) < tol


TOL = 1e-2          # north_star: bf16 results within 1e-2 relative of the fp32 reference


def check(name, measured, tol=TOL):
    """Assert against north_star's tolerance and record the measured error (gpurun_out/parity_errors.json)."""
    from tests.conftest import record_parity
    record_parity(name, measured, tol)
    assert measured < tol, (name, measured, tol)


@pytest.mark.parametrize("name", ["small_ws7", "mid_ws14", "full"])
def test_swin_matches_reference_golden(golden, name):
    model = cases.make_swin(name)
    x = synth.images(cases.SWIN_BATCH[name], cases.SWIN_CASES[name]["img_size"], seed=cases.SEED)
    model = model.to(DEV)
    feats = model.forward_features(x.to(DEV))
    logits = model(x.to(DEV))
    torch.cuda.synchronize()
    ref = golden["swin"][name]
    check(f"swin[{name}].features rel-L2 vs reference-module golden", rel_err(feats, ref["features"]))
    check(f"swin[{name}].logits max-rel vs reference-module golden", logits_err(logits, ref["logits"]))
    assert torch.equal(logits.cpu().argmax(1), ref["logits"].argmax(1))


def test_swin_block_taps_against_oracle():
    """Per-block relative error against the fp32 oracle: localises drift instead of only testing the end."""
    name = "small_ws7"
    model = cases.make_swin(name)
    x = synth.images(cases.SWIN_BATCH[name], cases.SWIN_CASES[name]["img_size"], seed=cases.SEED)
    taps = {}
    ref = oswin.forward_features(model.state_dict(), cases.swin_geometry(name), x, taps=taps)
    feats = model.to(DEV).forward_features(x.to(DEV))
    torch.cuda.synchronize()
    check("swin[small_ws7].features rel-L2 vs oracle", rel_err(feats, ref))


def test_swin_boundary_errors():
    cfg = mv.default_config()
    m = mv.build_model(cfg)
    assert m.output_num() == 1024 and abs(m.flops() / 1e9 - 79.57) < 0.01
    m = cases.make_swin("small_ws7").to(DEV)
    with pytest.raises(AssertionError):
        m.forward_features(torch.zeros(1, 3, 64, 64, device=DEV))
    m.train()
    with pytest.raises(RuntimeError):
        m.forward_features(torch.zeros(1, 3, 112, 112, device=DEV))


@pytest.mark.parametrize("full", [False, True])
def test_unixcoder_matches_oracle(golden, full):
    m = cases.make_roberta(full=full)
    cfg = m.config
    B = cases.ROBERTA_BATCH
    ids = synth.token_ids(B, cases.ROBERTA_L, cfg.vocab_size, seed=cases.SEED)
    tok_ref, sent_ref = oroberta.encode(m.state_dict(), cases.roberta_geometry(cfg), ids, prefix="encoder.")
    md = m.to(DEV)
    vec, _ = md.get_repr(ids.to(DEV))
    tok, _ = md.get_xcode_vec(ids.to(DEV))
    torch.cuda.synchronize()
    tag = "full" if full else "small"
    check(f"unixcoder[{tag}].sentence rel-L2 vs oracle", rel_err(vec, sent_ref))
    if not full:
        check("unixcoder[small].sentence rel-L2 vs HF golden", rel_err(vec, golden["roberta"]["sent"]))
    mask = ids.ne(cfg.pad_token_id).unsqueeze(-1).float()
    check(f"unixcoder[{tag}].tokens rel-L2 vs oracle", rel_err(tok.cpu() * mask, tok_ref * mask))
    assert torch.isfinite(tok).all()


def test_fusion_matches_oracle(golden):
    model = cases.make_fusion()
    g = synth.cpg_batch(cases.FUSION_BATCH, seed=cases.SEED)
    ge = torch.Generator().manual_seed(cases.SEED)
    img = torch.randn(cases.FUSION_BATCH, 1024, generator=ge)
    txt = torch.randn(cases.FUSION_BATCH, 768, generator=ge)
    ref = golden["graph"]["fusion_logits"]
    logits = model.to(DEV)(g.to(DEV), img.to(DEV), txt.to(DEV))
    torch.cuda.synchronize()
    check("fusion.logits max-rel vs oracle golden", logits_err(logits, ref))
    assert torch.equal(logits.cpu().argmax(1), ref.argmax(1))


def test_gat_ablation_variant_matches_oracle():
    """``Multi_DefectModel`` (GraphModel.py:214-304), the dgl.mean_nodes readout variant."""
    torch.manual_seed(cases.SEED)
    m = mv.Multi_DefectModel(mv.default_config()).eval()
    cases.round_matrices_to_bf16(synth.randomize_for_parity(m, seed=cases.SEED))
    g = synth.cpg_batch(5, seed=cases.SEED + 11)
    gen = torch.Generator().manual_seed(2)
    img, txt = torch.randn(5, 1024, generator=gen), torch.randn(5, 768, generator=gen)
    ref = ofusion.gat_variant_forward(m.state_dict(), cases.to_host_batch(g), img, txt)
    out = m.to(DEV)(g.to(DEV), img.to(DEV), txt.to(DEV))
    assert logits_close(out, ref, 1e-2), (out.cpu(), ref)
    assert torch.equal(out.cpu().argmax(1), ref.argmax(1))
    keys = set(m.state_dict())
    assert {"gat.attn_l", "gat2.fc.weight", "hbn.running_mean", "hfc.weight", "final_fc_bn.weight", "fconly.bias"} <= keys


@pytest.mark.parametrize("name", ["Multi_DefectModel_noGraph", "Multi_DefectModel_000", "Multi_DefectModel_001",
                                  "Multi_DefectModel_100", "Multi_DefectModel_NOGAT2", "Multi_DefectModel_noFunc",
                                  "Multi_DefectModel_noGlobalImage"])
def test_gat_free_ablation_variants_match_oracle(name):
    """the GATConv-free RQ2 / RQ3 classes (GraphModel.py:306-615, 1277-1384) and the two RQ2
    head variants of the live graph branch (new_model.py:81-319)."""
    torch.manual_seed(cases.SEED)
    m = mv.ABLATIONS[name](mv.default_config()).eval()
    cases.round_matrices_to_bf16(synth.randomize_for_parity(m, seed=cases.SEED))
    B = 5
    g = synth.cpg_batch(B, seed=cases.SEED + 13)           # graph sizes on both sides of the 100-slot truncation
    gen = torch.Generator().manual_seed(3)
    img, txt = torch.randn(B, 1024, generator=gen), torch.randn(B, 768, generator=gen)
    ref = ofusion.ablation_forward(name, m.state_dict(), cases.to_host_batch(g), img, txt)
    out = m.to(DEV)(g.to(DEV), img.to(DEV), txt.to(DEV))
    torch.cuda.synchronize()
    assert out.shape == (B, 2) and torch.isfinite(out).all()
    assert logits_close(out, ref, 1e-2), (name, out.cpu(), ref)
    assert torch.equal(out.cpu().argmax(1), ref.argmax(1))


@pytest.mark.parametrize("key", ["Multi_DefectModel_110", "Multi_DefectModel_GATPOS", "Multi_DefectModel_011",
                                 "Multi_DefectModel_NOGAT", "Multi_DefectModel_NOGAT3", "Multi_DefectModel_NOGAT4",
                                 "myModels.Multi_DefectModel"])
def test_grid_and_gating_fusion_classes_match_oracle_and_reference_class(golden, key):
    """SURVEY.md section 8f.3, the remaining classes: the RQ3 grid of GraphModel.py:618-1273 and the GRU-projection /
    gating-attention class of myModels.py:280-428 -- CUDA mirror vs the oracle AND vs the logits of the reference's own
    class (tests/golden/fusion_classes.pt)."""
    mirror = {k: c for k, _, _, c in cases.fusion_class_cases()}[key]
    m = cases.fusion_class_model(key, mirror)
    g, img, txt = cases.fusion_class_inputs()
    ref = ofusion.class_forward(key, m.state_dict(), cases.to_host_batch(g), img, txt)
    gold = golden["fusion_classes"][key]["logits"]
    out = m.to(DEV)(g.to(DEV), img.to(DEV), txt.to(DEV))
    torch.cuda.synchronize()
    assert out.shape == gold.shape and torch.isfinite(out).all()
    check(f"{key} logits max-rel vs oracle", logits_err(out, ref))
    check(f"{key} logits max-rel vs the reference class", logits_err(out, gold))
    assert torch.equal(out.cpu().argmax(1), gold.argmax(1))


def test_gru_sequence_kernel_matches_gru_oracle():
    """mvuld_gru_sequence (myModels.py:324,385-387) vs oracle.fusion.gru_last_state (itself pinned against torch.nn.GRU
    on the CPU): 64 rows, 333 steps, hidden 512, fp32."""
    from mvuld_b200 import _lib
    torch.manual_seed(5)
    B, T, H = 64, 333, 512
    gru = torch.nn.GRU(H, H, 1, batch_first=True)
    x = torch.randn(B, T, H) * 0.5
    ref = ofusion.gru_last_state({"g." + k: v for k, v in gru.state_dict().items()}, "g.", x)
    gi = (x.reshape(B * T, H) @ gru.weight_ih_l0.detach().t() + gru.bias_ih_l0.detach()).contiguous().to(DEV)
    out = torch.empty(B, H, device=DEV)
    ws = torch.empty(int(_lib.load().mvuld_gru_sequence_workspace(B, H)), device=DEV, dtype=torch.uint8)
    _lib.call("mvuld_gru_sequence", gi, gru.weight_hh_l0.detach().to(DEV).contiguous(),
              gru.bias_hh_l0.detach().to(DEV).contiguous(), out, ws, B, T, H)
    torch.cuda.synchronize()
    check("gru_sequence max abs error vs oracle (|h| <= 1)", float((out.cpu() - ref).abs().max()), 1e-4)


def test_fusion_rejects_zero_in_degree():
    model = cases.make_fusion().to(DEV)
    g = mv.graph.graph((torch.tensor([0, 1]), torch.tensor([1, 2])), num_nodes=3)      # node 0 has no in-edge
    g.ndata["_UNIX_NODE_EMB"] = torch.randn(3, 768)
    g.ndata["pos_emb"] = torch.zeros(3, 4)
    with pytest.raises(RuntimeError):
        model(g.to(DEV), torch.randn(1, 1024, device=DEV), torch.randn(1, 768, device=DEV))


def test_ggnn_matches_oracle(golden):
    gm = cases.make_ggnn()
    g = synth.ggnn_batch(cases.GGNN_BATCH, seed=cases.SEED, n_etypes=cases.GGNN_T)
    hb = cases.to_host_batch(g)
    prob_r, logit_r, sum_r, h_r = ofusion.ggnn_sum_forward(gm.state_dict(), hb, cases.GGNN_D, cases.GGNN_STEPS,
                                                           cases.GGNN_T)
    assert torch.allclose(sum_r, golden["graph"]["ggnn_sum"], rtol=1e-4, atol=1e-3)
    gd = g.to(DEV)
    md = gm.to(DEV)
    h = md.node_states(gd)
    prob, logit = md(gd)
    torch.cuda.synchronize()
    check("ggnn.node_states rel-L2 vs oracle", rel_err(h, h_r))
    check("ggnn.graph_sums rel-L2 vs oracle", rel_err(md._last_sum, sum_r))
    check("ggnn.logits max-rel vs oracle", logits_err(logit, logit_r))
    # edge types outside [0, n_etypes) are rejected like DGL's assert
    gd.edata["_ETYPE"] = gd.edata["_ETYPE"] + 10
    with pytest.raises(AssertionError):
        md.node_states(gd)


def test_composed_forward_matches_oracle():
    """Full MVulD forward at small Swin / small RoBERTa geometry: image + tokens + CPG -> logits."""
    torch.manual_seed(cases.SEED)
    cfg = mv.default_config()
    cfg.defrost()
    cfg.DATA.IMG_SIZE = 224
    cfg.MODEL.SWINV2.DEPTHS = [2, 2, 2, 2]
    cfg.MODEL.SWINV2.NUM_HEADS = [4, 8, 16, 32]
    cfg.MODEL.SWINV2.WINDOW_SIZE = 7
    cfg.MODEL.SWINV2.PRETRAINED_WINDOW_SIZES = [6, 6, 6, 6]
    cfg.freeze()
    rcfg = mv.roberta_base_config(vocab_size=1000, num_hidden_layers=2)
    model = mv.MVulD(cfg, rcfg).eval()
    cases.round_matrices_to_bf16(synth.randomize_for_parity(model, seed=cases.SEED))
    B = 3
    img = synth.images(B, 224, seed=1)
    ids = synth.token_ids(B, 512, 1000, seed=1)
    g = synth.cpg_batch(B, seed=1)
    # oracle, piece by piece
    from oracle.swin import SwinGeometry
    geo = SwinGeometry(img_size=224, embed_dim=128, depths=(2, 2, 2, 2), num_heads=(4, 8, 16, 32), window_size=7,
                       pretrained_window_sizes=(6, 6, 6, 6))
    sd_swin = model.swin.state_dict()
    f_img = oswin.forward_features(sd_swin, geo, img)
    f_txt = oroberta.get_repr(model.unix.state_dict(), cases.roberta_geometry(rcfg), ids)
    ref = ofusion.fusion_forward(model.fusion.state_dict(), cases.to_host_batch(g), f_img, f_txt)
    out = model.to(DEV)(img.to(DEV), ids.to(DEV), g.to(DEV))
    torch.cuda.synchronize()
    check("composed[small].logits max-rel vs oracle", logits_err(out, ref))
    assert torch.equal(out.cpu().argmax(1), ref.argmax(1))


def test_graph_branch_on_a_side_stream_gives_identical_logits():
    """MVulD.overlap_graph_branch: graph_features on a second stream + head == the single-stream forward."""
    torch.manual_seed(cases.SEED)
    model = mv.MVulD(mv.default_config(), cases.roberta_small_config()).eval()
    synth.randomize_for_parity(model, seed=777)
    model = model.to(DEV)
    B = 3
    img, ids = synth.images(B, 448, seed=5).to(DEV), synth.token_ids(B, 512, vocab=1000, seed=5).to(DEV)
    g = synth.cpg_batch(B, seed=5).to(DEV)
    want = model(img, ids, g)
    model.overlap_graph_branch = True
    for _ in range(3):
        got = model(img, ids, g)
        assert torch.equal(got, want)


def test_device_prefetcher_yields_identical_batches_and_results():
    """mvuld_b200.prefetch: batches staged on the copy stream give the same logits as synchronous staging."""
    from mvuld_b200.prefetch import DevicePrefetcher, ResultSink
    fus = cases.make_fusion().to(DEV)
    batches = []
    for i in range(3):
        g = synth.cpg_batch(2, seed=cases.SEED + i)
        gen = torch.Generator().manual_seed(i)
        for k in list(g.ndata):
            g.ndata[k] = g.ndata[k].pin_memory()
        batches.append(dict(g=g, img=torch.randn(2, 1024, generator=gen).pin_memory(),
                            txt=torch.randn(2, 768, generator=gen).pin_memory()))
    want = [fus(b["g"].to(DEV), b["img"].to(DEV), b["txt"].to(DEV)).cpu() for b in batches]
    sink = ResultSink(3)
    for d in DevicePrefetcher(batches, DEV):
        sink.push(fus(d["g"], d["img"], d["txt"]))
    got = sink.results()
    assert len(got) == 3 and all(torch.equal(a, b) for a, b in zip(got, want))
    with pytest.raises(RuntimeError):
        DevicePrefetcher(batches, "cpu")


def test_packed_line_encoding_matches_padded_reference_semantics():
    """per-node line vectors (data_list.py:292-299 -> unixcoder.py:56-68).  The packed,
    block-diagonal run must give what the reference computes with every line padded to 512 tokens: checked against the
    fp32 oracle on the padded batch (1e-2, bf16) and against this repo's own padded GPU path."""
    m = cases.make_roberta()
    cfg = m.config
    ids = synth.line_token_ids(37, vocab=cfg.vocab_size, seed=cases.SEED)
    g = torch.Generator().manual_seed(5)                                   # a few long lines, one full 512-token line
    for row, n in ((3, 300), (11, 512), (20, 131)):
        ids[row] = 1
        ids[row, :n] = torch.cat([torch.tensor([0, 6, 2]), torch.randint(4, cfg.vocab_size, (n - 4,), generator=g),
                                  torch.tensor([2])])
    ref = oroberta.get_repr(m.state_dict(), cases.roberta_geometry(cfg), ids)
    m = m.to(DEV)
    packed = m.myEncode_ids(ids)
    padded, _ = m.get_repr(ids.to(DEV))
    torch.cuda.synchronize()
    assert packed.shape == (37, cfg.hidden_size)
    assert rel_err(packed, ref) < 1e-2, rel_err(packed, ref)
    assert rel_err(packed, padded) < 5e-3, rel_err(packed, padded)
    # list-of-lists input, small pass size (several passes) -> same vectors
    as_lists = [r[r != 1].tolist() for r in ids]
    again = m.encoder.encode_lines(as_lists, rows_per_pass=1)
    assert rel_err(again, packed) < 2e-3
    with pytest.raises(RuntimeError):
        m.myEncode_ids(ids.to(DEV))                                        # packing happens on the host
    with pytest.raises(RuntimeError):
        m.myEncode(["int a = 0;"])                                         # no tokenizer offline


def test_deferred_validity_checks_raise_at_the_sync_point():
    """defer_checks postpones the per-call read-back of the kernels' input-validity flags (pipelined callers)."""
    from mvuld_b200 import graph as G
    fus = cases.make_fusion().to(DEV)
    g = synth.cpg_batch(2, seed=cases.SEED)
    gen = torch.Generator().manual_seed(1)
    img, txt = torch.randn(2, 1024, generator=gen).to(DEV), torch.randn(2, 768, generator=gen).to(DEV)
    want = fus(g.to(DEV), img, txt)
    fus.defer_checks = True
    got = fus(g.to(DEV), img, txt)
    fus.raise_if_invalid()                                     # nothing pending is wrong
    assert torch.equal(got, want)
    # a graph whose node 2 has no in-edge (no self loops): the immediate mode raises in forward, the deferred one later
    bad = G.graph((torch.tensor([0, 1]), torch.tensor([1, 0])), num_nodes=3)
    bad.ndata["_UNIX_NODE_EMB"] = torch.randn(3, 768, generator=gen)
    bad.ndata["pos_emb"] = torch.zeros(3, 4)
    bad = G.batch([bad])
    fus(bad.to(DEV), img[:1], txt[:1])                         # deferred: returns
    with pytest.raises(RuntimeError, match="0-in-degree"):
        fus.raise_if_invalid()
    fus.defer_checks = False
    with pytest.raises(RuntimeError, match="0-in-degree"):
        fus(bad.to(DEV), img[:1], txt[:1])


def test_packed_text_branch_gives_the_padded_result_in_the_composed_forward():
    """get_repr / MVulD.forward with the batch packed at data-loading time (pad tokens dropped, first-fit-decreasing)
    == the same batch as padded [B, 512] rows; a CPU id tensor is packed on the fly."""
    m = cases.make_roberta().to(DEV)
    ids = synth.token_ids(12, 512, vocab=m.config.vocab_size, seed=cases.SEED + 3)
    packed_host = m.encoder.pack_host(ids)
    assert packed_host.n_rows < 12 and not packed_host.passes[0]["ids"].is_cuda
    with pytest.raises(RuntimeError):
        m.get_repr(packed_host)                                       # must be moved to the device first
    padded, _ = m.get_repr(ids.to(DEV))
    packed, _ = m.get_repr(packed_host.to(DEV))
    from_cpu, _ = m.get_repr(ids)
    assert rel_err(packed, padded) < 5e-3 and rel_err(from_cpu, padded) < 5e-3
    # through the prefetcher (PackedLines is a batch element like a tensor or a Graph)
    from mvuld_b200.prefetch import DevicePrefetcher
    got = [m.get_repr(d["ids"])[0] for d in DevicePrefetcher([dict(ids=packed_host)] * 2, DEV)]
    assert torch.equal(got[0], packed) and torch.equal(got[1], packed)


def test_edge_cases_ragged_and_extreme_sizes():
    """Ragged / extreme inputs: a 2-node and a 1 500-node CPG in one batch (pad to 100 slots vs truncate,
    GraphModel.py:30-54), batch size 1, a single GGNN graph, one / zero code lines, a full 512-token line."""
    from mvuld_b200 import graph as G
    fus = cases.make_fusion()
    gen = torch.Generator().manual_seed(3)

    def cpg(n):
        src = torch.cat([torch.arange(0, n - 1), torch.randint(0, n, (n,), generator=gen)])
        dst = torch.cat([torch.arange(1, n), torch.randint(0, n, (n,), generator=gen)])
        g = G.graph((src, dst), num_nodes=n)
        g.ndata["_UNIX_NODE_EMB"] = torch.randn(n, 768, generator=gen) * 0.5
        g.ndata["pos_emb"] = torch.rand(n, 4, generator=gen)
        return G.add_self_loop(g)

    batch = G.batch([cpg(2), cpg(1500), cpg(100), cpg(101)])
    img, txt = torch.randn(4, 1024, generator=gen), torch.randn(4, 768, generator=gen)
    ref = ofusion.fusion_forward(fus.state_dict(), cases.to_host_batch(batch), img, txt)
    fus = fus.to(DEV)
    out = fus(batch.to(DEV), img.to(DEV), txt.to(DEV))
    assert logits_close(out, ref, 1e-2) and torch.equal(out.cpu().argmax(1), ref.argmax(1))
    one = G.batch([cpg(37)])
    ref1 = ofusion.fusion_forward({k: v.cpu() for k, v in fus.state_dict().items()}, cases.to_host_batch(one), img[:1], txt[:1])
    out1 = fus(one.to(DEV), img[:1].to(DEV), txt[:1].to(DEV))
    assert logits_close(out1, ref1, 1e-2)
    # GGNN: one graph
    gm = cases.make_ggnn()
    g1 = synth.ggnn_batch(1, seed=9, n_etypes=cases.GGNN_T)
    pref, lref, _, _ = ofusion.ggnn_sum_forward(gm.state_dict(), cases.to_host_batch(g1), cases.GGNN_D, cases.GGNN_STEPS,
                                                cases.GGNN_T)
    prob, logit = gm.to(DEV)(g1.to(DEV))
    assert logits_close(logit, lref, 1e-2)
    # line encoding: one line, a full 512-token line, no lines
    m = cases.make_roberta().to(DEV)
    one_line = synth.line_token_ids(1, vocab=m.config.vocab_size, seed=1)
    full_line = torch.cat([torch.tensor([0, 6, 2]), torch.randint(4, m.config.vocab_size, (508,), generator=gen),
                           torch.tensor([2])]).view(1, 512)
    for ids in (one_line, full_line):
        want, _ = m.get_repr(ids.to(DEV))
        got = m.myEncode_ids(ids)
        assert got.shape == (1, m.config.hidden_size) and rel_err(got, want) < 5e-3
    assert m.myEncode_ids(torch.ones(0, 512, dtype=torch.int64)).shape == (0, m.config.hidden_size)
    with pytest.raises(ValueError):
        m.myEncode_ids(torch.ones(1, 512, dtype=torch.int64))              # a line of pad tokens only
    # Swin: batch of one
    sw = cases.make_swin("small_ws7")
    x = synth.images(1, cases.SWIN_CASES["small_ws7"]["img_size"], seed=2)
    r = oswin.forward_features(sw.state_dict(), cases.swin_geometry("small_ws7"), x)
    check("swin[small_ws7, batch 1].features rel-L2 vs oracle", rel_err(sw.to(DEV).forward_features(x.to(DEV)), r))


def test_swin_sub_batches_on_streams_give_identical_features():
    """forward_features with ``streams = 2``: sub-batches on two CUDA streams, same result as one batch."""
    name = "small_ws7"
    model = cases.make_swin(name).to(DEV)
    x = synth.images(16, cases.SWIN_CASES[name]["img_size"], seed=cases.SEED + 3).to(DEV)
    one = model.forward_features(x).clone()
    model.streams = 2
    two = model.forward_features(x).clone()
    model.streams = 0
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    again = model.forward_features(x)
    assert torch.equal(one, again)


@pytest.mark.parametrize("B", [4, 16])
def test_cuda_graph_replay_with_concurrent_branches_is_bit_identical(B):
    """mvuld_b200.graphs.GraphedMVulD: the image and padded-text branches replayed from CUDA graphs on two side streams
    while the graph branch runs eagerly -- same launches, so the logits equal the single-stream eager forward bit for
    bit, on every replay.  (At 16 functions the SwinV2 cluster kernels run next to the text encoder's kernels: the case
    that exposed an early release of the residual ring in csrc/gemm_ln_cluster.cu.)"""
    from mvuld_b200.graphs import GraphedMVulD
    torch.manual_seed(cases.SEED)
    model = mv.MVulD(mv.default_config()).eval()
    synth.randomize_for_parity(model, seed=777)
    model = model.to(DEV)
    img, ids = synth.images(B, 448, seed=B).to(DEV), synth.token_ids(B, 512, seed=B).to(DEV)
    g = synth.cpg_batch(B, seed=B)
    g.ndata.pop("_FUNC_EMB", None)
    g = g.to(DEV)
    want = model(img, ids, g).clone()
    fast = GraphedMVulD(model, concurrent_below=64)
    for _ in range(6):
        assert torch.equal(fast(img, ids, g), want)
    seq = GraphedMVulD(model, concurrent_below=0)
    assert torch.equal(seq(img, ids, g), want)
