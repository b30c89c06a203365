"""functions/s of the full forward at small and large per-GPU batches, eager vs CUDA-graph replay."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvuld_b200 as mv
from mvuld_b200 import synth
from mvuld_b200.graphs import GraphedMVulD

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = mv.MVulD(mv.default_config()).eval()
synth.randomize_for_parity(model, seed=777)
model = model.to(dev)
fast = GraphedMVulD(model)
out = {}
for B in (1, 4, 16, 64):
    img = synth.images(B, 448, seed=B).to(dev)
    ids = synth.token_ids(B, 512, seed=B).to(dev)
    g = synth.cpg_batch(B, seed=B)
    g.ndata.pop("_FUNC_EMB", None)
    g = g.to(dev)
    packed = model.unix.encoder.pack_host(ids.cpu()).to(dev)
    res = {}
    ref = None
    for name, fn in (("eager_padded", lambda: model(img, ids, g)), ("graph_padded", lambda: fast(img, ids, g)),
                     ("eager_packed", lambda: model(img, packed, g))):
        for _ in range(3):
            g._csr = None
            y = fn()
        torch.cuda.synchronize()
        if name == "eager_padded":
            ref = y.clone()
        elif name == "graph_padded":
            assert torch.equal(ref, y), "graph replay differs from the eager forward"
        n = max(5, 200 // B)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(n):
            g._csr = None
            y = fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        res[name] = dict(ms_per_step=round(ms, 3), functions_per_s=round(B / ms * 1e3, 1),
                         host_ms_per_step=round((time.perf_counter() - t0) / n * 1e3, 3))
    out[f"batch_{B}"] = res
    print(B, json.dumps(res), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "batch_sweep.json"), "w"), indent=1)
