"""Standalone GGNN forward (configs[2]: 4096 graphs) for timing / ncu captures of the segment-reduce kernels."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvuld_b200 as mv
from mvuld_b200 import synth
B = int(os.environ.get("PB", 4096))
torch.manual_seed(12345)
model = mv.GGNNSum(132, 200, max_edge_types=4, num_steps=6).eval()
synth.randomize_for_parity(model, seed=777)
model = model.cuda()
g = synth.ggnn_batch(B, seed=12345, n_etypes=4).to("cuda")
for _ in range(2):
    out = model(g)[1]
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
out = model(g)[1]
e.record()
torch.cuda.synchronize()
print("ggnn forward ms", s.elapsed_time(e), "nodes", g.num_nodes(), "edges", g.num_edges())
