"""Standalone launch of the Rs_GCN affinity kernel (64 graphs) for timing / ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
B, n, C = int(os.environ.get("PB", 64)), 100, 512
g = torch.Generator().manual_seed(0)
tpg = (torch.randn(B * n, 3 * C, generator=g) * 0.3).cuda()
y3 = torch.empty(B * n, 3 * C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.call("mvuld_rs_gcn_affinity_f32", tpg, y3, None, B, n, C)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    _lib.call("mvuld_rs_gcn_affinity_f32", tpg, y3, None, B, n, C)
e.record()
torch.cuda.synchronize()
print("affinity us per launch", s.elapsed_time(e) / 20 * 1e3)
