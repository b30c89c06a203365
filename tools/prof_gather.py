"""Standalone timing of the GGNN segment-reduce kernel (configs[2] shape: 4096 graphs, 4 edge types, D = 200)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib, synth
B, T, D = int(os.environ.get("PB", 4096)), 4, 200
g = synth.ggnn_batch(B, seed=12345, n_etypes=T).to("cuda")
N, E = g.num_nodes(), g.num_edges()
indptr, idx_src, eids = g.in_csr()
et = g.edata["_ETYPE"]
status = torch.zeros(1, device="cuda", dtype=torch.int32)
et_sorted = torch.empty(E, device="cuda", dtype=torch.uint8)
_lib.call("mvuld_gather_etype", et, eids, E, T, et_sorted, status)
msgs = torch.randn(N, T * D, device="cuda").to(torch.bfloat16)
out = torch.empty(N, 2 * D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.call("mvuld_ggnn_gather_sum", msgs, indptr, idx_src, et_sorted, out, 2 * D, N, T, D)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    _lib.call("mvuld_ggnn_gather_sum", msgs, indptr, idx_src, et_sorted, out, 2 * D, N, T, D)
e.record()
torch.cuda.synchronize()
us = s.elapsed_time(e) / 20 * 1e3
alg = E * D * 2 + E * 5 + (N + 1) * 4 + N * D * 2          # SURVEY 8(d) row 2, bf16 messages and output
print(f"variant {os.environ.get('MVULD_GGNN_VARIANT', '0')}: gather {us:.1f} us, algorithmic {alg / 1e9:.2f} GB -> {alg / us / 1e3:.0f} GB/s; "
      f"N {N} E {E} checksum {float(out.float().sum()):.3f}")
