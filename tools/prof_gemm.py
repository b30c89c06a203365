"""Standalone launch of one GEMM shape (default: stage-2 fc1 + GELU) for timing / ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
M, N, K = [int(x) for x in os.environ.get("MNK", "50176,2048,512").split(",")]
act = int(os.environ.get("ACT", "1"))
g = torch.Generator().manual_seed(0)
A = (torch.randn(M, K, generator=g) * 0.5).to("cuda", torch.bfloat16)
W = (torch.randn(N, K, generator=g) * 0.05).to("cuda", torch.bfloat16)
bias = torch.randn(N, generator=g).cuda()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"gemm M={M} N={N} K={K} act={act}: {ms * 1000:.1f} us  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
