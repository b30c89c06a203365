"""Standalone launch of the Swin window-attention kernel (stage-2 shape) for ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
B, H, W, C, nH, ws, shift = int(os.environ.get("PB", 16)), 28, 28, 512, 16, 28, 0
n = B * H * W * C
g = torch.Generator().manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(B * nH, ws * ws, 32, generator=g), dim=-1).mul(14.0).to("cuda", torch.float16)
k = torch.nn.functional.normalize(torch.randn(B * nH, ws * ws, 32, generator=g), dim=-1).to("cuda", torch.float16)
v = torch.randn(B * nH, ws * ws, 32, generator=g).to("cuda", torch.bfloat16)
side = 2 * ws - 1
tab = (torch.rand(nH, side * side, generator=g) * 16 * 1.4427).cuda()
tmax = tab.max(1).values.contiguous()
qn = torch.full((nH,), 14.0, device="cuda") if os.environ.get("FIXED_REF", "1") == "1" else None
out = torch.empty(B * H * W, C, device="cuda", dtype=torch.bfloat16)
ENTRY = os.environ.get("ENTRY", "mvuld_swin_window_attention")
for _ in range(3):
    _lib.call(ENTRY, q, k, v, tab, tmax, qn, out, B, H, W, C, nH, ws, shift)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    _lib.call(ENTRY, q, k, v, tab, tmax, qn, out, B, H, W, C, nH, ws, shift)
e.record()
torch.cuda.synchronize()
print(ENTRY, "ms per launch", s.elapsed_time(e) / 10, "B", B)
