"""SwinV2 qkv projection (GEMM + per-head normalise + window scatter) at the stage-0 / 1 / 2 sizes of a 64-image batch."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
for B, H, C, nH in ((64, 112, 128, 4), (64, 56, 256, 8), (64, 28, 512, 16)):
    M = B * H * H
    x = rn(M, C).to(torch.bfloat16)
    w = (rn(3 * C, C) * 0.05).to(torch.bfloat16)
    qb, vb, qs = rn(C) * 0.1, rn(C) * 0.1, torch.full((nH,), 14.4, device="cuda")
    q = torch.empty(M * C, device="cuda", dtype=torch.float16); k = torch.empty_like(q)
    v = torch.empty(M * C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: _lib.call("mvuld_swin_qkv", x, w, qb, vb, qs, q, k, v, B, H, H, C, nH, 28, 14)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) / 20 * 1000
    print(f"swin_qkv M={M} C={C}: {t:7.1f} us  {2.0 * M * 3 * C * C / t / 1e6:5.0f} TFLOP/s  {M * C * 8 / t / 1e6:5.2f} TB/s", flush=True)
