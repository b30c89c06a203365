"""Per-tile floor of the GEMM: time vs K at fixed M, N for the bf16 (TMA-store) epilogue with and without GELU."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
M, N = 50176, 2048
g = torch.Generator().manual_seed(0)
for act in (0, 1):
    for K in (64, 128, 256, 512, 1024, 2048):
        A = (torch.randn(M, K, generator=g) * 0.5).to("cuda", torch.bfloat16)
        W = (torch.randn(N, K, generator=g) * 0.05).to("cuda", torch.bfloat16)
        bias = torch.randn(N, generator=g).cuda()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        tiles_per_cta = (M / 128) * (N / 256) / 148
        print(f"act={act} K={K:5d}: {ms * 1000:7.1f} us  {2.0 * M * N * K / ms / 1e9:6.0f} TFLOP/s  "
              f"{ms * 1000 / tiles_per_cta:5.2f} us per 128x256 tile (MMA {K / 16 * 128 / 1965:5.2f} us at 1965 MHz)", flush=True)
