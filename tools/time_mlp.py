"""SwinV2 Mlp + norm2 + residual at the stage-0 / stage-1 sizes of a 64-image batch: two kernels (fc1 GEMM with GELU ->
hidden in HBM -> fc2 GEMM + LayerNorm) against the fused kernel (csrc/mlp_ln.cu)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1000


for M, C in ((802816, 128), (200704, 256)):
    X = (rn(M, C) * 0.7).to(torch.bfloat16)
    W1, b1 = (rn(4 * C, C) * 0.08).to(torch.bfloat16), rn(4 * C) * 0.3
    W2, b2 = (rn(C, 4 * C) * 0.05).to(torch.bfloat16), rn(C) * 0.3
    gamma, beta = 1 + 0.1 * rn(C), 0.1 * rn(C)
    x32, xb = rn(M, C), X.clone()
    hid = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)

    def two():
        _lib.gemm(xb, W1, bias=b1, act=_lib.ACT_GELU, out_bf16=hid)
        _lib.gemm_ln(hid, W2, gamma, beta, 1e-5, bias=b2, shortcut=x32, x32=x32, xb=xb)

    def fused():
        _lib.mlp_ln(xb, W1, b1, W2, b2, gamma, beta, 1e-5, shortcut=x32, x32=x32, xb=xb)

    t2, t1 = timed(two), timed(fused)
    fl = 2.0 * M * C * 4 * C * 2
    by = 12.0 * M * C
    print(f"M={M} C={C}: two kernels {t2:7.1f} us, fused {t1:7.1f} us ({fl / t1 / 1e6:5.0f} TFLOP/s, "
          f"{by / t1 / 1e6:5.2f} TB/s of algorithmic bytes)", flush=True)
