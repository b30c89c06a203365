"""Back-to-back launch timing of the training step's row / reduction kernels at the SwinV2-B shapes of a 32-image batch
(events around 20 launches: no host gaps), against the bytes each must move."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)


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


for M, C in ((25088, 512), (401408, 128), (100352, 256), (6272, 1024)):
    y = rn(M, C).to(torch.bfloat16)
    dout, gamma = rn(M, C), 1 + 0.1 * rn(C)
    dvb = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    part = _lib.ln_rows_bwd_partials(M, C, dev)
    t = timed(lambda: _lib.call("mvuld_ln_rows_bwd", y, None, gamma, dout, dvb, None, dg, db, dg, part, M, C, 1e-5, 1))
    by = M * C * (2 + 4 + 2)
    print(f"ln_rows_bwd M={M} C={C}: {t:7.1f} us  {by / t / 1e6:5.2f} TB/s", flush=True)
for M, C, mode in ((25088, 512, 1), (16384, 768, 2), (401408, 128, 1)):      # with the dense bias gradient (third column sum)
    y = rn(M, C).to(torch.bfloat16)
    dout, gamma, sc = rn(M, C), 1 + 0.1 * rn(C), rn(M, C)
    dvb, dv32 = torch.empty(M, C, device=dev, dtype=torch.bfloat16), torch.empty(M, C, device=dev)
    dg, db, dbi = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    part = _lib.ln_rows_bwd_partials(M, C, dev)
    t = timed(lambda: _lib.call("mvuld_ln_rows_bwd", y, sc if mode == 2 else None, gamma, dout, dvb, dv32 if mode == 2 else None,
                                dg, db, dbi, part, M, C, 1e-5, mode))
    by = M * C * (2 + 4 + 2 + (8 if mode == 2 else 0))
    print(f"ln_rows_bwd+dbias M={M} C={C} mode={mode}: {t:7.1f} us  {by / t / 1e6:5.2f} TB/s", flush=True)
for M, C in ((25088, 2048), (16384, 3072), (401408, 512)):
    pre, dh = rn(M, C).to(torch.bfloat16), rn(M, C).to(torch.bfloat16)
    dpre, dbias = torch.empty_like(pre), torch.zeros(C, device=dev)
    t = timed(lambda: _lib.gelu_bwd_colsum(pre, dh, dpre, dbias))
    print(f"gelu_bwd_colsum M={M} C={C}: {t:7.1f} us  {M * C * 6 / t / 1e6:5.2f} TB/s", flush=True)
for R, C in ((25088, 512), (25088, 1536), (25088, 2048), (401408, 128), (401408, 512), (401408, 384)):
    x = rn(R, C).to(torch.bfloat16)
    out = torch.zeros(C, device=dev)
    t = timed(lambda: _lib.colsum(x, 1, C, out, R, C))
    print(f"colsum R={R} C={C}: {t:7.1f} us  {R * C * 2 / t / 1e6:5.2f} TB/s", flush=True)
for n_win, nH in ((32, 16), (128, 8), (512, 4)):
    ws, N = 28, 784
    gt = rn(n_win * nH, N, N).to(torch.bfloat16)
    side = 2 * ws - 1
    dtab = torch.zeros(nH, side * side, device=dev)
    splits = _lib.load().mvuld_swin_bias_grad_splits(n_win, nH, ws)
    part = torch.empty(splits, nH, ws, ws, side, device=dev)
    t = timed(lambda: _lib.call("mvuld_swin_bias_grad", gt, n_win, nH, ws, N, part, dtab))
    print(f"bias_grad n_win={n_win} nH={nH} splits={splits}: {t:7.1f} us  {gt.numel() * 2 / t / 1e6:5.2f} TB/s", flush=True)
for n in (51380224, 205520896):
    pre, dh = rn(n // 512, 512).to(torch.bfloat16), rn(n // 512, 512).to(torch.bfloat16)
    dpre = torch.empty_like(pre)
    t = timed(lambda: _lib.call("mvuld_gelu_bwd", pre, dh, dpre, n))
    print(f"gelu_bwd n={n}: {t:7.1f} us  {n * 6 / t / 1e6:5.2f} TB/s", flush=True)
