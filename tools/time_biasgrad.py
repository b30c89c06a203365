import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for n_win, nH in ((32, 16), (128, 8), (512, 4)):
    ws, N = 28, 784
    gt = torch.randn(n_win * nH, N, N, device=dev, generator=g).to(torch.bfloat16)
    side = 2 * ws - 1
    dtab = torch.zeros(nH, side * side, device=dev)
    splits = _lib.load().mvuld_swin_bias_grad_splits(n_win, nH, ws)
    part = torch.empty(splits, nH, ws, ws, side, device=dev)
    fn = lambda: _lib.call("mvuld_swin_bias_grad", gt, n_win, nH, ws, N, part, dtab)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) / 20 * 1000
    print(f"bias_grad n_win={n_win} nH={nH} splits={splits}: {t:7.1f} us  {gt.numel() * 2 / t / 1e6:5.2f} TB/s", flush=True)
