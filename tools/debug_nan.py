"""Debug helper: one FusionTrainer step, then the eval forward with every C-ABI call checked for non-finite outputs."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib, train
from tests import cases
from tests.test_gpu_train import _inputs, _make_trainer
DEV = "cuda"
model, sd, tr = _make_trainer(0.0, lr=1e-3, weight_decay=0.005)
steps = int(os.environ.get("STEPS", 3))
for step in range(steps):
    g, img, txt, labels = _inputs(seed=cases.SEED + step)
    loss, logits = tr.step(g.to(DEV), img.to(DEV), txt.to(DEV), labels.to(DEV))
    torch.cuda.synchronize()
    print("step", step, "loss", float(loss), "logits finite", bool(torch.isfinite(logits).all()))
bad = {k: v for k, v in model.state_dict().items() if not torch.isfinite(v.float()).all()}
print("non-finite state entries:", list(bad)[:10])
for k, v in model.state_dict().items():
    if "running_var" in k and float(v.min()) <= 0:
        print("non-positive running_var", k, float(v.min()))
orig = _lib.call
first = []
def checked(name, *args):
    rc = orig(name, *args)
    torch.cuda.synchronize()
    for i, a in enumerate(args):
        t = a.t if hasattr(a, "t") and isinstance(getattr(a, "t"), torch.Tensor) else a
        if isinstance(t, torch.Tensor) and t.is_floating_point() and t.numel() and not torch.isfinite(t.float()).all():
            if not first:
                first.append(name)
                if name == "mvuld_rs_gcn_affinity_f32":
                    tp = args[0]
                    print("  tpg absmax", float(tp.abs().max()), "shape", tuple(tp.shape), "stride", tp.stride(), "contig", tp.is_contiguous(),
                          "B n C", args[3:6], "ptr%16", tp.data_ptr() % 16)
                    B, n, C = args[3:6]
                    th, ph, gg = (tp[:, q * C:(q + 1) * C].reshape(B, n, C) for q in range(3))
                    Rr = th @ ph.transpose(1, 2) / n
                    yr = Rr @ gg
                    print("  torch ref: R absmax", float(Rr.abs().max()), "y absmax", float(yr.abs().max()))
                    yk = args[1].float().view(B, n, 3 * C)
                    print("  kernel y nan per graph", [int(torch.isnan(yk[b]).sum()) for b in range(B)])
                print("FIRST non-finite after", name, "arg", i, tuple(t.shape), t.dtype,
                      "nan", int(torch.isnan(t.float()).sum()), "inf", int(torch.isinf(t.float()).sum()))
    return rc
_lib.call = checked
g, img, txt, _ = _inputs(seed=cases.SEED + 7)
out = model.eval()(g.to(DEV), img.to(DEV), txt.to(DEV))
print("eval out", out)

# replay the failing call in isolation
_lib.call = orig
torch.manual_seed(0)
B, n, C = 6, 100, 512
for trial, tp in enumerate([torch.randn(B * n, 3 * C, device=DEV) * 0.3, torch.randn(B * n, 3 * C, device=DEV) * 3.0]):
    y3 = torch.zeros(B * n, 3 * C, device=DEV, dtype=torch.bfloat16)
    _lib.call("mvuld_rs_gcn_affinity_f32", tp, y3, None, B, n, C)
    torch.cuda.synchronize()
    print("replay", trial, "nan", int(torch.isnan(y3.float()).sum()), "absmax", float(y3.float().abs().max()))
    R = torch.zeros(B, n, n, device=DEV)
    _lib.call("mvuld_rs_gcn_affinity_f32", tp, y3, R, B, n, C)
    torch.cuda.synchronize()
    th, ph = tp[:, :C].view(B, n, C), tp[:, C:2 * C].view(B, n, C)
    Rref = th @ ph.transpose(1, 2) / n
    print("   with r_out: y nan", int(torch.isnan(y3.float()).sum()), "R nan", int(torch.isnan(R).sum()),
          "R err", float((R - Rref).abs().max()))
