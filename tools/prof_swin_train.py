"""Per-entry-point GPU time of one SwinTrainer step (SwinV2-B 448 / w28) -- events around every C-ABI call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mvuld_b200 import synth, swin_train
import mvuld_b200 as mv
B = int(os.environ.get("PB", 32))
torch.manual_seed(0)
model = mv.build_model(mv.default_config()).eval()
synth.randomize_for_parity(model, seed=777)
model = model.cuda()
tr = swin_train.SwinTrainer(model, lr=1e-5, world_size=1)
x = synth.images(B, 448, seed=1).cuda()
y = torch.randint(0, 2, (B,)).cuda()
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); s.record()
for _ in range(3):
    tr.step(x, y)
e.record(); torch.cuda.synchronize()
print(f"SwinTrainer.step B={B}: {s.elapsed_time(e) / 3:.2f} ms per step (wall {1e3 * (time.time() - t0) / 3:.2f} ms)")
inst = bench.Instrument(); inst.install()
tr.step(x, y)
inst.remove()
fam, per = inst.summary()
tot = sum(v["ms"] for v in per.values())
print(f"instrumented total {tot:.2f} ms")
for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{v['ms']:9.3f} ms  {v['launches']:5d}  {k}")
# gemm breakdown by shape
shapes = {}
for name, a, s_, e_ in inst.records:
    if name == "mvuld_gemm_bf16":
        key = (a["M"], a["N"], a["K"])
        d = shapes.setdefault(key, [0.0, 0]); d[0] += s_.elapsed_time(e_); d[1] += 1
for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:14]:
    fl = 2.0 * k[0] * k[1] * k[2] * v[1]
    print(f"gemm M={k[0]:7d} N={k[1]:5d} K={k[2]:7d}: {v[0]:8.3f} ms x{v[1]:3d}  {fl / v[0] / 1e9:7.1f} TFLOP/s")
print("max memory GB", torch.cuda.max_memory_allocated() / 1e9)
# row / reduction kernels by their integer arguments (shape signature)
for ent in ("mvuld_ln_rows_bwd", "mvuld_colsum", "mvuld_swin_bias_grad", "mvuld_swin_attention_bwd", "mvuld_swin_attention_bwd_prep",
            "mvuld_gelu_bwd", "mvuld_swin_qkv_bwd", "mvuld_gemm_dw", "mvuld_transpose_bf16", "mvuld_ln_rows"):
    sig = {}
    for name, a, s_, e_ in inst.records:
        if name == ent:
            key = tuple(int(v) for v in a if isinstance(v, int) and not isinstance(v, bool))
            d = sig.setdefault(key, [0.0, 0]); d[0] += s_.elapsed_time(e_); d[1] += 1
    for k, v in sorted(sig.items(), key=lambda kv: -kv[1][0])[:6]:
        print(f"{ent} {k}: {v[0]:8.3f} ms x{v[1]:3d}  ({1e3 * v[0] / v[1]:7.1f} us each)")
