"""Per-entry-point GPU time of one MVulDTrainer step (SwinV2-B + RoBERTa-base + fusion, 32 functions) -- events around
every C-ABI call; the SwinV2 entries are listed by tools/prof_swin_train.py, this one is for the text / fusion side."""
import os, sys, time, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
args = types.SimpleNamespace(batch=0, workload="full", padded_text=False)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
w = bench.build_train_workload(args, 0, dev, 1, encoders=os.environ.get("ENCODERS", "1") == "1")
d = w["to_dev"]()
for _ in range(2):
    w["step"](d)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); s.record()
for _ in range(3):
    w["step"](d)
e.record(); torch.cuda.synchronize()
print(f"MVulDTrainer.step B={w['units']}: {s.elapsed_time(e) / 3:.2f} ms per step (wall {1e3 * (time.time() - t0) / 3:.2f} ms)")
inst = bench.Instrument(); inst.install()
w["step"](d)
inst.remove()
fam, per = inst.summary()
print(f"instrumented total {sum(v['ms'] for v in per.values()):.2f} ms")
for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    print(f"{v['ms']:9.3f} ms  {v['launches']:5d}  {k}")
shapes = {}
for name, a, s_, e_ in inst.records:
    if name == "mvuld_gemm_bf16":
        key = (a["M"], a["N"], a["K"])
        q = shapes.setdefault(key, [0.0, 0]); q[0] += s_.elapsed_time(e_); q[1] += 1
for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"gemm M={k[0]:7d} N={k[1]:5d} K={k[2]:7d}: {v[0]:8.3f} ms x{v[1]:3d}  {2.0 * k[0] * k[1] * k[2] * v[1] / v[0] / 1e9:7.1f} TFLOP/s")
for ent in ("mvuld_seq_attention_bwd", "mvuld_seq_qkv_bwd", "mvuld_seq_attention_bwd_prep", "mvuld_embed_grad_rows", "mvuld_gemm_dw",
            "mvuld_colsum", "mvuld_ln_rows_bwd", "mvuld_gelu_bwd_colsum", "mvuld_transpose_bf16"):
    sig = {}
    for name, a, s_, e_ in inst.records:
        if name == ent:
            key = tuple(int(v) for v in a if isinstance(v, int) and not isinstance(v, bool))
            q = sig.setdefault(key, [0.0, 0]); q[0] += s_.elapsed_time(e_); q[1] += 1
    for k, v in sorted(sig.items(), key=lambda kv: -kv[1][0])[:4]:
        print(f"{ent} {k}: {v[0]:8.3f} ms x{v[1]:3d}  ({1e3 * v[0] / v[1]:7.1f} us each)")
