"""Diagnostic: per-tensor gradient errors of the SwinV2 backward against the oracle, with reference norms."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases
from tests.test_gpu_swin_train import _encoder_case, rel_err
name = sys.argv[1] if len(sys.argv) > 1 else "small_ws7"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tr, feat, feats_ref, grads, gref = _encoder_case(name, B, cases.SEED + 21)
print("features rel err", rel_err(feat, feats_ref))
rows = []
for n, r in gref.items():
    if n in grads:
        rows.append((rel_err(grads[n].reshape(r.shape), r), float(r.norm()), n))
rows.sort(reverse=True)
tot = sum(r[1] ** 2 for r in rows) ** 0.5
print("flat norm", tot)
for e, nr, n in rows[:25]:
    print(f"{e:9.3e}  |ref| {nr:10.3e}  {n}")
