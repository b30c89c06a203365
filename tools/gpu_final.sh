#!/bin/bash
# Final evidence run of a round: smoke, the bench lines of every workload (1 GPU), then the ncu artefacts.
tag=${1:-r1}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; echo "full exit=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref exit=$?"
for w in swin ggnn train lines; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "$w exit=$?"
done
python - <<PY
import json
for w in ("full", "ref", "swin", "ggnn", "train", "lines"):
    try:
        d = json.load(open(f"gpurun_out/bench_{w}_$tag.json"))
        print(w, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"))
        if "train" in d: print("   train leg", round(d["train"]["value"], 1), round(d["train"]["e2e"]["value"], 1))
    except Exception as e:
        print(w, "FAILED", e)
PY
