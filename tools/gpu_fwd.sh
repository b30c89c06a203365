#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/all_tests.log 2>&1; echo "tests exit=$? $(tail -n 3 gpurun_out/all_tests.log)"
timeout 900 python bench.py --no-cpu-baseline --no-train --no-sub > gpurun_out/bench_fwd.json 2> gpurun_out/bench_fwd.err; echo "bench exit=$?"
timeout 600 python bench.py --workload swin --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swin.json 2> gpurun_out/bench_swin.err; echo "bench swin exit=$?"
python - <<'P'
import json
for f in ('gpurun_out/bench_fwd.json','gpurun_out/bench_swin.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'roofline', d.get('roofline',{}).get('frac'))
            print(json.dumps(d.get('entry_points_ms')))
P
