#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/all_tests.log 2>&1; echo "tests exit=$? $(tail -n 3 gpurun_out/all_tests.log)"
timeout 300 python tools/prof_swin_train.py > gpurun_out/prof_swin_train.log 2>&1; head -14 gpurun_out/prof_swin_train.log
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench exit=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_train.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','metric')})
        for k in ('train_encoders',):
            if k in d: print(k, round(d[k]['value'],1), round(d[k]['ms_per_step'],2))
P
