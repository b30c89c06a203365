#!/bin/bash
# quick GPU check: selected test groups + attention micro-timing (+ optional bench)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_session.sh ${GROUPS_:-"attn models"} > /dev/null 2>&1
cat gpurun_out/summary.txt
grep -E "Error|error|assert|FAILED" gpurun_out/attn.log gpurun_out/models.log | head -20
PB=64 python tools/prof_attn.py
PB=16 python tools/prof_attn.py
if [ -n "$BENCH" ]; then
  MVULD_BENCH_DETAIL=gpurun_out/detail_full_$BENCH.json python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_$BENCH.json 2> gpurun_out/bench_full_$BENCH.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_full_$BENCH.json')); print(d['value'], d['ms_per_step'], d['kernel_families'], d['entry_points_ms'])"
fi
