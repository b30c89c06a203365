#!/bin/bash
# all gpu tests + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2_gputests.log 2>&1
echo "gputests exit=$? $(tail -n 3 gpurun_out/r2_gputests.log)"
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench exit=$?"; tail -5 gpurun_out/r2_bench_default.err
