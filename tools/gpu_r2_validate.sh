#!/bin/bash
# round-2 validation: all gpu tests in one process, default bench line, reference arm (short)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider > gpurun_out/r2_gputests.log 2>&1
echo "gputests exit=$? $(tail -n 3 gpurun_out/r2_gputests.log)"
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench exit=$?"; cat gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref exit=$?"; cat gpurun_out/r2_bench_ref.json
