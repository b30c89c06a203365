#!/bin/bash
# swin training tests (full size) + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_swin_train.py -q -x -m gpu --timeout=900 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/swin_train_tests.log
echo "tests exit=$?"; tail -30 gpurun_out/swin_train_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench exit=$?"; tail -5 gpurun_out/r2_bench_default.err
