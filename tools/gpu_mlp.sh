#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "mlp_ln" -p no:cacheprovider > gpurun_out/mlp_tests.log 2>&1; echo "tests exit=$? $(tail -n 15 gpurun_out/mlp_tests.log)"
timeout 120 python tools/time_mlp.py 2>&1 | tee gpurun_out/mlp_time.log
