#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" -p no:cacheprovider > gpurun_out/pairs_tests.log 2>&1; echo "tests exit=$? $(tail -n 3 gpurun_out/pairs_tests.log)"
timeout 200 python tools/time_gemm_shapes.py 2>&1 | tee gpurun_out/pairs_time.log
