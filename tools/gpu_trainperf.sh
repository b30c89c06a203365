#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_swin_train.py tests/test_gpu_train.py tests/test_gpu_autograd_boundary.py -x -q -m gpu -p no:cacheprovider > gpurun_out/train_tests.log 2>&1; echo "tests exit=$? $(tail -n 3 gpurun_out/train_tests.log)"
timeout 300 python tools/prof_swin_train.py > gpurun_out/prof_swin_train.log 2>&1; head -8 gpurun_out/prof_swin_train.log
