#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_swin_train.py tests/test_gpu_autograd_boundary.py -q -x -m gpu --timeout=900 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/train_tests.log
echo "tests exit=$?"; tail -30 gpurun_out/train_tests.log
python tools/prof_swin_train.py 2>&1 | tail -50 > gpurun_out/prof_swin_train.log; cat gpurun_out/prof_swin_train.log
