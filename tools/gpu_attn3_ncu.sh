#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_autograd_boundary.py -q -x -m gpu 2>&1 | tail -25 > gpurun_out/train_tests.log
ENTRY=mvuld_swin_window_attention_fixed PB=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_swin3 -s 3 -c 1 -f -o gpurun_out/prof_attn3 python tools/prof_attn.py > gpurun_out/ncu_attn3.log 2>&1
echo ncu exit=$?
cat gpurun_out/train_tests.log
