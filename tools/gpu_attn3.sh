#!/bin/bash
# three-group window attention: parity + timing against the two-group kernel (+ optional trace)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k swin_attention 2>&1 | tail -5 | tee gpurun_out/attn3_test.log
PB=64 timeout 60 python tools/prof_attn.py 2>&1 | tee gpurun_out/attn3_time.log
ENTRY=mvuld_swin_window_attention_fixed PB=64 timeout 60 python tools/prof_attn.py 2>&1 | tee -a gpurun_out/attn3_time.log
if [ -f mvuld_b200/csrc/build/variants/lib_trace.so ]; then
  ENTRY=mvuld_swin_window_attention_fixed MVULD_LIB=mvuld_b200/csrc/build/variants/lib_trace.so timeout 60 python tools/trace_attn.py
fi
