#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_session.sh attn > /dev/null 2>&1
cat gpurun_out/summary.txt
PB=16 python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 && cat gpurun_out/prof_attn_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/prof_attn_$1 env PB=16 python tools/prof_attn.py > gpurun_out/ncu_attn_$1.log 2>&1
echo ncu exit=$?
