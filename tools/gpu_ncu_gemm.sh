#!/bin/bash
mkdir -p gpurun_out
ACT=1 python tools/prof_gemm.py; ACT=0 python tools/prof_gemm.py; MNK=32768,3072,768 ACT=1 python tools/prof_gemm.py; MNK=50176,512,2048 ACT=0 python tools/prof_gemm.py
ACT=1 python tools/prof_gemm.py > gpurun_out/prof_gemm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 3 -c 1 -f -o gpurun_out/prof_gemm_$1 env ACT=1 python tools/prof_gemm.py > gpurun_out/ncu_gemm_$1.log 2>&1
echo ncu exit=$?
