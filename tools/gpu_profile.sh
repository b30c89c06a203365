#!/bin/bash
# One GPU-box session that refreshes the evidence under gpurun_out/ (copied into profiles/ afterwards):
#   launch list of the bench command, then one `ncu --set full` capture each of the GEMM, the window attention, the
#   GGNN segment-reduce and the LayerNorm row kernel -- every capture only after the same command exited 0 without ncu.
# Usage: bash tools/gpu_profile.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list exit=$?"
ACT=1 python tools/prof_gemm.py > gpurun_out/prof_gemm_plain_$tag.log 2>&1 && cat gpurun_out/prof_gemm_plain_$tag.log &&
$NCU -k regex:gemm_tn -s 3 -c 1 -o gpurun_out/prof_gemm_fc1_$tag env ACT=1 python tools/prof_gemm.py > gpurun_out/ncu_gemm_$tag.log 2>&1
echo "gemm capture exit=$?"
MNK=32768,768,3072 ACT=0 python tools/prof_gemm.py > gpurun_out/prof_gemm2_plain_$tag.log 2>&1 && cat gpurun_out/prof_gemm2_plain_$tag.log &&
$NCU -k regex:gemm_tn -s 3 -c 1 -o gpurun_out/prof_gemm_fc2_$tag env MNK=32768,768,3072 ACT=0 python tools/prof_gemm.py > gpurun_out/ncu_gemm2_$tag.log 2>&1
echo "gemm2 capture exit=$?"
PB=64 python tools/prof_attn.py > gpurun_out/prof_attn_plain_$tag.log 2>&1 && cat gpurun_out/prof_attn_plain_$tag.log &&
$NCU -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn_$tag env PB=64 python tools/prof_attn.py > gpurun_out/ncu_attn_$tag.log 2>&1
echo "attention capture exit=$?"
python tools/prof_ggnn.py > gpurun_out/prof_ggnn_plain_$tag.log 2>&1 && cat gpurun_out/prof_ggnn_plain_$tag.log &&
$NCU -k regex:"ggnn_gather_sum|gru_gates|segment_sum" -s 14 -c 3 -o gpurun_out/prof_ggnn_$tag python tools/prof_ggnn.py > gpurun_out/ncu_ggnn_$tag.log 2>&1
echo "ggnn capture exit=$?"
$NCU -k regex:ln_rows -s 40 -c 1 -o gpurun_out/prof_ln_$tag python bench.py --workload swin --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_ln_$tag.log 2>&1
echo "ln capture exit=$?"
ls -la gpurun_out/*.ncu-rep
