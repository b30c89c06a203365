#!/bin/bash
# One ncu measurement of every kernel of the full step that has no dedicated `--set full` capture (graph / fusion / row
# kernels, the fused GEMM epilogue variants, the sequence attention): the roofline metrics only (a few replay passes per
# launch), written as CSV (an .ncu-rep of ~100 launches exceeds what gpurun copies back).
# Usage: bash tools/gpu_ncu_rest.sh <tag>      (run only after the same bench command exited 0 without ncu)
tag=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-train"
$CMD > gpurun_out/plain_rest_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_rest_$tag.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_bytes.sum
KRE='rs_gcn_affinity|gat_aggregate|gat_scores|gemm_ln_kernel|patch_embed|fusion_head|segment_mean|patch_merge_gather|unbatch_pad_bn|pos_branch|roberta_embed|ln_meanpool|csr_finish|csr_prepare|attn_fwd_kernel<1|attn_fwd_kernel<0, 32, 14|collate_edges|EpiQkvSwin|EpiQkvHeads|EpiGeneric|split3'
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k "regex:$KRE" -c ${NCU_COUNT:-160} --csv \
    --log-file gpurun_out/ncu_rest_$tag.csv $CMD > gpurun_out/ncu_rest_$tag.log 2>&1
echo "rest capture exit=$?"; tail -2 gpurun_out/ncu_rest_$tag.log; ls -la gpurun_out/ncu_rest_$tag.csv
