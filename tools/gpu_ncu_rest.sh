#!/bin/bash
# One `ncu --set full` launch of every kernel of the full step that has no dedicated capture yet (graph / fusion /
# row kernels, the fused GEMM epilogue variants, the sequence attention): first matching launches of one bench step.
# Usage: bash tools/gpu_ncu_rest.sh <tag>      (run only after the same bench command exited 0 without ncu)
tag=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-train"
$CMD > gpurun_out/plain_rest_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_rest_$tag.log; exit 1; }
NCU="timeout 600 ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
KRE='rs_gcn_affinity|gat_aggregate|gat_scores|gemm_ln_kernel|patch_embed|fusion_head|segment_mean|patch_merge_gather|unbatch_pad_bn|pos_branch|roberta_embed|ln_meanpool|csr_finish|csr_prepare|attn_fwd_kernel<1|attn_fwd_kernel<0, 32, 14|collate_edges'
$NCU -k "regex:$KRE" -c ${NCU_COUNT:-56} -o gpurun_out/prof_rest_$tag $CMD > gpurun_out/ncu_rest_$tag.log 2>&1
echo "rest capture exit=$?"; tail -2 gpurun_out/ncu_rest_$tag.log
$NCU -k "regex:EpiQkvSwin" -s 8 -c 1 -o gpurun_out/prof_qkvswin_$tag $CMD > gpurun_out/ncu_qkvswin_$tag.log 2>&1
echo "qkv swin capture exit=$?"
$NCU -k "regex:EpiQkvHeads" -c 1 -o gpurun_out/prof_qkvheads_$tag $CMD > gpurun_out/ncu_qkvheads_$tag.log 2>&1
echo "qkv heads capture exit=$?"
ls -la gpurun_out/prof_rest_$tag.ncu-rep gpurun_out/prof_qkv*_$tag.ncu-rep
