#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context, not the next group).
# Usage on the GPU box:  bash tools/gpu_session.sh [group ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  echo "=== $name ===" | tee -a gpurun_out/summary.txt
  timeout "$to" python -m pytest "$@" -q -m gpu -x --timeout=600 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
groups=${@:-"probe gemm attn rows graph models"}
for g in $groups; do
  case $g in
    probe)  run probe 300 tests/test_gpu_kernels.py -k "probe" ;;
    gemm)   run gemm 300 tests/test_gpu_kernels.py -k "gemm" ;;
    attn)   run attn 600 tests/test_gpu_kernels.py -k "attention" ;;
    rows)   run rows 300 tests/test_gpu_kernels.py -k "ln_rows or patch_embed" ;;
    graph)  run graph 600 tests/test_gpu_kernels.py -k "csr or segment or gat or rs_gcn" ;;
    models) run models 1500 tests/test_gpu_models.py ;;
  esac
done
cat gpurun_out/summary.txt
