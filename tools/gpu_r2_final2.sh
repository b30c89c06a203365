#!/bin/bash
# Refresh of the round-2 evidence after the last training-step changes (the forward path is unchanged since
# tools/gpu_r2_final.sh ran): `-m gpu` suite, smoke, default bench line, training-step profiles, row-kernel timings.
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -n 1 gpurun_out/${tag}_pytest_gpu.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/${tag}_smoke.log)"
python bench.py > gpurun_out/${tag}_bench_full_1gpu.json 2> gpurun_out/${tag}_bench_full_1gpu.err; echo "bench exit=$?"
python tools/prof_swin_train.py > gpurun_out/${tag}_prof_swin_train.log 2>&1; echo "swin train profile exit=$?"
python tools/prof_joint_train.py > gpurun_out/${tag}_prof_joint_train.log 2>&1; echo "joint train profile exit=$?"
python tools/time_rowkernels.py > gpurun_out/${tag}_rowkernels.log 2>&1; echo "row kernels exit=$?"
