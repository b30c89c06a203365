#!/bin/bash
# Round-end evidence in ONE GPU-box session: the whole `-m gpu` suite as the driver runs it, smoke, the bench lines of
# every workload, the ncu launch list of the bench command and a `--set full` capture of the GGNN segment-reduce.
# Usage: bash tools/gpu_round_end.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest -m gpu exit=$? $(tail -n 1 gpurun_out/pytest_gpu_$tag.log)"
bash tools/gpu_final.sh $tag
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train > gpurun_out/plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list exit=$?"
python tools/prof_gather.py > gpurun_out/prof_gather_plain_$tag.log 2>&1 && cat gpurun_out/prof_gather_plain_$tag.log &&
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:ggnn_gather_sum -s 3 -c 1 -o gpurun_out/prof_gather_$tag \
    python tools/prof_gather.py > gpurun_out/ncu_gather_$tag.log 2>&1
echo "gather capture exit=$?"
du -sh gpurun_out
