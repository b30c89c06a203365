#!/bin/bash
# One GPU-box session: parity tests, the three bench workloads, then the ncu launch list of the bench command.
# Usage: bash tools/gpu_round.sh <tag> [ncu]
tag=${1:-r1}
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash tools/gpu_session.sh > /dev/null 2>&1
cat gpurun_out/summary.txt
MVULD_BENCH_DETAIL=gpurun_out/detail_full_$tag.json python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err
echo "bench full exit=$?"; cat gpurun_out/bench_full_$tag.json
python bench.py --workload swin --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swin_$tag.json 2> gpurun_out/bench_swin_$tag.err
echo "bench swin exit=$?"; cat gpurun_out/bench_swin_$tag.json
python bench.py --workload ggnn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ggnn_$tag.json 2> gpurun_out/bench_ggnn_$tag.err
echo "bench ggnn exit=$?"; cat gpurun_out/bench_ggnn_$tag.json
if [ "$2" = "ncu" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch_$tag.log 2>&1
  echo "ncu launch list exit=$?"
fi
