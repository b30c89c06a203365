#!/bin/bash
# Round-2 evidence in ONE GPU-box session: the `-m gpu` suite, the bench lines (default line with every sub-object,
# reference arm, line encoding), the ncu launch list of the bench command and `--set full` captures of the window
# attention, the cluster GEMM + LayerNorm and the fc1 GEMM.  Every ncu pass runs only after its command exited 0 plain.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -n 1 gpurun_out/${tag}_pytest_gpu.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/${tag}_smoke.log)"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm exit=$?"
python bench.py > gpurun_out/${tag}_bench_full_1gpu.json 2> gpurun_out/${tag}_bench_full_1gpu.err; echo "bench exit=$?"
python bench.py --workload lines --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_lines_1gpu.json 2>/dev/null; echo "lines exit=$?"
python tools/batch_sweep.py > gpurun_out/${tag}_batch_sweep.log 2>&1; echo "sweep exit=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train --no-sub"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_full.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "launch list exit=$?"
ENTRY=mvuld_swin_window_attention_fixed PB=64 python tools/prof_attn.py > gpurun_out/${tag}_prof_attn_plain.log 2>&1 &&
ENTRY=mvuld_swin_window_attention_fixed PB=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_swin3 -s 3 -c 1 -f -o gpurun_out/${tag}_prof_attn python tools/prof_attn.py > gpurun_out/${tag}_ncu_attn.log 2>&1
echo "attention capture exit=$?"
python tools/run_glc_once.py 2048 > /dev/null 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_cluster -s 2 -c 1 -f -o gpurun_out/${tag}_prof_glc python tools/run_glc_once.py 2048 > gpurun_out/${tag}_ncu_glc.log 2>&1
echo "cluster kernel capture exit=$?"
ACT=1 python tools/prof_gemm.py > /dev/null 2>&1 &&
ACT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 3 -c 1 -f -o gpurun_out/${tag}_prof_gemm_fc1 python tools/prof_gemm.py > gpurun_out/${tag}_ncu_gemm.log 2>&1
echo "gemm capture exit=$?"
for f in attn glc gemm_fc1; do ncu -i gpurun_out/${tag}_prof_$f.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_${f}_raw.csv 2>/dev/null; done
du -sh gpurun_out
