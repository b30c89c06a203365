#!/bin/bash
# Final round-2 evidence in ONE GPU-box session (after the fused Mlp kernel, the CTA-pair GEMM and the training row
# kernels): the `-m gpu` suite, smoke, the bench lines, the ncu launch list of the bench command, `--set full` captures of
# the fused Mlp kernel and the CTA-pair GEMM, the training-step profile.  Every ncu pass runs after its command exited 0.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest -m gpu exit=$? $(tail -n 1 gpurun_out/${tag}_pytest_gpu.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/${tag}_smoke.log)"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm exit=$?"
python bench.py > gpurun_out/${tag}_bench_full_1gpu.json 2> gpurun_out/${tag}_bench_full_1gpu.err; echo "bench exit=$?"
python tools/batch_sweep.py > gpurun_out/${tag}_batch_sweep.log 2>&1; echo "sweep exit=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-train --no-sub"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_full.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "launch list exit=$?"
python tools/time_mlp.py > gpurun_out/${tag}_time_mlp.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_ln -s 3 -c 1 -f -o gpurun_out/${tag}_prof_mlp python tools/time_mlp.py > gpurun_out/${tag}_ncu_mlp.log 2>&1
echo "mlp capture exit=$?"
MNK=16896,768,3072 ACT=0 python tools/prof_gemm.py > gpurun_out/${tag}_prof_gemm_pair_plain.log 2>&1 &&
MNK=16896,768,3072 ACT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 3 -c 1 -f -o gpurun_out/${tag}_prof_gemm_pair python tools/prof_gemm.py > gpurun_out/${tag}_ncu_gemm_pair.log 2>&1
echo "pair gemm capture exit=$?"
python tools/prof_swin_train.py > gpurun_out/${tag}_prof_swin_train.log 2>&1; echo "train profile exit=$?"
python tools/time_rowkernels.py > gpurun_out/${tag}_rowkernels.log 2>&1; echo "row kernels exit=$?"
python tools/time_gemm_shapes.py > gpurun_out/${tag}_gemm_shapes.log 2>&1
for f in mlp gemm_pair; do ncu -i gpurun_out/${tag}_prof_$f.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_${f}_raw.csv 2>/dev/null; done
du -sh gpurun_out
