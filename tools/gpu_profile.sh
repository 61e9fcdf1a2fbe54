#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full-set capture of the hot kernels (C4).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tcgen05|k_filter_project|k_filter_apply|k_affinity_B|k_jacobi$|k_rayleigh_partial' -s 12 -c 5 -o gpurun_out/prof_c4 $CMD > gpurun_out/ncu_c4.log 2>&1
echo "c4 full rc=$?"
ls -la gpurun_out | head -30
tail -3 gpurun_out/ncu_c4.log gpurun_out/ncu_launch.log | cut -c1-300
