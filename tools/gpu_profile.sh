#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full-set captures of the hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
CMDH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload hd"
$CMDH > gpurun_out/plain_hd.log 2> gpurun_out/plain_hd.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tcgen05|k_filter_project|k_filter_apply|k_affinity_B|k_jacobi$' -s 10 -c 5 -o gpurun_out/prof_hd $CMDH > gpurun_out/ncu_hd.log 2>&1
echo "hd full rc=$?"
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tcgen05|k_filter_project|k_filter_apply|k_affinity_B' -s 8 -c 4 -o gpurun_out/prof_c4 $CMD > gpurun_out/ncu_c4.log 2>&1
echo "c4 full rc=$?"
ls -la gpurun_out | head -30
tail -3 gpurun_out/ncu_c4.log gpurun_out/ncu_hd.log gpurun_out/ncu_launch.log | cut -c1-300
