#!/bin/bash
# One gpurun call: one ncu full-set capture of the extrapolation GEMM with Phi not stored (config 4).
mkdir -p gpurun_out
python tools/run_nostore.py 0 6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tcgen05' -s 4 -c 1 -f -o gpurun_out/prof_nostore python tools/run_nostore.py 0 6 > gpurun_out/ncu_nostore.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_nostore.log | cut -c1-200
