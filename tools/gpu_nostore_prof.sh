#!/bin/bash
# One gpurun call: bench c4, then one ncu full-set capture of the extrapolation GEMM with Phi not stored.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_grp.json 2> gpurun_out/bench_c4_grp.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4_grp.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'kernels', d['kernel_ms_median'])
print('no store', d['no_phi_store_ms'])
PY
python tools/run_nostore.py 0 6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tcgen05' -s 4 -c 1 -o gpurun_out/prof_nostore python tools/run_nostore.py 0 6 > gpurun_out/ncu_nostore.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_nostore.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
