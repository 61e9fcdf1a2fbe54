#!/bin/bash
# One gpurun call on the final build: full GPU suite, smoke, bench c4 with the driver's step counts, then the ncu launch
# list of five plain steps (the run under ncu is evidence for SHARES only, never a bench value).
bash tools/gpu_round.sh
echo "== bench c4, --steps 20 --warmup 5 (the driver's call) =="
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c4_s20.json 2> gpurun_out/bench_c4_s20.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4_s20.json').read().strip().splitlines()[-1])
print('s20', round(d['value'], 1), round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
PY
echo "== launch list =="
timeout 600 python tools/run_c4_once.py 5 > gpurun_out/once.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python tools/run_c4_once.py 5 > gpurun_out/ncu_final.log 2>&1
echo "launch list rc=$?"; cat gpurun_out/once.log | tail -2
echo "== ncu --set full: the kernels that changed last (patch affinity, Jacobi, Rayleigh) =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_patch_affinity|k_jacobi$|k_rayleigh_cols' --launch-skip 3 -c 3 -o gpurun_out/r02_final_small python tools/run_c4_once.py 3 > gpurun_out/ncu_final_small.log 2>&1
echo "ncu full rc=$?"
