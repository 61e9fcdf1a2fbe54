#!/bin/bash
# gpu_round.sh plus the colour eighth of config 5
bash tools/gpu_round.sh
echo "== bench c5s =="
timeout 600 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5s.json 2> gpurun_out/bench_c5s.err
python - <<'PY'
import json
for wl in ('c4', 'c5s'):
    d = json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    print(wl, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'kernels', d['kernel_ms_median'])
    print(wl, 'phi stored', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['phi_stored_ms'].items() if k != 'note'})
    print(wl, 'stage calls', round(d['stage_calls_ms']['ms_per_step'], 3), 'staged', {k: round(v, 3) for k, v in d['staged_ms'].items() if k != 'note'})
PY
