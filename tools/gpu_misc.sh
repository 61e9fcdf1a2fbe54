#!/bin/bash
mkdir -p gpurun_out
for wl in c2 hd; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', {k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['value'], d['stage_ms'])
PY
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gram-schmidt 1 > gpurun_out/bench_c4_gs.json 2> gpurun_out/bench_c4_gs.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_c4_gs.json').read().strip().splitlines()[-1])
print('c4 gs', {k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['stage_ms'])
PY
tail -3 gpurun_out/bench_c4_gs.err
