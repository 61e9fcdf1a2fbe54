#!/bin/bash
# One gpurun --gpus N call: config 4 (and config 5 with C5=1) at N GPUs, 10 steps; prints the line's headline fields.
N=${1:-2}
mkdir -p gpurun_out
for wl in c4 ${C5:+c5}; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > gpurun_out/r02_scale_${wl}_n$N.json 2> gpurun_out/r02_scale_${wl}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r02_scale_${wl}_n$N.json').read().strip().splitlines()[-1])
    print('$wl N=$N', {k: d.get(k) for k in ('value', 'ms_per_step')}, 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity']['ok'], d['stage_ms'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r02_scale_${wl}_n$N.err').read()[-2000:])
PY
done
