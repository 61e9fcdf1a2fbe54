#!/bin/bash
# One gpurun --gpus 8 call: the world-4 / world-8 parity cases, then config 5 and config 4 on 8 GPUs (each line self-checks its parity).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu8.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA --tb=short --timeout 300 -k "4-320 or 8-320 or cat_small" > gpurun_out/tests_mgpu8.log 2>&1
grep -E 'passed|failed|FAILED|SKIPPED|world=|Error|error' gpurun_out/tests_mgpu8.log | cut -c1-300 | tail -12
for wl in c5 c4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --workload $wl > gpurun_out/bench_${wl}_n8.json 2> gpurun_out/bench_${wl}_n8.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${wl}_n8.json').read().strip().splitlines()[-1])
    print('$wl', {k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'parity')}, d['e2e']['value'], d['stage_ms'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_${wl}_n8.err').read()[-2500:])
PY
done
