#!/bin/bash
# One gpurun --gpus N call: multi-GPU tests, then bench at 1..N GPUs (C4) and C5 at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== multi-GPU tests =="
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA --tb=short --timeout 600 > gpurun_out/tests_mgpu.log 2>&1
grep -E 'passed|failed|FAILED|SKIPPED|world=|Error|error' gpurun_out/tests_mgpu.log | cut -c1-300 | tail -30
for n in ${NLIST:-2 4 8}; do
  if [ $n -le $N ]; then
    echo "== bench C4 N=$n =="
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    fi
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d['stage_ms'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_n$n.err').read()[-1500:])
PY
  fi
done
if [ $N -ge 2 ]; then
  echo "== bench C5 N=$N =="
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c5_n$N.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d['stage_ms'], d['kb_cutoff'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_c5_n$N.err').read()[-2500:])
PY
fi
