#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c5_n$N.json').read().strip().splitlines()[-1])
    for k in ('value', 'ms_per_step', 'n_gpus', 'e2e', 'stage_ms', 'kb_cutoff', 'gemm', 'roofline', 'roofline_gemm_dense', 'config'): print(k, '=', d.get(k))
except Exception as e:
    print('bench parse failed', e)
    import re
    t = open('gpurun_out/bench_c5_n$N.err').read()
    print('\n'.join([l for l in t.splitlines() if 'GLError' in l or 'Error' in l][:10]))
PY
