#!/bin/bash
# One gpurun call: the Phi-free fallback test, then config 5 (8192x8192 colour, p = 2000) on ONE GPU.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --tb=short --timeout 300 -k "does_not_fit or fused_filter or column_strip" > gpurun_out/c5one_tests.log 2>&1
grep -E 'passed|failed|FAILED|Error|^E ' gpurun_out/c5one_tests.log | cut -c1-300 | tail -12
echo "== bench c5 on one GPU =="
GLB200_VERBOSE=0 timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c5_n1.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'phi_stored', 'stage_ms', 'kb_cutoff', 'first_call_ms', 'phi_stored_ms', 'stage_calls_ms'):
    print(k, '=', d.get(k))
PY
tail -5 gpurun_out/bench_c5_n1.err | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv
