#!/bin/bash
# One gpurun call: tests, smoke, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
export GLB200_VERBOSE=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== full GPU suite =="
timeout 1200 python -m pytest tests -m gpu -q -rA --tb=short --timeout 300 > gpurun_out/tests.log 2>&1
grep -E 'passed|failed|FAILED|err_|Error|Fatal|^E ' gpurun_out/tests.log | cut -c1-300 | tail -60
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c4 =="
unset GLB200_VERBOSE
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'stage_ms', 'kernel_ms_median', 'kb_cutoff', 'gemm', 'roofline', 'roofline_filter', 'cpu_baseline'):
    print(k, '=', d.get(k))
PY
tail -20 gpurun_out/bench_c4.err | cut -c1-300
