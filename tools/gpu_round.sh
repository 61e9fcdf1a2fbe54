#!/bin/bash
# One gpurun call: tests, smoke, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
export GLB200_VERBOSE=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== full GPU suite =="
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --tb=line --timeout 300 > gpurun_out/tests.log 2>&1
grep -E 'passed|failed|FAILED|err_|Error|Fatal|^/root' gpurun_out/tests.log | cut -c1-300 | tail -80
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c4 =="
unset GLB200_VERBOSE
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -c 3000 gpurun_out/bench_c4.json; tail -20 gpurun_out/bench_c4.err | cut -c1-300
