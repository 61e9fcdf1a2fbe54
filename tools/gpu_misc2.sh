#!/bin/bash
# One gpurun call: new tests (column strips, eigenvector dumps), then the small configs through bench.py.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rA --tb=short --timeout 300 -k "column_strip or eigenvector_dumps or binary" > gpurun_out/misc2_tests.log 2>&1
grep -E 'passed|failed|FAILED|Error|^E ' gpurun_out/misc2_tests.log | cut -c1-300 | tail -30
for wl in c2 hd c5s; do
  echo "== bench $wl =="
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - $wl <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
for k in ('config', 'value', 'ms_per_step', 'e2e', 'stage_ms', 'kernel_ms_median', 'cpu_baseline'):
    print(k, '=', d.get(k))
PY
  tail -3 gpurun_out/bench_$wl.err | cut -c1-300
done
