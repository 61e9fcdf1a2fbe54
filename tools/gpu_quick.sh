#!/bin/bash
# One gpurun call: a parity subset around the fused GEMM epilogue, then bench c4 and c5s (no CPU baseline).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 \
  -k "pipeline_matches_golden or fused_filter or synthetic_against_oracle or config5 or c4_ or does_not_fit or kb_block or gram_schmidt" \
  > gpurun_out/quick_tests.log 2>&1
grep -E 'passed|failed|FAILED|Error|^E ' gpurun_out/quick_tests.log | cut -c1-300 | tail -12
for wl in c4 c5s; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err || tail -5 gpurun_out/bench_$wl.err
done
python - <<'PY'
import json
for wl in ('c4', 'c5s'):
    d = json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    print(wl, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 3), 'kernels', d['kernel_ms_median'])
    print(wl, 'phi stored', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['phi_stored_ms'].items() if k != 'note'})
    print(wl, 'stage calls', round(d['stage_calls_ms']['ms_per_step'], 3), 'staged', {k: round(v, 3) for k, v in d['staged_ms'].items() if k != 'note'})
PY
