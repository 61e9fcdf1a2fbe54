#!/bin/bash
# One gpurun call: option kb_block=32 (32-slot K_B blocks) against the default, tests and bench.
mkdir -p gpurun_out
echo "== kb_block test (both sizes) =="
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --tb=short --timeout 300 -k "kb_block" > gpurun_out/kb32_a.log 2>&1
grep -E 'passed|failed|FAILED|kb_block|Error|^E ' gpurun_out/kb32_a.log | cut -c1-300 | tail -20
echo "== suite subset with GLB200_KB_BLOCK=32 =="
GLB200_KB_BLOCK=32 timeout 900 python -m pytest tests -m gpu -q -rA --tb=short --timeout 300 \
  -k "pipeline_matches_golden or fused_filter or kb_cutoff or synthetic_against_oracle or config5 or c4_ or reference_python or gram_schmidt or large_sample" \
  > gpurun_out/kb32_b.log 2>&1
grep -E 'passed|failed|FAILED|err_|Error|^E ' gpurun_out/kb32_b.log | cut -c1-300 | tail -30
for blk in 64 32; do
  echo "== bench c4 kb_block=$blk =="
  GLB200_KB_BLOCK=$blk timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kb$blk.json 2> gpurun_out/bench_kb$blk.err
  python - $blk <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_kb{sys.argv[1]}.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'e2e', 'kernel_ms_median', 'kb_cutoff', 'roofline'):
    print(k, '=', d.get(k))
PY
  tail -5 gpurun_out/bench_kb$blk.err | cut -c1-300
done
