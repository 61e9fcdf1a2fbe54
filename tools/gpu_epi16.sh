#!/bin/bash
# One gpurun call: sixteen epilogue warps (option gemm_epi_warps) against eight: parity subset, then bench c5s and c4.
mkdir -p gpurun_out
for ew in 16 8; do
  echo "== parity subset GLB200_EPI_WARPS=$ew =="
  GLB200_EPI_WARPS=$ew timeout 900 python -m pytest tests -m gpu -q -rA --tb=short --timeout 300 \
    -k "pipeline_matches_golden or fused_filter or synthetic_against_oracle or config5 or c4_ or python_interface or kb_block" \
    > gpurun_out/epi${ew}_tests.log 2>&1
  grep -E 'passed|failed|FAILED|Error|^E ' gpurun_out/epi${ew}_tests.log | cut -c1-300 | tail -12
done
for wl in c5s c4; do
  for ew in 8 16; do
    echo "== bench $wl epi_warps=$ew =="
    GLB200_EPI_WARPS=$ew timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_epi$ew.json 2> gpurun_out/bench_${wl}_epi$ew.err
    python - $wl $ew <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_{sys.argv[1]}_epi{sys.argv[2]}.json').read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'kernel_ms_median'):
    print(k, '=', d.get(k))
print('e2e', d['e2e']['value'], 'no_phi_store', d.get('no_phi_store', d.get('side_legs', {})) if False else '')
PY
    tail -3 gpurun_out/bench_${wl}_epi$ew.err | cut -c1-300
  done
done
