#!/bin/bash
# One gpurun call: isolated processes so that a faulting kernel cannot poison the other groups.
mkdir -p gpurun_out
export GLB200_VERBOSE=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== group 1: sampling/affinity/eigen ==" 
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "sampling or synthetic_image or affinity or eigensolver or errors" 2>&1 | tail -40 | tee gpurun_out/g1.log
echo "== group 2: pipeline with the CUDA-core checker GEMM =="
GLB200_GEMM=simple timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "pipeline_matches_golden or synthetic_against_oracle or filter_options or stage_by_stage or gram_schmidt" 2>&1 | tail -60 | tee gpurun_out/g2.log
echo "== group 3: tcgen05 diag =="
timeout 300 python tools/diag_gemm.py 2>&1 | tail -40 | tee gpurun_out/g3.log
echo "== group 4: full suite (tcgen05) =="
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -60 | tee gpurun_out/g4.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
