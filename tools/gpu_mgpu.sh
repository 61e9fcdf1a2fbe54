#!/bin/bash
# One gpurun --gpus N call: whole GPU suite (single- and multi-GPU tests, host binary), then bench at 1 and N GPUs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== GPU suite =="
timeout 1500 python -m pytest tests -m gpu -q -rA --tb=short --timeout 600 > gpurun_out/tests.log 2>&1
grep -E 'passed|failed|FAILED|SKIPPED|world=|Error|error' gpurun_out/tests.log | cut -c1-300 | tail -60
echo "== bench N=1 =="
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err | cut -c1-300
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n =="
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    tail -c 1500 gpurun_out/bench_n$n.json; tail -5 gpurun_out/bench_n$n.err | cut -c1-300
  fi
done
