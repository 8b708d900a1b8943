#!/bin/bash
# 8-GPU weak-scaling bench line only:  gpurun --gpus 8 -- 'bash tools/gpu_r2_n8_bench.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_n8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_n8.err | grep "^{\"metric\"" > gpurun_out/bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print('N=8 clips/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
