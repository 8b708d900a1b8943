#!/bin/bash
# Round 2, final 8-GPU evidence:  gpurun --gpus 8 -- 'bash tools/gpu_r2_n8_final.sh'   (weak-scaling bench line + strong-scaling sweep)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
N=8
python -c "import __graft_entry__ as g; g.build()" > $O/build_n$N.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build_n$N.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/nccl_n8.%p.log timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 2>$O/bench_n8.err | grep "^{\"metric\"" > $O/bench_n8.json; echo "bench N=8 rc=${PIPESTATUS[0]}" | tee -a $O/summary_n8.txt
cat $O/bench_n8.json | cut -c1-700 | tee -a $O/summary_n8.txt
cat $O/nccl_n8.*.log 2>/dev/null | grep -m4 -i "NVLS\|via P2P\|Connected all" | tee -a $O/summary_n8.txt; rm -f $O/nccl_n8.*.log
timeout 600 $TR tools/sweep_multi.py 2>$O/sweep_n$N.err | grep '^{"n_gpus"' > $O/sweep_multi_n$N.jsonl; echo "sweep N=$N rc=${PIPESTATUS[0]} lines=$(wc -l < $O/sweep_multi_n$N.jsonl)" | tee -a $O/summary_n8.txt
cat $O/sweep_multi_n$N.jsonl | tee -a $O/summary_n8.txt
echo DONE | tee -a $O/summary_n8.txt
