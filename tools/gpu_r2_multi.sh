#!/bin/bash
# Round 2 multi-GPU evidence on ONE box with N GPUs:  gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
#   N = 2: the 2-GPU bit-equality tests + strong-scaling sweep;  N = 4: sweep;  N = 8: sweep + config 4 (DDP training step over NCCL) + config 3
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build_n$N.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build_n$N.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee $O/multigpu_tests_n2.log
fi
timeout 900 $TR tools/sweep_multi.py 2>$O/sweep_n$N.err | grep '^{"n_gpus"' > $O/sweep_multi_n$N.jsonl; echo "sweep N=$N rc=${PIPESTATUS[0]} lines=$(wc -l < $O/sweep_multi_n$N.jsonl)" | tee -a $O/summary_multi_n$N.txt
cat $O/sweep_multi_n$N.jsonl | tee -a $O/summary_multi_n$N.txt
if [ "$N" = "8" ]; then
  timeout 400 $TR tools/train_bench.py --batch 32 --segments 14 --steps 5 --warmup 3 2>$O/train_n8.err | tail -1 > $O/train_bench_sync_n8.json; echo "config 4 (DDP, 8 x 32 clips) rc=${PIPESTATUS[0]}" | tee -a $O/summary_multi_n$N.txt
  cat $O/train_bench_sync_n8.json | tee -a $O/summary_multi_n$N.txt
  timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 2>$O/bench_n8.err | grep "^{\"metric\"" > $O/bench_n8.json; echo "bench N=8 rc=${PIPESTATUS[0]}" | tee -a $O/summary_multi_n$N.txt
  cat $O/bench_n8.json | cut -c1-600 | tee -a $O/summary_multi_n$N.txt
  grep -m3 -i "NVLS\|Using network\|via P2P" $O/bench_n8.err | tee -a $O/summary_multi_n$N.txt
fi
echo DONE | tee -a $O/summary_multi_n$N.txt
