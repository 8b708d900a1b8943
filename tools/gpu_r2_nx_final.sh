#!/bin/bash
# Round 2, final N-GPU evidence (N = 2 or 4):  gpurun --gpus N -- 'bash tools/gpu_r2_nx_final.sh N'   (weak-scaling bench line + strong-scaling
# sweep; at N = 2 also the 2-GPU bit-equality tests)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build_n$N.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build_n$N.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee $O/multigpu_tests_n2.log
fi
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 2>$O/bench_n$N.err | grep "^{\"metric\"" > $O/bench_n$N.json; echo "bench N=$N rc=${PIPESTATUS[0]}" | tee -a $O/summary_n$N.txt
cat $O/bench_n$N.json | cut -c1-400 | tee -a $O/summary_n$N.txt
timeout 900 $TR tools/sweep_multi.py 2>$O/sweep_n$N.err | grep '^{"n_gpus"' > $O/sweep_multi_n$N.jsonl; echo "sweep N=$N rc=${PIPESTATUS[0]} lines=$(wc -l < $O/sweep_multi_n$N.jsonl)" | tee -a $O/summary_n$N.txt
cat $O/sweep_multi_n$N.jsonl | tee -a $O/summary_n$N.txt
echo DONE | tee -a $O/summary_n$N.txt
