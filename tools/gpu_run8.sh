#!/bin/bash
# 2-GPU run: NCCL sharded-forward equality test + weak-scaling bench at N=1 and N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
nvidia-smi -L > $O/gpus.txt
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_multigpu_gpu.py > $O/pytest_multigpu.log 2>&1; echo "multigpu test rc=$?" >> $O/summary.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n1.log 2>&1; echo "bench n1 rc=$?" >> $O/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.log 2>&1; echo "bench n2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/pytest_multigpu.log
tail -1 $O/bench_n1.log | cut -c1-200; tail -1 $O/bench_n2.log | cut -c1-200
