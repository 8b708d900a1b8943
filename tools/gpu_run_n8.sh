#!/bin/bash
# 8-GPU box: weak-scaling bench at N=8 and N=4 (N=1, 2 measured separately)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
nvidia-smi -L > $O/gpus.txt
for N in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.log 2>&1; echo "bench n$N rc=$?" >> $O/summary.txt
done
cat $O/summary.txt
for N in 8 4; do tail -1 $O/bench_n$N.log | cut -c1-220; done
