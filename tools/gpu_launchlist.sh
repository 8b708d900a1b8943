#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
# one full-size step (B=64, S=8): skip weight preparation + 3 warm-up steps (about 170 + 3*292 launches), take the next 292
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1046 -c 292 --csv --log-file $O/launches_b64.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
grep -c gemm $O/launches_b64.csv
