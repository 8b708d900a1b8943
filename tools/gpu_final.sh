#!/bin/bash
# What the driver runs at round end: full GPU test-suite, smoke(), default bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" >> $O/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $O/bench_ref.log 2>&1; echo "bench reference rc=$?" >> $O/summary.txt
timeout 300 python bench.py > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; tail -1 $O/bench_ref.log | cut -c1-300; tail -1 $O/bench.log
