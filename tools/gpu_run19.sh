#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 100 $PT tests/test_kernels_gpu.py -k "divided" > $O/pytest_attn.log 2>&1; echo "attn tests (variant 2) rc=$?" >> $O/summary.txt
timeout 200 $PT tests/test_model_gpu.py > $O/pytest_model.log 2>&1; echo "model rc=$?" >> $O/summary.txt
timeout 120 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench v2 rc=$?" >> $O/summary.txt
SFB_ATTN_TC_VARIANT=1 timeout 120 python tools/microbench.py > $O/microbench_v1.log 2>&1; echo "microbench v1 rc=$?" >> $O/summary.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1; echo "bench v2 rc=$?" >> $O/summary.txt
SFB_ATTN_TC_VARIANT=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_v1.log 2>&1; echo "bench v1 rc=$?" >> $O/summary.txt
cat $O/summary.txt
grep -E "^E  |passed|failed" $O/pytest_attn.log | head -5; grep -E "passed|failed" $O/pytest_model.log | tail -2
echo "-- variant 2"; grep -E "^attn space|^attn time|^layer" $O/microbench.log; echo "-- variant 1"; grep -E "^attn space|^attn time|^layer" $O/microbench_v1.log
tail -1 $O/bench.log | cut -c1-200; tail -1 $O/bench_v1.log | cut -c1-200
