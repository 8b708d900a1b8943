#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 100 $PT tests/test_kernels_gpu.py -k "divided" > $O/pytest_attn.log 2>&1; echo "attn tests (tma) rc=$?" >> $O/summary.txt
SFB_ATTN_TMA=0 timeout 100 $PT tests/test_kernels_gpu.py -k "divided" > $O/pytest_attn_notma.log 2>&1; echo "attn tests (cp.async) rc=$?" >> $O/summary.txt
timeout 150 $PT tests/test_model_gpu.py -s > $O/pytest_model.log 2>&1; echo "model rc=$?" >> $O/summary.txt
timeout 120 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
SFB_ATTN_TMA=0 timeout 120 python tools/microbench.py > $O/microbench_notma.log 2>&1; echo "microbench notma rc=$?" >> $O/summary.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
grep -E "^E  |passed|failed" $O/pytest_attn.log | head -5; grep -E "passed|failed" $O/pytest_attn_notma.log; grep -E "parity|passed|failed" $O/pytest_model.log | tail -3
grep -E "^attn|^layer" $O/microbench.log; echo "-- no tma"; grep -E "^attn space" $O/microbench_notma.log
tail -1 $O/bench.log | cut -c1-250
