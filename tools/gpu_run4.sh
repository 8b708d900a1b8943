#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider"
timeout 300 $PT tests/test_kernels_gpu.py -k "tc1cta or im2col" > $O/t_1cta.log 2>&1; echo "1cta rc=$?" >> $O/summary.txt
timeout 120 $PT tests/test_kernels_gpu.py -k "tcpair" > $O/t_pair.log 2>&1; echo "pair rc=$?" >> $O/summary.txt
timeout 300 $PT tests/test_kernels_gpu.py -k "gemm and tc and not tc1cta and not tcpair" > $O/t_auto.log 2>&1; echo "auto rc=$?" >> $O/summary.txt
timeout 600 $PT tests/test_model_gpu.py > $O/pytest_model.log 2>&1; echo "pytest model rc=$?" >> $O/summary.txt
timeout 300 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
SFB_GEMM_IMPL=2 timeout 300 python tools/microbench.py > $O/microbench_1cta.log 2>&1; echo "microbench 1cta rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/t_1cta.log; tail -3 $O/t_pair.log; tail -3 $O/t_auto.log; tail -3 $O/pytest_model.log
echo "--- pair (auto)"; grep -E "^gemm" $O/microbench.log
echo "--- 1cta"; grep -E "^gemm" $O/microbench_1cta.log | head -6
tail -1 $O/bench.log | cut -c1-400
