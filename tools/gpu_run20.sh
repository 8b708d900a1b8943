#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 100 $PT tests/test_kernels_gpu.py -k "tcpair or (gemm and tc and not tc1cta and not tcquad)" > $O/t_pair.log 2>&1; echo "pair rc=$?" >> $O/summary.txt
timeout 100 $PT tests/test_kernels_gpu.py -k "tcquad" > $O/t_quad.log 2>&1; echo "quad rc=$?" >> $O/summary.txt
timeout 120 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench cl2 rc=$?" >> $O/summary.txt
SFB_GEMM_CL=4 timeout 120 python tools/microbench.py > $O/microbench_cl4.log 2>&1; echo "microbench cl4 rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -4 $O/t_pair.log; grep -E "^E  |passed|failed|Error" $O/t_quad.log | head -8
echo "-- cl2"; grep -E "^gemm" $O/microbench.log | head -6; echo "-- cl4"; grep -E "^gemm" $O/microbench_cl4.log | head -6
