#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider > $O/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?" >> $O/summary.txt
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider > $O/pytest_model.log 2>&1; echo "pytest model rc=$?" >> $O/summary.txt
timeout 600 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
SFB_GEMM_ORDER=1 timeout 600 python tools/microbench.py > $O/microbench_mfast.log 2>&1; echo "microbench mfast rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/pytest_kernels.log; tail -3 $O/pytest_model.log
grep -E "^gemm|^attn|^layer|^mel" $O/microbench.log
echo "--- m-fastest"; grep -E "^gemm" $O/microbench_mfast.log | head -6
