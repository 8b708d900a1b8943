#!/bin/bash
# Development loop on a GPU box: kernel + model tests (fail fast, per-test timeout), per-kernel microbench, short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 150 $PT tests/test_kernels_gpu.py > $O/pytest_kernels.log 2>&1; echo "kernels rc=$?" >> $O/summary.txt
timeout 250 $PT tests/test_model_gpu.py > $O/pytest_model.log 2>&1; echo "model rc=$?" >> $O/summary.txt
timeout 120 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/pytest_kernels.log; tail -3 $O/pytest_model.log
grep -E "^gemm|^attn|^layer|^mel" $O/microbench.log
tail -1 $O/bench.log | cut -c1-300
