#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 150 $PT tests/test_kernels_gpu.py -k "gemm or patch_embed" > $O/t_gemm.log 2>&1; echo "gemm tests rc=$?" >> $O/summary.txt
timeout 120 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/t_gemm.log
grep -E "^gemm" $O/microbench.log | head -6
tail -1 $O/bench.log | cut -c1-200
