#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python tools/microbench.py > $O/microbench.log 2>&1; echo "microbench rc=$?" >> $O/summary.txt
# launch list of one small step (B=8 clips): shares per kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_b8.csv \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
# full capture of the GEMM at the benchmark shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 6 -c 4 -o $O/prof_gemm \
   python tools/microbench.py 512 > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -5 $O/pytest_gpu.log
cat $O/microbench.log | head -30
