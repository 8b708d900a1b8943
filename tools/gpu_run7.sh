#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 60 --timeout-method thread"
timeout 150 $PT tests/test_model_gpu.py -s > $O/pytest_model.log 2>&1; echo "model rc=$?" >> $O/summary.txt
timeout 200 python bench.py --steps 5 --warmup 3 > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
# launch list of one full-size step (B=64): 3 warm-up steps skipped by -s
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file $O/launches_b64.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 7 -c 4 -o $O/prof_gemm_r1 \
   python tools/microbench.py 512 > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_mma_kernel -c 1 -o $O/prof_attn_space_r1 \
   python tools/microbench.py 512 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" >> $O/summary.txt
cat $O/summary.txt
grep -E "parity|passed|failed" $O/pytest_model.log | tail -3
tail -1 $O/bench.log
