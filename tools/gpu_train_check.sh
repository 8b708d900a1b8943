#!/bin/bash
# First GPU contact of the N3 / N1 (training) kernels.  Order: cheapest and most diagnostic first, every step bounded by its own timeout,
# results in gpurun_out/.  --runxfail turns the provisional xfail markers into hard failures.
#   gpurun --timeout 2400 -- 'bash tools/gpu_train_check.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu --runxfail -p no:cacheprovider 2>&1 | tail -40 > $O/train_tests_n3.log; echo "N3 tests rc=$?" | tee -a $O/summary_train.txt
timeout 900 python -m pytest tests/test_train_encoders_gpu.py -q -m gpu --runxfail -p no:cacheprovider 2>&1 | tail -40 > $O/train_tests_n1.log; echo "N1 tests rc=$?" | tee -a $O/summary_train.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py tests/test_train_encoders_gpu.py -q -m gpu --runxfail \
    -p no:cacheprovider -k "dropout or transpose or colsum or layernorm_backward or sync_head_backward or droppath or (attention_train_forward and 30)" 2>&1 | tail -25 > $O/train_memcheck.log
echo "memcheck rc=$?" | tee -a $O/summary_train.txt
timeout 300 python tools/train_bench.py --batch 32 --segments 14 --steps 5 --warmup 3 2>&1 | tail -3 > $O/train_bench_sync_n1.json; echo "sync train bench rc=$?" | tee -a $O/summary_train.txt
timeout 600 python tools/train_bench.py --mode avclip --batch 8 --segments 8 --steps 3 --warmup 2 2>&1 | tail -3 > $O/train_bench_avclip_n1.json; echo "avclip train bench rc=$?" | tee -a $O/summary_train.txt
timeout 300 python tools/train_bench.py --mode avclip_fwd --batch 16 --segments 8 --steps 5 --warmup 3 2>&1 | tail -3 > $O/bench_avclip_fwd_n1.json; echo "avclip fwd bench (config 3) rc=$?" | tee -a $O/summary_train.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/train_launches.csv \
    python tools/train_bench.py --batch 8 --segments 14 --steps 1 --warmup 1 > $O/train_ncu.log 2>&1
python tools/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt 2>&1 || true
SFB_MB_CUBLAS=1 timeout 300 python tools/microbench.py 512 > $O/microbench_cublas.log 2>&1; echo "microbench rc=$?" | tee -a $O/summary_train.txt
cat $O/summary_train.txt; tail -5 $O/train_tests_n3.log; tail -5 $O/train_tests_n1.log; cat $O/train_bench_sync_n1.json $O/train_bench_avclip_n1.json
