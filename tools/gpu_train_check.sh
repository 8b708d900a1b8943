#!/bin/bash
# First GPU contact of the N3 (training-step) kernels: run their parity tests for real (--runxfail turns the provisional xfail markers
# into hard failures), memcheck the small cases, then time the training step (BASELINE config 4 shape, one GPU).
#   gpurun --timeout 1500 -- 'bash tools/gpu_train_check.sh'
mkdir -p gpurun_out
python -m pytest tests/test_train_gpu.py -q -m gpu --runxfail -x 2>&1 | tail -40 | tee gpurun_out/train_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py -q -m gpu --runxfail \
    -k "dropout or transpose or colsum or layernorm_backward or sync_head_backward or (attention_train_forward and 30)" 2>&1 | tail -25 | tee gpurun_out/train_memcheck.log
python tools/train_bench.py --batch 32 --segments 14 --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/train_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py --batch 8 --segments 14 --steps 1 --warmup 1 > gpurun_out/train_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/train_launches.csv > gpurun_out/train_launches_summary.txt 2>&1 || true
