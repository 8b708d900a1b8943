#!/bin/bash
# Round 2, first GPU call: gate (pytest -m gpu without xfail + the new tap tests), bench A/B of the fused time-attention CLS query,
# the reference-GPU denominator, the cuBLAS yardstick (+ one ncu capture of its kernel), configs 3 / 4 at N = 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu_info.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | tail -60 > $O/gpu_tests.log; echo "gpu tests rc=${PIPESTATUS[0]}" | tee -a $O/summary1.txt
grep -h "per-stage\|parity vs\|latency" $O/gpu_tests.log | tee -a $O/summary1.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_a.log 2>&1; echo "bench A rc=$?" | tee -a $O/summary1.txt
SFB_TIME_CLS_FUSED=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_b.log 2>&1; echo "bench B (SFB_TIME_CLS_FUSED=1) rc=$?" | tee -a $O/summary1.txt
for f in a b; do tail -1 $O/bench_$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$f', 'clips/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm frac', round(d['roofline']['frac'],3), 'checksum', d.get('logits_checksum'), d.get('clocks'))" | tee -a $O/summary1.txt; done
SYNCHFORMER_REF=$PWD/baseline/_ref timeout 900 python tools/ref_gpu_timing.py $O/reference_gpu.json > $O/reference_gpu.log 2>&1; echo "reference gpu timing rc=$?" | tee -a $O/summary1.txt
tail -12 $O/reference_gpu.log | tee -a $O/summary1.txt
SFB_MB_CUBLAS=1 timeout 300 python tools/microbench.py 512 > $O/microbench_cublas.log 2>&1; echo "microbench rc=$?" | tee -a $O/summary1.txt
head -30 $O/microbench_cublas.log | tee -a $O/summary1.txt
for dbg in 1 2 3 4 7; do echo "== SFB_GEMM_DBG=$dbg (1 no stores, 2 same-tile loads, 4 no residual loads)" >> $O/microbench_dbg.log; SFB_GEMM_DBG=$dbg SFB_MB_GEMM_ONLY=1 timeout 120 python tools/microbench.py 512 >> $O/microbench_dbg.log 2>&1; done
cat $O/microbench_dbg.log | tee -a $O/summary1.txt
for s in qkv fc1; do
  timeout 300 ncu --set full --clock-control none -k regex:'nvjet|cutlass|gemm|xmma' -s 4 -c 1 -o $O/prof_cublas_$s python tools/cublas_probe.py $s > $O/ncu_cublas_$s.log 2>&1; echo "ncu cublas $s rc=$?" | tee -a $O/summary1.txt
done
timeout 300 python tools/train_bench.py --batch 32 --segments 14 --steps 5 --warmup 3 2>&1 | tail -3 > $O/train_bench_sync_n1.json; echo "sync train bench rc=$?" | tee -a $O/summary1.txt
timeout 300 python tools/train_bench.py --mode avclip_fwd --batch 16 --segments 8 --steps 5 --warmup 3 2>&1 | tail -3 > $O/bench_avclip_fwd_n1.json; echo "avclip fwd bench (config 3) rc=$?" | tee -a $O/summary1.txt
timeout 600 python tools/train_bench.py --mode avclip --batch 8 --segments 8 --steps 3 --warmup 2 2>&1 | tail -3 > $O/train_bench_avclip_n1.json; echo "avclip train bench rc=$?" | tee -a $O/summary1.txt
cat $O/train_bench_sync_n1.json $O/bench_avclip_fwd_n1.json $O/train_bench_avclip_n1.json | tee -a $O/summary1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/train_launches.csv \
    python tools/train_bench.py --batch 8 --segments 14 --steps 1 --warmup 1 > $O/train_ncu.log 2>&1
python tools/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt 2>&1 || true
echo DONE | tee -a $O/summary1.txt
