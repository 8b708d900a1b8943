#!/bin/bash
# Round 2, GPU call 3: single-pass (norm-bound) softmax in the tcgen05 space attention; EMIT_LN with the transposing reduction.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; }
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > $O/kernel_tests.log; echo "kernel tests rc=${PIPESTATUS[0]}" | tee -a $O/summary3.txt
tail -6 $O/kernel_tests.log | tee -a $O/summary3.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s --deselect tests/test_kernels_gpu.py 2>&1 | tail -60 > $O/gpu_tests.log; echo "other gpu tests rc=${PIPESTATUS[0]}" | tee -a $O/summary3.txt
grep -h "per-stage\|parity vs\|latency\|passed\|failed\|fused-LayerNorm" $O/gpu_tests.log | tee -a $O/summary3.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_default.log 2>&1; echo "bench default rc=$?" | tee -a $O/summary3.txt
SFB_LN_FUSED=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_lnfused.log 2>&1; echo "bench SFB_LN_FUSED=1 rc=$?" | tee -a $O/summary3.txt
for f in default lnfused; do tail -1 $O/bench_$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$f', 'clips/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm frac', round(d['roofline']['frac'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],1), 'launches', d['gpu_launches'], 'checksum', d.get('logits_checksum'), d.get('clocks')); print({k:(round(v['ms_per_step'],2), round(v['frac_of_hbm_peak'],2)) for k,v in d['roofline']['hbm_bound_kernels'].items()})" | tee -a $O/summary3.txt; done
timeout 300 python tools/microbench.py 512 > $O/microbench.log 2>&1; echo "microbench rc=$?" | tee -a $O/summary3.txt
head -34 $O/microbench.log | tee -a $O/summary3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_space_tc python tools/microbench.py 512 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" | tee -a $O/summary3.txt
echo DONE | tee -a $O/summary3.txt
