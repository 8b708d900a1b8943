#!/bin/bash
# Round 2, GPU call 4: gate + bench + N=1 sweep + launch list + ncu captures with the kernels as they stand
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | tail -40 > $O/gpu_tests.log; echo "gpu tests rc=${PIPESTATUS[0]}" | tee -a $O/summary4.txt
grep -h "passed\|failed" $O/gpu_tests.log | tee -a $O/summary4.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/summary4.txt
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_n1.log 2>$O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary4.txt
tail -1 $O/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('clips/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm frac', round(d['roofline']['frac'],3), 'step frac', round(d['roofline']['step_frac_canonical'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],1), 'launches', d['gpu_launches'], d.get('clocks'), d.get('cpu_baseline')); print({k:(round(v['ms_per_step'],2), round(v['frac_of_hbm_peak'],2)) for k,v in d['roofline']['hbm_bound_kernels'].items()})" | tee -a $O/summary4.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400 | tee -a $O/summary4.txt
timeout 600 python tools/sweep_multi.py 2>/dev/null | grep '^{"n_gpus"' > $O/sweep_multi_n1.jsonl; cat $O/sweep_multi_n1.jsonl | tee -a $O/summary4.txt
# launch list of one full-size step (skip weight preparation + 3 warm-up steps): count launches first
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary4.txt
python tools/summarize_launches.py $O/launches_all.csv > $O/launches_all_summary.txt 2>&1 || true
head -30 $O/launches_all_summary.txt | tee -a $O/summary4.txt
for k in "qkv:2" "proj_res:9" "fc1:16"; do n=${k%%:*}; s=${k##*:}; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s $s -c 1 -o $O/prof_gemm_$n python tools/microbench.py 512 > $O/ncu_$n.log 2>&1; echo "ncu gemm $n rc=$?" | tee -a $O/summary4.txt; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_space_tc python tools/microbench.py 512 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" | tee -a $O/summary4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm768 -s 2 -c 1 -o $O/prof_layernorm python tools/microbench.py 512 > $O/ncu_ln.log 2>&1; echo "ncu ln rc=$?" | tee -a $O/summary4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_time_mma -s 2 -c 1 -o $O/prof_attn_time python tools/microbench.py 512 > $O/ncu_time.log 2>&1; echo "ncu time rc=$?" | tee -a $O/summary4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_row1 -s 2 -c 1 -o $O/prof_attn_row1 python tools/microbench.py 512 > $O/ncu_row1.log 2>&1; echo "ncu row1 rc=$?" | tee -a $O/summary4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 2 -c 1 -o $O/prof_mel python tools/microbench.py 512 > $O/ncu_mel.log 2>&1; echo "ncu mel rc=$?" | tee -a $O/summary4.txt
echo DONE | tee -a $O/summary4.txt
