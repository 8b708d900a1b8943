#!/bin/bash
# Round 2, final 1-GPU evidence run: bench line, N=1 sweep, launch list of one full-size step, ncu --set full of the kernels that changed
# since call 4 (space attention + its fix-up kernel) and of the dominant GEMM (traffic for roofline.traffic).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee $O/summary_final.txt
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_n1.log 2>$O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary_final.txt
grep '^{"metric"' $O/bench_n1.log | tail -1 > $O/bench_n1.json
python - <<'PY' | tee -a $O/summary_final.txt
import json
d = json.load(open('gpurun_out/bench_n1.json'))
r = d['roofline']
print('clips/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'gemm frac', round(r['frac'], 3),
      'step frac', round(r['step_frac_canonical'], 3), 'gemm ms', round(r['gemm_ms_per_step'], 1), 'launches', d['gpu_launches'], d.get('clocks'), d.get('cpu_baseline'))
print({k: (round(v['ms_per_step'], 2), round(v['frac_of_hbm_peak'], 2)) for k, v in r['hbm_bound_kernels'].items()})
PY
timeout 600 python tools/sweep_multi.py 2>/dev/null | grep '^{"n_gpus"' > $O/sweep_multi_n1.jsonl; cat $O/sweep_multi_n1.jsonl | tee -a $O/summary_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary_final.txt
python tools/summarize_launches.py $O/launches_all.csv > $O/launches_all_summary.txt 2>&1 || true
head -30 $O/launches_all_summary.txt | tee -a $O/summary_final.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_space_tc python tools/microbench.py 512 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" | tee -a $O/summary_final.txt
timeout 300 ncu --set full --clock-control none -k regex:attn_space_fixup -s 2 -c 1 -o $O/prof_attn_space_fixup python tools/microbench.py 512 > $O/ncu_fixup.log 2>&1; echo "ncu fixup rc=$?" | tee -a $O/summary_final.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o $O/prof_gemm_qkv python tools/microbench.py 512 > $O/ncu_qkv.log 2>&1; echo "ncu gemm qkv rc=$?" | tee -a $O/summary_final.txt
SFB_MB_CUBLAS=0 timeout 300 python tools/microbench.py 512 2>/dev/null | grep -v '^{' > $O/microbench.txt; cat $O/microbench.txt | tee -a $O/summary_final.txt
echo DONE | tee -a $O/summary_final.txt
