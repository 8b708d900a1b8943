#!/bin/bash
# Round 2, last 1-GPU evidence refresh (after the TMA fp32 epilogue and the occupancy changes of the time / CLS attention kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $O/summary_final3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/summary_final3.txt
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_n1.log 2>$O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary_final3.txt
grep '^{"metric"' $O/bench_n1.log | tail -1 > $O/bench_n1.json
python - <<'PY' | tee -a $O/summary_final3.txt
import json
d = json.load(open('gpurun_out/bench_n1.json'))
r = d['roofline']
print('clips/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'gemm frac', round(r['frac'], 3),
      'step frac', round(r['step_frac_canonical'], 3), 'gemm ms', round(r['gemm_ms_per_step'], 1), 'launches', d['gpu_launches'], d.get('clocks'), d.get('cpu_baseline'))
print({k: (round(v['ms_per_step'], 2), round(v['frac_of_hbm_peak'], 2)) for k, v in r['hbm_bound_kernels'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary_final3.txt
python tools/summarize_launches.py $O/launches_all.csv > $O/launches_all_summary.txt 2>&1 || true
head -12 $O/launches_all_summary.txt | tee -a $O/summary_final3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_time_mma -s 2 -c 1 -o $O/prof_attn_time python tools/microbench.py 512 > $O/ncu_time.log 2>&1; echo "ncu time rc=$?" | tee -a $O/summary_final3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_space_tc python tools/microbench.py 512 > $O/ncu_attn.log 2>&1; echo "ncu space rc=$?" | tee -a $O/summary_final3.txt
timeout 300 ncu --set full --clock-control none -k regex:attn_row1 -s 2 -c 1 -o $O/prof_attn_row1 python tools/microbench.py 512 > $O/ncu_row1.log 2>&1; echo "ncu row1 rc=$?" | tee -a $O/summary_final3.txt
timeout 300 python tools/microbench.py 512 2>/dev/null | grep -v '^{' > $O/microbench.txt
echo DONE | tee -a $O/summary_final3.txt
