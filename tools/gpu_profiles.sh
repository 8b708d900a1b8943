#!/bin/bash
# Round evidence: launch list of one full-size step + full ncu captures of the dominant kernels (one launch each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
# one step = 292 launches in round 1 before the CLS fusion, 280+12 merge after; skip weight preparation (~170) + 3 warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1046 -c 292 --csv --log-file $O/launches_b64.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o $O/prof_gemm_qkv python tools/microbench.py 512 > $O/ncu1.log 2>&1; echo "gemm qkv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 16 -c 1 -o $O/prof_gemm_fc1 python tools/microbench.py 512 > $O/ncu2.log 2>&1; echo "gemm fc1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 9 -c 1 -o $O/prof_gemm_proj_res python tools/microbench.py 512 > $O/ncu5.log 2>&1; echo "gemm proj rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_space_tc python tools/microbench.py 512 > $O/ncu3.log 2>&1; echo "attn rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_time_mma -s 2 -c 1 -o $O/prof_attn_time python tools/microbench.py 512 > $O/ncu4.log 2>&1; echo "time rc=$?"
