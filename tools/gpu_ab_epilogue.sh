#!/bin/bash
# Same-box A/B of the whole benchmark step: per-thread vs TMA fp32 epilogue x separate LayerNorm launches vs LayerNorm fused into the GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 1" "0 1" "1 1" "1 0"; do
  set -- $cfg
  line=$(SFB_GEMM_F32_TMA=$1 SFB_LN_FUSED=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' | tail -1)
  echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('F32_TMA=$1 LN_FUSED=$2  clips/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'gemm ms', round(r['gemm_ms_per_step'],1), 'layernorm ms', round(r['hbm_bound_kernels'].get('layernorm',{}).get('ms_per_step',0),2), 'sm_mhz', d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_epilogue.txt
done
