#!/bin/bash
# Same-box A/B of the whole benchmark step: LayerNorm launches (0) vs the two qkv norms fused (2) vs all three fused (1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 2 1 2 0; do
  line=$(SFB_LN_FUSED=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' | tail -1)
  echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('SFB_LN_FUSED=$m  clips/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'gemm ms', round(r['gemm_ms_per_step'],1), 'layernorm ms', round(r['hbm_bound_kernels'].get('layernorm',{}).get('ms_per_step',0),2), 'sm_mhz', d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_lnfuse.txt
done
