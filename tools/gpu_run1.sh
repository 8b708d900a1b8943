#!/bin/bash
# First GPU contact: kernel tests in separate processes (a trap in one group must not poison the others), model parity,
# smoke, a short bench.  Everything is logged under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvsmi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider"
timeout 300 $PT tests/test_kernels_gpu.py -k "simple or layernorm or im2col or cast or mel or sync_tokens or aggregator" > $O/t1_simple.log 2>&1; echo "t1 rc=$?" >> $O/summary.txt
timeout 300 $PT tests/test_kernels_gpu.py -k "gemm and tc" > $O/t2_gemm_tc.log 2>&1; echo "t2 rc=$?" >> $O/summary.txt
timeout 300 $PT tests/test_kernels_gpu.py -k "linearity or patch_embed" > $O/t3_gemm_more.log 2>&1; echo "t3 rc=$?" >> $O/summary.txt
timeout 300 $PT tests/test_kernels_gpu.py -k "attention and tc" > $O/t4_attn_tc.log 2>&1; echo "t4 rc=$?" >> $O/summary.txt
timeout 600 $PT tests/test_model_gpu.py > $O/t5_model.log 2>&1; rc=$?; echo "t5 rc=$rc" >> $O/summary.txt
if [ $rc -ne 0 ]; then
  SFB_GEMM_IMPL=1 timeout 900 $PT tests/test_model_gpu.py -k "golden or fresh" > $O/t5b_model_simplegemm.log 2>&1; echo "t5b(simple gemm) rc=$?" >> $O/summary.txt
  SFB_GEMM_IMPL=1 SFB_ATTN_IMPL=1 timeout 900 $PT tests/test_model_gpu.py -k "golden" > $O/t5c_model_allsimple.log 2>&1; echo "t5c(all simple) rc=$?" >> $O/summary.txt
  SFB_ATTN_IMPL=1 timeout 900 $PT tests/test_model_gpu.py -k "golden" > $O/t5d_model_simpleattn.log 2>&1; echo "t5d(simple attn) rc=$?" >> $O/summary.txt
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.log 2>&1; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python tools/ref_gpu_timing.py > $O/ref_gpu.log 2>&1; echo "refgpu rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/bench.log
