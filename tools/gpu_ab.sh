#!/bin/bash
# A/B microbench of an experiment switch given as env assignment(s), e.g.  bash tools/gpu_ab.sh SFB_GEMM_DIRECT=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 120 python tools/microbench.py > gpurun_out/microbench_a.log 2>&1; echo "A rc=$?"
env "$@" timeout 120 python tools/microbench.py > gpurun_out/microbench_b.log 2>&1; echo "B ($*) rc=$?"
echo "-- A (default)"; grep -E "^gemm|^attn space" gpurun_out/microbench_a.log | head -8
echo "-- B ($*)"; grep -E "^gemm|^attn space" gpurun_out/microbench_b.log | head -8
