#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 500 python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log | head -12
