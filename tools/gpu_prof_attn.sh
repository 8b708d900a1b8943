#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_space_tc -s 2 -c 1 -o $O/prof_attn_tc python tools/microbench.py 512 > $O/ncu_attn_tc.log 2>&1; echo "rc=$?"
