#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 200 python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 100 --timeout-method thread tests/test_model_gpu.py -k "graph" -s > $O/pytest_graph.log 2>&1; echo "graph test rc=$?"
grep -E "latency|passed|failed|Error" $O/pytest_graph.log | head
