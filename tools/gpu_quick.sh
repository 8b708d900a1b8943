#!/bin/bash
# Quick GPU test of a pytest -k expression:  bash tools/gpu_quick.sh "forward_clip or golden"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 250 python -m pytest -q -m gpu -p no:cacheprovider -x --timeout 120 --timeout-method thread tests -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"
grep -E "^E  |passed|failed|Error" gpurun_out/pytest_quick.log | head -12
