#!/bin/bash
# A/B of an experiment switch on the whole benchmark step, e.g.  bash tools/gpu_ab_bench.sh SFB_TIME_CLS_FUSED=1
# (switches that act inside model.py / the attention dispatch are invisible to tools/microbench.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.log 2>&1; echo "A rc=$?"
env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.log 2>&1; echo "B ($*) rc=$?"
for f in a b; do python - "$f" <<'PY'
import json, sys
line = open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1]
d = json.loads(line)
print(sys.argv[1].upper(), 'clips/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 2), 'checksum', d.get('logits_checksum'), 'clocks', d.get('clocks'))
PY
done
