#!/bin/bash
# One-GPU iteration check: the tests named in $1 (a -k expression), a bench line, optional diagnostics.
#   gpurun --timeout 900 -- 'bash tools/r2_iter.sh "column or half_join" tag [diag]'
set -u
K=${1:-"column"}
TAG=${2:-r02i}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -15 | tee $O/${TAG}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench.json").read().strip().split("\n")[-1])
print("value", round(d["value"] / 1e6, 1), "M rows/s", round(d["ms_per_step"], 4), "ms/step", d.get("per_step_ms"), "host_enqueue", d.get("host_ms_per_step"), "e2e", round(d["e2e"]["value"] / 1e6, 1), "parity", (d.get("parity") or {}).get("ok"))
for t in d["roofline"]["top_kernels"][:6]:
    print("  ", t)
PY
if [ "${3:-}" = "ab" ]; then
  for i in 1 2 3; do
    timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_join_many" 2>&1 | tail -1 | sed 's/^/[steal=1] /'
    MZGPU_PROBE_STEAL=0 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_join_many" 2>&1 | tail -1 | sed 's/^/[steal=0] /'
  done
  MZGPU_PROBE_STEAL=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('[steal=0] value', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step', d.get('per_step_ms'))
for t in d['roofline']['top_kernels'][:3]: print('  ', t)"
fi
if [ "${3:-}" = "diag" ] || [ "${4:-}" = "diag" ]; then
  timeout 200 python tools/diag_bulk.py cfg4 > $O/${TAG}_diag_cfg4.log 2>&1; tail -45 $O/${TAG}_diag_cfg4.log
  timeout 200 python tools/diag_bulk.py cfg2 10000000 > $O/${TAG}_diag_cfg2.log 2>&1; tail -45 $O/${TAG}_diag_cfg2.log
fi
