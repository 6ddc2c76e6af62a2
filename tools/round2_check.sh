#!/bin/bash
# GPU validation + measurement in one call:  gpurun --timeout 1200 -- 'bash tools/round2_check.sh [tag]'
# Stages run smallest-first under their own short timeouts (a hung cooperative kernel must not take
# the box with it); a failing stage is re-run with the fast MSD path off (MZGPU_FUSED_FAST=0) and with
# deferred merges off (MZGPU_DEFER_MERGES=0) to say which layer broke; later stages still run.
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
LOG=$O/${TAG}_pytest.log
: > $LOG
FAILED=0
stage() {
  echo "== $1" | tee -a $LOG
  timeout "$2" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$3" 2>&1 | tail -12 | tee -a $LOG
  if [ "${PIPESTATUS[0]}" -ne 0 ]; then
    FAILED=1
    echo "   FAILED: $1; bisect:" | tee -a $LOG
    MZGPU_FUSED_FAST=0 timeout "$2" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$3" 2>&1 | tail -3 | sed 's/^/   [fast=0] /' | tee -a $LOG
    MZGPU_DEFER_MERGES=0 timeout "$2" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$3" 2>&1 | tail -3 | sed 's/^/   [defer=0] /' | tee -a $LOG
  fi
}
stage "fused kernel (consolidate / seal / merge)" 180 "consolidate or batcher_seal or batch_merge"
stage "multi-job seal" 90 "seal_many"
stage "spine (deferred merges)" 120 "spine"
stage "chained probes" 90 "half_join_many"
stage "fused first stage" 90 "delta_first_stage"
stage "reduce" 180 "reduce or topk"
stage "joins" 240 "join"
echo "== whole suite" | tee -a $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee -a $LOG
line() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step  e2e', round(d['e2e']['value'] / 1e6, 1),
          ' launches', d.get('gpu_launches'), [(k['kernel'], k['share'], k['launches_per_step']) for k in d['roofline']['top_kernels'][:6]])
except Exception as e:
    print('$1 no bench line:', e)"
}
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json | line "fast+defer   "
MZGPU_FUSED_FAST=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_bench_nofast.err | tee $O/${TAG}_bench_nofast.json | line "fast=0       "
MZGPU_DEFER_MERGES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_bench_nodefer.err | tee $O/${TAG}_bench_nodefer.json | line "defer=0      "
timeout 120 python tools/fused_phases.py > $O/${TAG}_fused_phases.json 2> $O/${TAG}_fused_phases.err
python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_fused_phases.json"))
    for s in d["synthetic"]:
        for p in s["phases"]:
            print("synthetic", s["n"], p["path"], p["total_us"], p["stamps"])
    for p in d["q3"][:12]:
        print("q3", p["rows"], p["row_bytes"], p["path"], p["ctas"], p["total_us"], p["stamps"])
except Exception as e:
    print("no phase file:", e)
PY
tail -3 $O/${TAG}_bench.err
echo "FAILED=$FAILED"
