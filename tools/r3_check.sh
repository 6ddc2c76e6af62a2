#!/bin/bash
# One-GPU validation of the opt-in merge-path kernels / side stream against the defaults (short):
#   gpurun --timeout 900 -- 'bash tools/r3_check.sh tag'
set -u
TAG=${1:-r03}
O=gpurun_out
mkdir -p $O
line() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step', d.get('per_step_ms'), 'host', d.get('host_ms_per_step'),
          'e2e', round(d['e2e']['value'] / 1e6, 1), 'launches', d.get('gpu_launches'), 'parity', (d.get('parity') or {}).get('ok'))
    for t in d['roofline']['top_kernels'][:6]: print('    ', t)
except Exception as e:
    print('$1 no bench line:', e)"
}
echo "== merges alone, default path"
timeout 120 python tools/merge_bench.py 2>&1 | tail -8 | tee $O/${TAG}_merge_bench_default.log
echo "== merges alone, merge-path kernels"
MZGPU_MERGE_KERNELS=1 timeout 120 python tools/merge_bench.py 2>&1 | tail -8 | tee $O/${TAG}_merge_bench_kernels.log
echo "== GPU suite, merge-path kernels on the main stream"
MZGPU_MERGE_KERNELS=1 timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/${TAG}_pytest_gpu_mk.log
echo "== GPU suite, merge-path kernels on the side stream"
MZGPU_MERGE_KERNELS=1 MZGPU_SIDE_STREAM=1 timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/${TAG}_pytest_gpu_side.log
echo "== bench: defaults"
timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench_default.err | tee $O/${TAG}_bench_default.json | line "default     "
echo "== bench: merge-path kernels, main stream"
MZ_ORACLE_WORKERS=32 MZGPU_MERGE_KERNELS=1 timeout 400 python bench.py --steps 40 --warmup 5 2> $O/${TAG}_bench_mk.err | tee $O/${TAG}_bench_mk.json | line "merge kernels"
echo "== bench: merge-path kernels, side stream"
MZ_ORACLE_WORKERS=32 MZGPU_MERGE_KERNELS=1 MZGPU_SIDE_STREAM=1 timeout 400 python bench.py --steps 40 --warmup 5 2> $O/${TAG}_bench_side.err | tee $O/${TAG}_bench_side.json | line "side merges "
tail -3 $O/${TAG}_bench_side.err
