#!/bin/bash
# One GPU call that validates and measures the current tree (one GPU):
#   gpurun --timeout 1500 -- 'bash tools/r2_bundle.sh tag'
set -u
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
line() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step', d.get('per_step_ms'), 'host', d.get('host_ms_per_step'),
          'e2e', round(d['e2e']['value'] / 1e6, 1), 'launches', d.get('gpu_launches'), 'parity', (d.get('parity') or {}).get('ok'))
    for t in d['roofline']['top_kernels'][:5]: print('    ', t)
except Exception as e:
    print('$1 no bench line:', e)"
}
echo "== full GPU suite (defaults)"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/${TAG}_pytest_gpu.log
echo "== bench (defaults)"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json | line "default         "
echo "== merge-path kernels: merges alone"
MZGPU_MERGE_KERNELS=1 timeout 120 python tools/merge_bench.py 2>&1 | tail -8 | tee $O/${TAG}_merge_bench_kernels.log
echo "== merge-path kernels on the side stream: full GPU suite"
MZGPU_MERGE_KERNELS=1 MZGPU_SIDE_STREAM=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $O/${TAG}_pytest_gpu_side.log
echo "== merge-path kernels on the side stream: bench with the oracle parity check of the timed steps"
MZGPU_MERGE_KERNELS=1 MZGPU_SIDE_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 2> $O/${TAG}_bench_side.err | tee $O/${TAG}_bench_side.json | line "side merges     "
tail -3 $O/${TAG}_bench_side.err
echo "== the same, 60 timed steps"
MZGPU_MERGE_KERNELS=1 MZGPU_SIDE_STREAM=1 timeout 600 python bench.py --steps 60 --warmup 5 2> /dev/null | tee $O/${TAG}_bench_side60.json | line "side merges, 60 "
echo "== merge-path kernels on the main stream"
MZGPU_MERGE_KERNELS=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> /dev/null | line "merge kernels   "
