#!/bin/bash
# One GPU call that validates and measures the current tree (one GPU):
#   gpurun --timeout 1200 -- 'bash tools/r2_bundle.sh tag'
set -u
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $O/${TAG}_pytest_gpu.log
line() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step', d.get('per_step_ms'), 'host_enqueue', d.get('host_ms_per_step'),
          'e2e', round(d['e2e']['value'] / 1e6, 1), 'launches', d.get('gpu_launches'))
    for t in d['roofline']['top_kernels'][:4]: print('    ', t)
except Exception as e:
    print('$1 no bench line:', e)"
}
echo "== bench (default switches)"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json | line "default      "
echo "== bench, equal CTA shares per chain"
MZGPU_PROBE_SHARE=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> /dev/null | line "share=0      "
echo "== merge-path kernels: merges alone"
MZGPU_MERGE_KERNELS=1 timeout 120 python tools/merge_bench.py 2>&1 | tail -8 | tee $O/${TAG}_merge_bench_kernels.log
timeout 120 python tools/merge_bench.py 2>&1 | tail -8 | tee $O/${TAG}_merge_bench_fused.log
echo "== merge-path kernels: full GPU suite"
MZGPU_MERGE_KERNELS=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $O/${TAG}_pytest_gpu_merge_kernels.log
echo "== merge-path kernels: bench (with the oracle parity check of the timed steps)"
MZGPU_MERGE_KERNELS=1 python bench.py --steps 20 --warmup 5 2> /dev/null | tee $O/${TAG}_bench_merge_kernels.json | line "merge kernels"
python -c "
import json
d = json.loads(open('$O/${TAG}_bench_merge_kernels.json').read().strip().split('\n')[-1]); print('    parity', d.get('parity'))"
echo "== bulk regimes"
timeout 200 python tools/diag_bulk.py cfg4 > $O/${TAG}_diag_cfg4.log 2>&1; grep -E "rep|groups|big blocks" $O/${TAG}_diag_cfg4.log | tail -12
timeout 200 python tools/diag_bulk.py cfg2 10000000 > $O/${TAG}_diag_cfg2.log 2>&1; grep -E "rep . (seals|work)|out rows" $O/${TAG}_diag_cfg2.log | tail -9
