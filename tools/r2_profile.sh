#!/bin/bash
# ncu captures of the step's top kernels (one GPU):  gpurun --timeout 1500 -- 'bash tools/r2_profile.sh r02b'
set -u
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correction or half_join or first_stage" 2>&1 | tail -5 | tee $O/${TAG}_pytest_new.log
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_probe_(lb|chains)" -s 6 -c 2 \
  -o $O/${TAG}_probe python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_probe.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 12 -c 4 \
  -o $O/${TAG}_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_fused.log 2>&1
ls -la $O | tail -8
