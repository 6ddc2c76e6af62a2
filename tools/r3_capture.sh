#!/bin/bash
# Final one-GPU capture of this round (under gpurun): GPU suite, bench line (CPU sweep included), ncu launch
# list of the bench command, ncu --set full of the step's top kernels (fused seal, probe chains, merge tiles).
#   gpurun --timeout 1200 -- 'bash tools/r3_capture.sh r02c'
# tools/import_profiles.py (no GPU) turns gpurun_out/ into the committed profiles/ summaries.
set -u
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/${TAG}_pytest_gpu.log
python bench.py --steps 40 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 900 $O/${TAG}_bench.json; echo
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 12 -c 4 \
  -o $O/${TAG}_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_fused.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:k_probe_(lb|chains)" -s 6 -c 2 \
  -o $O/${TAG}_probe python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_probe.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:k_mrg_(tiles|index|partition)" -s 30 -c 6 \
  -o $O/${TAG}_mrg python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_mrg.log 2>&1
timeout 100 python tools/fused_phases.py > $O/${TAG}_fused_phases.json 2> /dev/null
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${TAG}_clocks_idle.csv
ls -la $O | tail -16
