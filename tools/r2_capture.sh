#!/bin/bash
# Round-2 one-GPU capture (under gpurun): full GPU suite, bench line, phase stamps, ncu launch list of the
# bench command, ncu --set full of the step's top kernels, kernel-level bench and its two ncu captures.
#   gpurun --timeout 1500 -- 'bash tools/r2_capture.sh r02'
# tools/import_profiles.py (no GPU) turns gpurun_out/ into the committed profiles/ summaries.
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 600 $O/${TAG}_bench.json
timeout 100 python tools/fused_phases.py > $O/${TAG}_fused_phases.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 12 -c 4 \
  -o $O/${TAG}_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_probe_(lb|chains)" -s 6 -c 2 \
  -o $O/${TAG}_probe python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_probe.log 2>&1
timeout 500 python bench_kernels.py --out $O/${TAG}_kernels.json > /dev/null 2> $O/${TAG}_kernels.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rs_onesweep -s 10 -c 1 \
  -o $O/${TAG}_onesweep python tools/big_kernels.py sort > $O/${TAG}_ncu_onesweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^k_probe" -s 2 -c 2 \
  -o $O/${TAG}_probe_bulk python tools/big_kernels.py join > $O/${TAG}_ncu_probe_bulk.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${TAG}_clocks_idle.csv
ls -la $O | tail -24
