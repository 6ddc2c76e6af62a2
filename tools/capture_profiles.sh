#!/bin/bash
# Run on the GPU box (under gpurun): bench line, kernel-level bench, ncu launch list of the bench
# command, and ncu --set full captures of the top kernels.  Outputs land in gpurun_out/; tools/
# import_profiles.py (run here, no GPU) turns them into the committed profiles/ summaries.
set -u
TAG=${1:-r01}
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 500 python bench_kernels.py --out $O/${TAG}_kernels.json > /dev/null 2> $O/${TAG}_kernels.err
timeout 100 python tools/fused_phases.py > $O/${TAG}_fused_phases.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 24 -c 3 \
  -o $O/${TAG}_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_probe_(lb|chains)" -s 8 -c 2 \
  -o $O/${TAG}_probe python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rs_onesweep -s 10 -c 1 \
  -o $O/${TAG}_onesweep python tools/big_kernels.py sort > $O/${TAG}_ncu_onesweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^k_probe$" -s 2 -c 2 \
  -o $O/${TAG}_probe_bulk python tools/big_kernels.py join > $O/${TAG}_ncu_probe_bulk.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${TAG}_clocks_idle.csv
ls -la $O | tail -20
