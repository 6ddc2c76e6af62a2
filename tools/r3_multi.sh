#!/bin/bash
# 2-GPU validation of the defaults (merge-path kernels on the side stream) against the previous ones:
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r3_multi.sh 2 tag'
set -u
N=${1:-2}
TAG=${2:-r03m}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value']/1e6,1), 'M rows/s', round(d['ms_per_step'],4), 'ms/step', d.get('per_step_ms'), 'host', d.get('host_ms_per_step'), 'e2e', round(d['e2e']['value']/1e6,1), 'hydration', d['hydration'], 'parity', d.get('parity'))
    for t in d['roofline']['top_kernels'][:8]: print('    ', t)
except Exception as e:
    print('$1 no bench line:', e)"
}
echo "== parity, peer-memory exchange, defaults" | tee $O/${TAG}_multi.log
timeout 240 $TR tools/q3_multi_gpu_check.py 2>&1 | grep -E "PARITY|batch|hydrate|exchange|Error|error" | tail -14 | tee -a $O/${TAG}_multi.log
echo "== bench, defaults (merge kernels on the side stream)" | tee -a $O/${TAG}_multi.log
MZ_ORACLE_WORKERS=64 timeout 500 $TR bench.py --gpus $N --steps 30 --warmup 5 2> $O/${TAG}_bench_n$N.err | tee $O/${TAG}_bench_n$N.json | show "side  " | tee -a $O/${TAG}_multi.log
tail -3 $O/${TAG}_bench_n$N.err
echo "== bench, previous defaults (fused merges on the main stream)" | tee -a $O/${TAG}_multi.log
MZGPU_SIDE_STREAM=0 MZGPU_MERGE_KERNELS=0 timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench_old_n$N.err | tee $O/${TAG}_bench_old_n$N.json | show "main  " | tee -a $O/${TAG}_multi.log
