#!/bin/bash
# 2-GPU check of the mid-size block cache: parity, bench with the per-step parity record, A/B without the cache.
#   gpurun --gpus 2 --timeout 700 -- 'bash tools/r3_multi2.sh tag'
set -u
TAG=${1:-r03n}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
show() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value']/1e6,1), 'M rows/s', round(d['ms_per_step'],4), 'ms/step', d.get('per_step_ms'), 'e2e', round(d['e2e']['value']/1e6,1), 'hydration', d['hydration'], 'parity', (d.get('parity') or {}).get('ok'))
    print('    host', d.get('host_ms_per_step'))
    print('    kernel time/step', round(d['roofline']['kernel_time_per_step_ms'], 4))
    for t in d['roofline']['top_kernels'][:8]: print('    ', t)
except Exception as e:
    print('$1 no bench line:', e)"
}
{
echo "== parity, peer-memory exchange"
timeout 240 $TR tools/q3_multi_gpu_check.py 2>&1 | grep -E "PARITY|Error|error" | tail -4
echo "== bench, defaults"
MZ_ORACLE_WORKERS=64 timeout 400 $TR bench.py --gpus 2 --steps 30 --warmup 5 2> $O/${TAG}_bench_n2.err | tee $O/${TAG}_bench_n2.json | show "cache "
grep -h "\[mzgpu\]" $O/${TAG}_bench_n2.err | tail -2
echo "== bench, MZGPU_MID_BLOCK_MB=0 (the driver's pool for every block below 512 MB)"
MZGPU_MID_BLOCK_MB=0 timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench_nocache_n2.err | tee $O/${TAG}_bench_nocache_n2.json | show "pool  "
echo "== bench, defaults again"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench2_n2.err | tee $O/${TAG}_bench2_n2.json | show "cache2"
} 2>&1 | tee $O/${TAG}_multi.log
