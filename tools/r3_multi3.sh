#!/bin/bash
# One 2-GPU call: the GPU suite (one GPU), the 2-GPU bench twice, the 1-GPU bench.
set -u
TAG=${1:-r03p}
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
except Exception as e:
    print('$1 no bench line:', e)"
}
{
echo "== GPU suite"
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== 2 GPUs"
MZGPU_DEBUG=1 MZ_ORACLE_WORKERS=64 timeout 400 $TR bench.py --gpus 2 --steps 30 --warmup 5 2> $O/${TAG}_bench_n2.err | tee $O/${TAG}_bench_n2.json | show "n2   "
grep -h "\[mzgpu\]" $O/${TAG}_bench_n2.err | tail -2 | cut -c1-400
echo "== 2 GPUs again"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_bench2_n2.err | tee $O/${TAG}_bench2_n2.json | show "n2(2)"
echo "== 1 GPU"
MZ_ORACLE_WORKERS=32 timeout 300 python bench.py --steps 40 --warmup 5 2> $O/${TAG}_bench_n1.err | tee $O/${TAG}_bench_n1.json | show "n1   "
} 2>&1 | tee $O/${TAG}_multi.log
