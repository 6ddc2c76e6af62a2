#!/bin/bash
# N-GPU validation: parity of the key-sharded dataflow (peer-memory exchange and NCCL exchange), then the bench.
#   gpurun --gpus N --timeout 1200 -- 'bash tools/r2_multi.sh N tag'
set -u
N=${1:-2}
TAG=${2:-r02m}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== one-GPU smoke of what changed last" | tee $O/${TAG}_multi.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_join or first_stage or join_core or q3_dataflow or peer_memory or incremental_full_size" 2>&1 | tail -4 | tee -a $O/${TAG}_multi.log
if [ "${PIPESTATUS[0]}" -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
echo "== parity, peer-memory exchange" | tee -a $O/${TAG}_multi.log
timeout 300 $TR tools/q3_multi_gpu_check.py 2>&1 | grep -E "PARITY|batch|hydrate|exchange|Error|error" | tail -14 | tee -a $O/${TAG}_multi.log
echo "== parity, NCCL exchange" | tee -a $O/${TAG}_multi.log
MZGPU_P2P=0 timeout 300 $TR tools/q3_multi_gpu_check.py 2>&1 | grep -E "PARITY|Error|error" | tail -4 | tee -a $O/${TAG}_multi.log
echo "== bench, peer-memory exchange" | tee -a $O/${TAG}_multi.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 2> $O/${TAG}_bench_n$N.err | tee $O/${TAG}_bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('p2p  ', round(d['value']/1e6,1), 'M rows/s', round(d['ms_per_step'],4), 'ms/step', d.get('per_step_ms'), 'e2e', round(d['e2e']['value']/1e6,1), 'hydration', d['hydration'], 'parity', d.get('parity'))" | tee -a $O/${TAG}_multi.log
tail -3 $O/${TAG}_bench_n$N.err
echo "== bench, NCCL exchange" | tee -a $O/${TAG}_multi.log
MZGPU_P2P=0 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_bench_nccl_n$N.err | tee $O/${TAG}_bench_nccl_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('nccl ', round(d['value']/1e6,1), 'M rows/s', round(d['ms_per_step'],4), 'ms/step', d.get('per_step_ms'), 'e2e', round(d['e2e']['value']/1e6,1), 'hydration', d['hydration'])" | tee -a $O/${TAG}_multi.log
