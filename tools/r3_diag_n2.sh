#!/bin/bash
# 2-GPU host-side diagnosis: where does the host time of a step go, and is the box starving the ranks?
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r3_diag_n2.sh tag'
set -u
TAG=${1:-r03d}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
box() {
  echo "-- nproc $(nproc)  loadavg $(cat /proc/loadavg)"
  echo "-- cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)  cpuset $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null | cut -c1-60)"
  grep -E "nr_throttled|throttled_usec|nr_periods" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo
}
show() {
  python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().split('\n')[-1])
    print('$1', round(d['value']/1e6,1), 'M rows/s', round(d['ms_per_step'],4), 'ms/step', d.get('per_step_ms'), 'e2e', round(d['e2e']['value']/1e6,1))
    print('    host', d.get('host_ms_per_step'))
except Exception as e:
    print('$1 no bench line:', e)"
}
{
box
nvidia-smi topo -m 2>/dev/null | head -6
for cfg in "A_defaults:" "B_blocking_sync:MZGPU_BLOCKING_SYNC=1" "C_no_sampler:MZ_CLOCK_SAMPLER=0" "D_defaults_again:"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== $name $envs"
  env $envs timeout 240 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2> $O/${TAG}_$name.err | tee $O/${TAG}_$name.json | show "$name"
  box
done
} 2>&1 | tee $O/${TAG}_diag.log
