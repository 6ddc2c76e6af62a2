#!/bin/bash
# First GPU call for the round2-prep branch: the whole parity suite, then the bench with the new
# batching on and (for a bisect) with the deferred merges off.  Usage under gpurun:
#   gpurun --timeout 900 -- 'bash tools/round2_check.sh'
set -u
O=gpurun_out
mkdir -p $O
# Smallest new pieces first, each under its own short timeout: a hung cooperative kernel must not
# take the box (and a gpurun strike) with it.  Stop at the first failure.
step() {
  echo "== $1"
  timeout "$2" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$3" 2>&1 | tail -8 | tee -a $O/r2_pytest.log
  if [ "${PIPESTATUS[0]}" -ne 0 ]; then echo "FAILED: $1"; exit 1; fi
}
: > $O/r2_pytest.log
step "single-job fused kernel after the body refactor" 120 "consolidate or batcher_seal or batch_merge"
step "multi-job seal" 90 "seal_many"
step "deferred merges (spine)" 120 "spine"
step "chained probes" 90 "half_join_many"
step "fused first stage" 90 "delta_first_stage"
step "reduce corrections without the sort launch" 120 "reduce"
step "joins (bulk single-pass probe)" 180 "join"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $O/r2_pytest.log
line() {
  python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$1', round(d['value'] / 1e6, 1), 'M rows/s', round(d['ms_per_step'], 4), 'ms/step  e2e', round(d['e2e']['value'] / 1e6, 1),
      ' launches', d.get('gpu_launches'), [(k['kernel'], k['share'], k['launches_per_step']) for k in d['roofline']['top_kernels'][:5]])"
}
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $O/r2_bench.err | tee $O/r2_bench.json | line "batched      "
MZGPU_DEFER_MERGES=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $O/r2_bench_nodefer.err | tee $O/r2_bench_nodefer.json | line "no defer     "
timeout 200 python bench_kernels.py --only cfg2 --out $O/r2_k_cfg2.json > /dev/null 2> $O/r2_k_cfg2.err
python -c "
import json
c = json.load(open('$O/r2_k_cfg2.json'))['cases'][0]
print('cfg2', round(c['rows_per_sec'] / 1e6, 1), 'M rows/s', [(k['kernel'], k['ms'], k['frac_of_measured_hbm']) for k in c['kernels'] if 'probe' in k['kernel']])"
tail -3 $O/r2_bench.err
