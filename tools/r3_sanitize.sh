#!/bin/bash
# compute-sanitizer memcheck over the tests that drive the round's new kernels (merge-path kernels on the side
# stream, peer-memory exchange, block caches):  gpurun --timeout 260 -- 'bash tools/r3_sanitize.sh'
O=gpurun_out; mkdir -p $O
timeout 200 compute-sanitizer --tool memcheck --log-file $O/r02c_sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch_merge or spine_structure or peer_memory or q3_dataflow" 2>&1 | tail -4 | tee $O/r02c_sanitizer_pytest.log
tail -5 $O/r02c_sanitizer_memcheck.log
grep -c "Invalid\|Error" $O/r02c_sanitizer_memcheck.log
