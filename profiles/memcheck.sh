#!/bin/bash
# compute-sanitizer memcheck over the two hot kernels on small inputs (run on the GPU box).
# usage: profiles/memcheck.sh > gpurun_out/memcheck.log 2>&1
cd "$(dirname "$0")/.."
for t in "tests/test_gpu_parity.py::test_decompress_all_block_types[dynamic6]" \
         "tests/test_gpu_parity.py::test_parse_fuzz[0]" \
         "tests/test_gpu_parity.py::test_zero_copy_reads_nothing_past_a_registered_buffer"; do
  echo "=== compute-sanitizer --tool memcheck :: $t"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
      python -m pytest "$t" -x -q -m gpu 2>&1 | grep -v "^$" | tail -15
  echo "exit: ${PIPESTATUS[0]}"
done
