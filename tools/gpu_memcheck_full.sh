#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_full.log \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x --deselect tests/test_parity_gpu.py::test_split_draws_continue_ids_and_order --deselect tests/test_parity_gpu.py::test_coarse_list_overflow_falls_back_on_the_device > gpurun_out/memcheck_full_pytest.log 2>&1
echo "memcheck exit $?"; tail -4 gpurun_out/memcheck_full_pytest.log | cut -c1-300; tail -4 gpurun_out/memcheck_full.log | cut -c1-300
