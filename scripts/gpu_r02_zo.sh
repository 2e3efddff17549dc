#!/bin/bash
# round 2, call zo: tc_assign with a lean converged-warp MMA issue loop: assignment / encode / training tests, encode and training timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_parity.py tests/test_gpu_update_fixed.py -x -q -k "not query" > gpurun_out/r02zo_tests.log 2>&1
rc=$?
echo "tests rc=$rc"; tail -5 gpurun_out/r02zo_tests.log | cut -c1-400
timeout 300 python scripts/bench_encode.py 2000000 300 30 5 > gpurun_out/r02zo_encode.log 2>&1; cat gpurun_out/r02zo_encode.log | cut -c1-260
timeout 300 python scripts/bench_encode.py 2000000 128 16 5 > gpurun_out/r02zo_encode128.log 2>&1; cat gpurun_out/r02zo_encode128.log | cut -c1-260
timeout 300 python scripts/bench_train.py 2000000 300 30 6 1 1 > gpurun_out/r02zo_train.log 2>&1; tail -2 gpurun_out/r02zo_train.log | cut -c1-300
