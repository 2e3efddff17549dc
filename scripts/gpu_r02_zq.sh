#!/bin/bash
# round 2, call zq: tc_assign with three sweep groups and 32-column TMEM loads: tests, encode and training timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_update_fixed.py -x -q > gpurun_out/r02zq_tests.log 2>&1
rc=$?
echo "tests rc=$rc"; tail -4 gpurun_out/r02zq_tests.log | cut -c1-400
timeout 300 python scripts/bench_encode.py 2000000 300 30 5 > gpurun_out/r02zq_encode.log 2>&1; grep tensor gpurun_out/r02zq_encode.log | cut -c1-200
timeout 300 python scripts/bench_encode.py 2000000 128 16 5 > gpurun_out/r02zq_encode128.log 2>&1; grep tensor gpurun_out/r02zq_encode128.log | cut -c1-200
timeout 300 python scripts/bench_train.py 2000000 300 30 6 1 1 > gpurun_out/r02zq_train.log 2>&1; tail -1 gpurun_out/r02zq_train.log | cut -c1-300
timeout 300 python scripts/bench_train.py 2000000 300 30 6 1 1 >> gpurun_out/r02zq_train.log 2>&1; tail -1 gpurun_out/r02zq_train.log | cut -c1-300
