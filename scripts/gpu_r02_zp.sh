#!/bin/bash
# round 2, call zp: one polling lane in the MMA warps (tc_assign and the tensor scan): tests, encode / training / scan timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_tscan.py tests/test_gpu_update_fixed.py -x -q > gpurun_out/r02zp_tests.log 2>&1
rc=$?
echo "tests rc=$rc"; tail -5 gpurun_out/r02zp_tests.log | cut -c1-400
timeout 300 python scripts/bench_encode.py 2000000 300 30 5 > gpurun_out/r02zp_encode.log 2>&1; grep tensor gpurun_out/r02zp_encode.log | cut -c1-200
timeout 300 python scripts/bench_train.py 2000000 300 30 6 1 1 > gpurun_out/r02zp_train.log 2>&1; tail -1 gpurun_out/r02zp_train.log | cut -c1-300
timeout 300 python scripts/bench_train.py 2000000 300 30 6 1 1 >> gpurun_out/r02zp_train.log 2>&1; tail -1 gpurun_out/r02zp_train.log | cut -c1-300
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py > gpurun_out/r02zp_c2.log 2>&1; cut -c1-160 gpurun_out/r02zp_c2.log
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02zp_c4.log 2>&1; cut -c1-160 gpurun_out/r02zp_c4.log
