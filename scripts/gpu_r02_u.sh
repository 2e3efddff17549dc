#!/bin/bash
# round 2, call u: lean MMA issue loops (converged warp + elect, 32-bit ring counters): tests, timing sweep, ncu of the pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02u_tests.log 2>&1
rc=$?
echo "tests rc=$rc"
tail -15 gpurun_out/r02u_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py > gpurun_out/r02u_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/r02u_bench.log | cut -c1-400
timeout 600 python scripts/bench_tscan.py 10000000 300 30 25000 10 > gpurun_out/r02u_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:filter2_kernel -s 2 -c 1 -o gpurun_out/r02u_filter2 python scripts/bench_tscan.py 10000000 300 30 25000 10 > gpurun_out/r02u_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02u_ncu.log
