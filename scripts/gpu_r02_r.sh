#!/bin/bash
# round 2, call r: first run of the tensor scan (tests, then a c2-shaped timing against the pruned scan)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02r_tests.log 2>&1
echo "tests rc=$?"
tail -30 gpurun_out/r02r_tests.log
timeout 900 python scripts/bench_tscan.py > gpurun_out/r02r_bench.log 2>&1
echo "bench rc=$?"
tail -8 gpurun_out/r02r_bench.log
