#!/bin/bash
# round 2, call s: tensor scan with L2-sized row splits: tests, sweep of the split size / stage ratio / boot rows, ncu of the filter
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02s_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/r02s_tests.log
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py > gpurun_out/r02s_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/r02s_bench.log | cut -c1-400
timeout 600 python scripts/bench_tscan.py 10000000 300 30 25000 10 > gpurun_out/r02s_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o gpurun_out/r02s_filter python scripts/bench_tscan.py 10000000 300 30 25000 10 > gpurun_out/r02s_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02s_ncu.log
