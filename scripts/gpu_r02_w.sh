#!/bin/bash
# round 2, call w: accumulator released before the survivor listing, warp-aggregated appends, wait flavours of the epilogue; launch list of one call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02w_tests.log 2>&1
rc=$?
echo "tests rc=$rc"
tail -15 gpurun_out/r02w_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py > gpurun_out/r02w_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/r02w_bench.log | cut -c1-400
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py > gpurun_out/r02w_plain.log 2>&1 &&
TSCAN_ONLY=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02w_launches.csv python scripts/bench_tscan.py > gpurun_out/r02w_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02w_ncu.log
