#!/bin/bash
# round 2, call zn: the 256-row boot kernel of the tensor scan: tests, c2 / c4 / c5 timings with both boots
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02zn_tests.log 2>&1
rc=$?
echo "tests rc=$rc"; tail -8 gpurun_out/r02zn_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
TSCAN_ONLY=1 TSCAN_SWEEP=1 timeout 600 python scripts/bench_tscan.py > gpurun_out/r02zn_c2.log 2>&1; cut -c1-230 gpurun_out/r02zn_c2.log
TSCAN_ONLY=1 TSCAN_SWEEP=1 timeout 600 python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02zn_c4.log 2>&1; cut -c1-230 gpurun_out/r02zn_c4.log
TSCAN_ONLY=1 TSCAN_SWEEP=1 timeout 600 python scripts/bench_tscan.py 1000000 1000 100 10000 100 > gpurun_out/r02zn_c5.log 2>&1; cut -c1-230 gpurun_out/r02zn_c5.log
