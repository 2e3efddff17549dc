#!/bin/bash
# round 2, call z: sentinel warp + named barrier for the epilogue wake-up: tests, c4 and c2 sweeps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02zf_tests.log 2>&1
rc=$?
echo "tests rc=$rc"
tail -5 gpurun_out/r02zf_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02zf_c4.log 2>&1
echo "c4 rc=$?"
cat gpurun_out/r02zf_c4.log | cut -c1-250
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py > gpurun_out/r02zf_c2.log 2>&1
echo "c2 rc=$?"
cat gpurun_out/r02zf_c2.log | cut -c1-250
