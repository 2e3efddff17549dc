#!/bin/bash
# round 2, call zl: the streamed-B form of the pair filter (D > 316: c5): tensor-scan tests, c5-shaped timing against the pruned scan
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tscan.py -x -q > gpurun_out/r02zl_tests.log 2>&1
rc=$?
echo "tests rc=$rc"; tail -12 gpurun_out/r02zl_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python scripts/bench_tscan.py 1000000 1000 100 10000 100 > gpurun_out/r02zl_c5.log 2>&1; cut -c1-420 gpurun_out/r02zl_c5.log
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py > gpurun_out/r02zl_c2.log 2>&1; cut -c1-200 gpurun_out/r02zl_c2.log
