#!/bin/bash
# parity of the pruned scan paths, then the pruned-scan knob sweep (c2 shape)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "pruned or overflow or ties" > gpurun_out/r01d_pytest_query.log 2>&1
echo "pytest exit $?" >> gpurun_out/r01d_pytest_query.log
tail -5 gpurun_out/r01d_pytest_query.log
timeout 900 python scripts/sweep_scan.py ${SWEEP:-30:0 20:32 15:32 12:32 10:32 8:32 6:32 5:32 4:32 8:16 8:64 0:32} > gpurun_out/r01d_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/r01d_sweep.log
cat gpurun_out/r01d_sweep.log
