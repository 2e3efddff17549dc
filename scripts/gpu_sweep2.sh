#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sweep_scan.py $SWEEP > gpurun_out/r01d_sweep2.log 2>&1
echo "sweep exit $?" >> gpurun_out/r01d_sweep2.log
cat gpurun_out/r01d_sweep2.log
