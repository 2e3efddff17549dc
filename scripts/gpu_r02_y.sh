#!/bin/bash
# round 2, call y: the c4 shard shape (12.5M x 128, m = 16) under the tensor scan: sweep of the kernel forms, ncu of the main stage; c1 shape
mkdir -p gpurun_out
TSCAN_SWEEP=1 timeout 900 python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02y_c4.log 2>&1
echo "c4 rc=$?"
cat gpurun_out/r02y_c4.log | cut -c1-330
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py 12500000 128 16 25000 10 > gpurun_out/r02y_plain.log 2>&1 &&
TSCAN_ONLY=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:filter2_kernel -s 4 -c 1 -o gpurun_out/r02y_filter2_c4 python scripts/bench_tscan.py 12500000 128 16 25000 10 > gpurun_out/r02y_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python scripts/bench_tscan.py 1000000 100 10 10000 10 > gpurun_out/r02y_c1.log 2>&1
echo "c1 rc=$?"
cat gpurun_out/r02y_c1.log | cut -c1-330
