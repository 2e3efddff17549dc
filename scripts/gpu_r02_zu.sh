#!/bin/bash
# round 2, call zu: ncu --set full of the final filter at the c4 shard shape (KP = 144, last stage) and of the streamed-B form at the c5 shape
mkdir -p gpurun_out
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02zu_c4_plain.log 2>&1 &&
TSCAN_ONLY=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:filter2_kernel -s 12 -c 2 -o gpurun_out/r02zu_filter2_c4_final python scripts/bench_tscan.py 12500000 128 16 100000 10 > gpurun_out/r02zu_c4_ncu.log 2>&1
echo "c4 ncu rc=$?"
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py 1000000 1000 100 10000 100 > gpurun_out/r02zu_c5_plain.log 2>&1 &&
TSCAN_ONLY=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:filter2_kernel -s 11 -c 1 -o gpurun_out/r02zu_filter2_c5_final python scripts/bench_tscan.py 1000000 1000 100 10000 100 > gpurun_out/r02zu_c5_ncu.log 2>&1
echo "c5 ncu rc=$?"
