#!/bin/bash
# round 2, call zj: ncu evidence of the FINAL tensor scan at the bench's launch shapes: --set full of the five stage launches of one
# 100 000-query call at c2 (the second call: index decoded, buffers allocated), and the launch list of the same command
mkdir -p gpurun_out
TSCAN_ONLY=1 timeout 600 python scripts/bench_tscan.py > gpurun_out/r02zj_plain.log 2>&1 &&
TSCAN_ONLY=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:filter2_kernel -s 5 -c 5 -o gpurun_out/r02zj_filter2_final python scripts/bench_tscan.py > gpurun_out/r02zj_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02zj_ncu.log
TSCAN_ONLY=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02zj_launches.csv python scripts/bench_tscan.py > gpurun_out/r02zj_ncu2.log 2>&1
echo "ncu2 rc=$?"
