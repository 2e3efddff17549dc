#!/bin/bash
# one full ncu capture of the main-stage pruned scan kernel (cfg from $CFG, default 10:32)
mkdir -p gpurun_out
CFG=${CFG:-10:32}
timeout 600 python scripts/sweep_scan.py --queries 2368 --reps 1 $CFG > gpurun_out/r01d_ncu_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pruned_scan_kernel \
  --launch-skip 3 --launch-count 1 -f -o gpurun_out/r01d_pscan python scripts/sweep_scan.py --queries 2368 --reps 1 $CFG \
  > gpurun_out/r01d_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/r01d_ncu.log
tail -3 gpurun_out/r01d_ncu.log
ncu -i gpurun_out/r01d_pscan.ncu-rep --page raw --csv > gpurun_out/r01d_pscan_raw.csv 2>/dev/null
cat gpurun_out/r01d_ncu_plain.log
