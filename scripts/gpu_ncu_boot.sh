mkdir -p gpurun_out
CMD="python scripts/sweep_scan.py --rows 1250000 --dim 300 --m 30 --queries 2368 --reps 1 20:0"
timeout 300 $CMD > gpurun_out/r01e_boot_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_scan_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r01e_boot $CMD > gpurun_out/r01e_boot_ncu.log 2>&1
tail -2 gpurun_out/r01e_boot_ncu.log; cat gpurun_out/r01e_boot_plain.log
