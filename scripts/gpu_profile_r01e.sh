# round 1e: pruned scan with the subset bound.  Every ncu pass runs only after the same command has
# exited 0 without ncu.  The bound is pinned to the 15 quantizers the default bench run converges to.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --queries 4736 --no-cpu-baseline --no-recall --opt pruned_lb_quantizers=15"
KR='regex:pruned_scan|fused_scan|lut_build|unpack_keys|select_pass|gather_lists|assign|adc_keys|offsets_kernel|qparams|qselect|qlut|rowcodes|normalize|merge_small'
timeout -s KILL 400 $CMD > gpurun_out/r01e_plain.log 2>&1 && timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$KR" -c 900 --csv --log-file gpurun_out/r01e_launches.csv $CMD > gpurun_out/r01e_ncu1.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pruned_scan_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r01e_pscan16 $CMD > gpurun_out/r01e_ncu2.log 2>&1
tail -1 gpurun_out/r01e_plain.log | cut -c1-400; tail -1 gpurun_out/r01e_ncu1.log gpurun_out/r01e_ncu2.log
