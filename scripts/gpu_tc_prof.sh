CMD="python scripts/bench_encode.py 1000000 300 30 1"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_assign_kernel.*unsigned" -s 1 -c 1 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
