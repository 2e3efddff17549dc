# round 1c: tensor-core encode/assign + fixed-point update.  Every ncu pass runs only after the same
# command has exited 0 without ncu.
CMD="python bench.py --steps 1 --warmup 1 --queries 4736 --no-cpu-baseline --no-recall"
KR='regex:pruned_scan|fused_scan|lut_build|unpack_keys|select_pass|gather_lists|assign|adc_keys|offsets_kernel|rm_|update_|count_diff|gather_rows|tc_prep|absmax|scale_kernel|finalize'
timeout -s KILL 400 $CMD > gpurun_out/r01c_plain.log 2>&1 && timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$KR" -c 600 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/r01c_ncu1.log 2>&1
E="python scripts/bench_encode.py 1000000 300 30 1"
timeout -s KILL 200 $E > gpurun_out/r01c_enc_plain.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_assign_kernel.*unsigned" -s 1 -c 1 -f -o gpurun_out/prof_tc $E > gpurun_out/r01c_ncu2.log 2>&1
T="python scripts/bench_train.py 1000000 300 30 2 1"
timeout -s KILL 200 $T > gpurun_out/r01c_train_plain.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:update_fixed -s 1 -c 1 -f -o gpurun_out/prof_upd $T > gpurun_out/r01c_ncu3.log 2>&1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$KR" -c 600 --csv --log-file gpurun_out/r01c_train_launches.csv $T > gpurun_out/r01c_ncu4.log 2>&1
tail -1 gpurun_out/r01c_plain.log | cut -c1-300; tail -2 gpurun_out/r01c_enc_plain.log; tail -1 gpurun_out/r01c_train_plain.log; tail -1 gpurun_out/r01c_ncu2.log gpurun_out/r01c_ncu3.log
