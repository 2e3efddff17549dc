CMD="python scripts/bench_train.py 1000000 300 30 2 1"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:update_fixed -s 1 -c 1 -f -o gpurun_out/prof_upd $CMD > gpurun_out/ncu_upd.log 2>&1
tail -2 gpurun_out/ncu_upd.log
