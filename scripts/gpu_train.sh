timeout -s KILL 200 python scripts/bench_train.py 2000000 300 30 10 1 2>&1 | tail -1
timeout -s KILL 200 python scripts/bench_train.py 2000000 300 30 10 0 2>&1 | tail -1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/bench_train.py 1000000 300 30 3 1 > gpurun_out/train_ncu.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_rm.csv python scripts/bench_train.py 1000000 300 30 3 0 > gpurun_out/train_ncu_rm.log 2>&1
tail -1 gpurun_out/train_ncu.log
