timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/bench_train.py 1000000 300 30 3 1 > gpurun_out/train_ncu.log 2>&1
tail -1 gpurun_out/train_ncu.log
