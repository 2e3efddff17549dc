# round 2, call E: fused Lloyd pass -- parity, then c3 timing fused / three-pass, then the launch list
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_update_fixed.py tests/test_gpu_tcassign.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02e_tests.log | cut -c1-400
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train or kmeans or sharded or compute" > gpurun_out/r02e_tests2.log 2>&1; echo "tests2 rc=$?"; tail -5 gpurun_out/r02e_tests2.log | cut -c1-400
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/r02e_clocks.csv &
SMI=$!
for f in 1 0; do timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 $f; done 2>&1 | tee gpurun_out/r02e_c3.txt
timeout -s KILL 300 python scripts/bench_train.py 5000000 300 30 25 1 1 2>&1 | tee -a gpurun_out/r02e_c3.txt
kill $SMI
sort gpurun_out/r02e_clocks.csv | uniq -c | sort -rn | head -8
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_train_launches.csv python scripts/bench_train.py 2000000 300 30 6 1 1 > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
