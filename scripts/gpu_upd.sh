timeout -s KILL 240 python -m pytest tests/test_gpu_update_fixed.py -q -x > gpurun_out/upd_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/upd_pytest.log | cut -c1-1500
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -q -x -k "update or clusters or sharded or pq_train" > gpurun_out/upd_pytest2.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/upd_pytest2.log | cut -c1-800
timeout -s KILL 200 python scripts/bench_train.py 2000000 300 30 10 1 2>&1 | tail -1
timeout -s KILL 200 python scripts/bench_train.py 2000000 300 30 10 0 2>&1 | tail -1
