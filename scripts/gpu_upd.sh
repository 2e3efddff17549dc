timeout -s KILL 240 python -m pytest tests/test_gpu_update_fixed.py -q -x > gpurun_out/upd_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/upd_pytest.log | cut -c1-1500
timeout -s KILL 200 python scripts/bench_train.py 2000000 300 30 10 1 2>&1 | tail -1
