GULON_TC_DEBUG=1 timeout -s KILL 120 python -m pytest tests/test_gpu_tcassign.py -q -x > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tc_pytest.log | cut -c1-1800
timeout -s KILL 120 python scripts/bench_encode.py 2000000 300 30 5 > gpurun_out/tc_bench.log 2>&1; tail -3 gpurun_out/tc_bench.log
