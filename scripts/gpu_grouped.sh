timeout -s KILL 300 python -m pytest tests/test_gpu_grouped.py -q -x > gpurun_out/grouped_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/grouped_pytest.log | cut -c1-400
