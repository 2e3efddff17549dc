GULON_TC_DEBUG=1 timeout -s KILL 600 python -m pytest tests/test_gpu_tcassign.py -q > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/tc_pytest.log
CMD="python scripts/bench_encode.py 1000000 300 30 1"
timeout -s KILL 300 $CMD > gpurun_out/tc_bench.log 2>&1 && timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc_assign -s 2 -c 1 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
tail -4 gpurun_out/tc_bench.log; tail -3 gpurun_out/ncu_tc.log
