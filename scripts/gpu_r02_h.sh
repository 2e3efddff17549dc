mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_update_fixed.py -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02h_tests.log | cut -c1-300
timeout -s KILL 300 python scripts/bench_encode.py 10000000 300 30 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_encode.py 1000000 1000 100 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_encode.py 12500000 128 16 3 2>&1 | tail -2 | head -1
for i in 1 2; do timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 1; done
