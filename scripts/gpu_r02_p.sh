mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tcassign.py -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 300 python scripts/bench_encode.py 10000000 300 30 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_encode.py 12500000 128 16 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 1
