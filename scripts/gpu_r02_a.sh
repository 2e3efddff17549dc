# round 2, call A: the sharded C entry point through thread hooks, then the whole GPU suite
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r02a_sharded.log 2>&1; echo "sharded rc=$?"; tail -15 gpurun_out/r02a_sharded.log | cut -c1-400
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02a_pytest.log | cut -c1-600
