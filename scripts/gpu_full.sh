mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-600
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
