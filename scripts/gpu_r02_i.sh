# round 2, call I: the whole GPU suite + smoke + the default bench (state to commit)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02i_pytest.log | cut -c1-400
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout -s KILL 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench.err
