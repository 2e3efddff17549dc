# round 2, call F: fused-fp32 prefilter in the tensor-core assignment -- parity, encode rate, c3
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_update_fixed.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02f_tests.log | cut -c1-400
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "assign or encode or train or kmeans or pq or c1 or c2 or c4 or c5" > gpurun_out/r02f_tests2.log 2>&1; echo "tests2 rc=$?"; tail -5 gpurun_out/r02f_tests2.log | cut -c1-400
timeout -s KILL 300 python scripts/bench_encode.py 10000000 300 30 2>&1 | tail -3 | tee gpurun_out/r02f_encode.txt
timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 1 2>&1 | tee gpurun_out/r02f_c3.txt
