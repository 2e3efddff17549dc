# round 2, call D: pruned scan with the TMA slice pipeline -- parity tests, then the headline
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q -k "query or nonfinite or k_" > gpurun_out/r02d_scan.log 2>&1; echo "scan tests rc=$?"; tail -15 gpurun_out/r02d_scan.log | cut -c1-400
timeout -s KILL 600 python bench.py --no-extra-legs --no-recall > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
