# round 2, call B: new tests (synth twin, rerank pipeline, edge cases, full-size oracle checks), then both bench arms
mkdir -p gpurun_out
nproc
timeout -s KILL 900 python -m pytest tests/test_synth.py tests/test_gpu_rerank.py tests/test_gpu_edge.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02b_new.log 2>&1; echo "new tests rc=$?"; tail -25 gpurun_out/r02b_new.log | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02b_full.log 2>&1; echo "fullsize rc=$?"; tail -25 gpurun_out/r02b_full.log | cut -c1-300
( time timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_ref.json 2> gpurun_out/r02b_ref.err ) 2>&1 | grep real; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02b_ref.json; tail -3 gpurun_out/r02b_ref.err
( time timeout -s KILL 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
