# round 2, call J: k-chunked fast scans, persistent encode staging -- edge + parity tests, encode e2e
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_edge.py tests/test_gpu_rerank.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02j_tests.log | cut -c1-400
timeout -s KILL 600 python bench.py --no-recall --steps 1 --warmup 1 --cpu-queries 64 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02j_bench.json'))
print(b["value"], b["encode"]["e2e"], b["train_c3"]["seconds_each"], b["rerank_c5"]["value"])
PY
