# round 2, call K: grouped work-list kernel -- tests, then the whole bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_storage.py tests/test_wordvectors.py -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02k_tests.log | cut -c1-400
( time timeout -s KILL 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r02k_bench.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02k_bench.json'))
print(b["value"], b["e2e"]["value"], json.dumps(b.get("grouped_ivf"))[:900])
PY
