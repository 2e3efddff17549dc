# round 2, call N: cooperative re-rank kernel, bench with the other-shapes leg
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_recall.py tests/test_gpu_update_fixed.py -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02n_tests.log | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k c5 > gpurun_out/r02n_tests2.log 2>&1; echo "c5 rc=$?"; tail -3 gpurun_out/r02n_tests2.log | cut -c1-300
( time timeout -s KILL 900 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r02n_bench.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02n_bench.json'))
print(b["value"], b["e2e"]["value"], b["roofline"].get("dram_frac"), b["roofline"].get("traffic"))
print("c5", b["rerank_c5"]["value"], b["rerank_c5"]["rerank"], b["rerank_c5"].get("matches_oracle"))
print("shapes", json.dumps(b.get("other_shapes"))[:1200])
PY
