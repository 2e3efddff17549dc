# round 2, call O: sweep-group count experiment (NWG) for tc_assign, re-rank kernel with unrolled loads
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_tcassign.py tests/test_gpu_rerank.py -m gpu -x -q > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.log | cut -c1-300
timeout -s KILL 300 python scripts/bench_encode.py 10000000 300 30 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_encode.py 12500000 128 16 3 2>&1 | tail -2 | head -1
timeout -s KILL 300 python scripts/bench_train.py 10000000 300 30 25 1 1
timeout -s KILL 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import gulon_b200 as g
from gulon_b200.synth import Mixture
from gulon_b200.pipeline import RerankPipeline
dev = torch.device('cuda', 0)
D, M, rows, Q, R, k = 1000, 100, 1000000, 10000, 100, 10
mix = Mixture(D, seed=5, device=dev)
X = mix.rows(0, rows)
pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(X[:65536].contiguous()), g.ProductQuantizerConfig(256, M, 2))
codes = pq.encode_dev(X)
pipe = RerankPipeline(pq, codes, X, 0, rows)
q = mix.rows(0, Q, stream_seed=1)
for _ in range(2): pipe.rerank_only(k, R, q)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): pipe.rerank_only(k, R, q)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("rerank_only ms", ms, "GB/s", Q * R * D * 4 / ms / 1e6)
PY
