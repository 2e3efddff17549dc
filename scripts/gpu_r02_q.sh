# round 2, call Q: lower-bound scan for wide indexes (K > 256)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; echo "wide rc=$?"; tail -25 gpurun_out/r02q_tests.log | cut -c1-300
timeout -s KILL 300 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import gulon_b200 as g
from gulon_b200.synth import Mixture
dev = torch.device('cuda', 0)
D, M, K, rows, Q, k = 128, 16, 1024, 4_000_000, 4736, 10
mix = Mixture(D, seed=7, device=dev)
X = mix.rows(0, rows)
pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(X[:131072].contiguous()), g.ProductQuantizerConfig(K, M, 2))
codes = pq.encode_dev(X)
ix = g.PQIndex.from_device_codes(pq, codes, rows)
q = mix.rows(0, Q, stream_seed=1)
for impl, name in ((g.SCAN_AUTO, "lower-bound scan"), (g.SCAN_SIMPLE, "plain tables + selection")):
    g.set_option("scan_impl", impl)
    nqq = Q if impl == g.SCAN_AUTO else 128
    for _ in range(2): r = ix.batch_query_dev(k, q[:nqq])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2): r = ix.batch_query_dev(k, q[:nqq])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("%s: K=%d, %d x %d-d rows, m=%d: %.0f queries/s (lb quantizers %d)" % (name, K, rows, D, M, nqq / dt, g._native.counter("pscan_lb_quantizers")))
    if impl == g.SCAN_AUTO: keep = (r[0][:128].cpu().numpy(), r[1][:128].cpu().numpy())
    else: print("same answers:", np.array_equal(keep[0], r[0].cpu().numpy()) and np.array_equal(keep[1], r[1].cpu().numpy()))
g.set_option("scan_impl", g.SCAN_AUTO)
PY
