"""Times PQIndex.batchQuery at a BASELINE shape under the tensor scan and the pruned scan on one resident index
and compares the answers bit for bit.  python scripts/bench_tscan.py [rows D M queries k]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import gulon_b200 as g
from gulon_b200 import _native as N
from gulon_b200.synth import Mixture


def main():
    a = [int(x) for x in sys.argv[1:6]] + [10_000_000, 300, 30, 100_000, 10][len(sys.argv) - 1:]
    rows, D, M, nq, k = a
    dev = torch.device("cuda", 0)
    mix = Mixture(D, device=dev)
    xt = mix.rows(0, 262144)
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, 8))
    del xt
    stride = (rows + 15) // 16 * 16
    codes = torch.zeros((M, stride), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    CH = 1 << 20
    for r0 in range(0, rows, CH):
        n = min(CH, rows - r0)
        x = mix.rows(r0, r0 + n)
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST, codes.data_ptr() + r0, stride, st))
        torch.cuda.synchronize()
        del x
    ix = g.PQIndex.from_device_codes(pq, codes, rows)
    Q = mix.rows(0, nq, stream_seed=1)
    res = {}
    import os
    only = os.environ.get("TSCAN_ONLY")
    for name, impl in ((("tensor", g.SCAN_TENSOR),) if only else (("tensor", g.SCAN_TENSOR), ("pruned", g.SCAN_PRUNED))):
        g.set_option("scan_impl", impl)
        g.set_option("profile", 1)
        t0 = time.perf_counter()
        out = ix.batch_query_dev(k, Q)
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        reps = 1 if only else (3 if name == "tensor" else 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = ix.batch_query_dev(k, Q)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = (out[0].cpu().numpy(), out[1].cpu().numpy())
        line = {"impl": name, "ms": ms, "qps": nq / ms * 1e3, "first_call_s": first}
        if name == "tensor":
            c = {n: N.counter("tscan_" + n) for n in ("kernel_ns", "kernel_launches", "tiles", "slow_paths", "survivors",
                                                     "candidates", "pairs", "fallbacks", "batches", "stages")}
            line.update(c)
            KP = (D + 4 + 15) // 16 * 16
            flop = 2.0 * 128 * 256 * KP * c["tiles"]
            line["filter_tflops"] = flop / max(c["kernel_ns"], 1) * 1e-3
            line["survivor_rate"] = c["survivors"] / max(c["pairs"], 1)
        print(line, flush=True)
        g.set_option("profile", 0)
    if os.environ.get("TSCAN_SWEEP"):
        g.set_option("scan_impl", g.SCAN_TENSOR)
        for name, vals in (("tensor_boot_rows", [8192, 0]), ("tensor_stage_ratio", [3, 6, 8, 0]),):
            for v in vals:
                g.set_option(name, v)
                g.set_option("profile", 1)
                ix.batch_query_dev(k, Q)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    out = ix.batch_query_dev(k, Q)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                same = np.array_equal(out[0].cpu().numpy(), res["tensor"][0])
                if name == "tensor_epi_wait" and v >= 8:
                    N.lib().gulon_set_option(b"tensor_epi_wait", 2)
                print({name: v, "ms": ms, "qps": nq / ms * 1e3, "filter_ms": N.counter("tscan_kernel_ns") / 3e6, "other_ms": ms - N.counter("tscan_kernel_ns") / 3e6,
                       "survivors": N.counter("tscan_survivors") // 3, "same": bool(same)}, flush=True)
                g.set_option("profile", 0)
    g.set_option("scan_impl", g.SCAN_AUTO)
    if only:
        return
    same = np.array_equal(res["tensor"][0], res["pruned"][0]) and np.array_equal(res["tensor"][1].view(np.uint32), res["pruned"][1].view(np.uint32))
    print({"tensor_equals_pruned": bool(same)}, flush=True)


main()
