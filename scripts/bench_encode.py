"""Resident-data timing of ProductQuantizer.encode: exact CUDA-core kernel vs tensor-core path.

usage: python scripts/bench_encode.py [rows] [D] [M] [reps]
Prints one JSON line per implementation: vectors/s, algorithmic GB/s (N*(D*4+M) bytes) and the
fraction of the measured HBM peak, plus candidate chunks per (row, window) for the tensor path.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import gulon_b200 as g
from gulon_b200 import _native as N
from gulon_b200.synth import Mixture


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    peak = 6548.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    mix = Mixture(D, device=dev)
    xt = mix.rows(0, 200_000)
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, 4))
    X = mix.rows(0, rows)
    stride = (rows + 15) // 16 * 16
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, impl in (("exact", 1), ("tensor", 2)):
        g.set_option("assign_impl", impl)
        codes = torch.zeros((M, stride), dtype=torch.uint8, device=dev)
        for _ in range(2):
            N.check(N.lib().gulon_pq_encode_dev(pq.handle, X.data_ptr(), rows, D, N.TIE_LOWEST,
                                                codes.data_ptr(), stride, st))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            N.check(N.lib().gulon_pq_encode_dev(pq.handle, X.data_ptr(), rows, D, N.TIE_LOWEST,
                                                codes.data_ptr(), stride, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = rows * (D * 4 + M) / (ms * 1e-3) / 1e9
        line = {"impl": name, "rows": rows, "D": D, "M": M, "ms": ms, "vectors_per_s": rows / (ms * 1e-3),
                "GBps": gbs, "hbm_frac": gbs / peak}
        if impl == 2:
            g.set_option("profile", 1)
            N.check(N.lib().gulon_pq_encode_dev(pq.handle, X.data_ptr(), rows, D, N.TIE_LOWEST,
                                                codes.data_ptr(), stride, st))
            torch.cuda.synchronize()
            line["chunks_per_row_window"] = N.counter("assign_tc_pairs") / max(1, N.counter("assign_tc_rows"))
            g.set_option("profile", 0)
        out[name] = codes
        print(json.dumps(line), flush=True)
    print(json.dumps({"codes_identical": bool(torch.equal(out["exact"], out["tensor"]))}))
    g.set_option("assign_impl", 0)


if __name__ == "__main__":
    main()
