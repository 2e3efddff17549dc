"""Resident-data timing of ProductQuantizer.train (k-means codebook training, c3 shape).

usage: python scripts/bench_train.py [rows] [D] [M] [iters] [update_mode: 0 running mean | 1 sum] [fused: 1 | 0]
Prints one JSON line: seconds, Lloyd passes, rows*windows assigned per second.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import gulon_b200 as g
from gulon_b200.synth import Mixture


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    fused = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    g.set_option("train_fused", fused)
    dev = torch.device("cuda", 0)
    X = Mixture(D, device=dev).rows(0, rows)
    pts = g.DevicePoints.from_torch(X)
    seen = []
    cfg = g.ProductQuantizerConfig(256, M, iters, report=seen.append, update_mode=mode)
    for rep in range(2):
        seen.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.ProductQuantizer.train(pts, cfg)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    passes = len(seen)
    print(json.dumps({"rows": rows, "D": D, "M": M, "max_iters": iters, "update_mode": mode, "fused": fused, "seconds": dt,
                      "reports": passes, "GB_per_pass": rows * D * 4 / 1e9,
                      "s_per_iter": dt / max(1, iters + 1)}))


if __name__ == "__main__":
    main()
