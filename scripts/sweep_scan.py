"""Sweep of the pruned scan's knobs on one resident index (GPU box only; not a bench value).
usage: python scripts/sweep_scan.py [--rows N --dim D --m M --queries Q] cfg cfg ...   cfg = lb:stage_div[:boot[:query_batch]]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=300)
    ap.add_argument("--m", type=int, default=30)
    ap.add_argument("--queries", type=int, default=23680)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--centres", type=int, default=4096)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    ap.add_argument("cfgs", nargs="*")
    a = ap.parse_args()
    import torch
    import gulon_b200 as g
    from gulon_b200 import _native as N
    from gulon_b200.synth import Mixture
    dev = torch.device("cuda", 0)
    for kv in a.opt:
        name, val = kv.split("=")
        g.set_option(name, int(val))
    D, M, K = a.dim, a.m, 256
    mix = Mixture(D, device=dev, centres=a.centres)
    xt = mix.rows(0, min(262144, a.rows))
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(K, M, 8))
    del xt
    stride = (a.rows + 15) // 16 * 16
    codes = torch.zeros((M, stride), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    CH = 1 << 20
    for r0 in range(0, a.rows, CH):
        n = min(CH, a.rows - r0)
        x = mix.rows(r0, r0 + n)
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST,
                                            codes.data_ptr() + r0, stride, st))
        torch.cuda.synchronize()
        del x
    ix = g.PQIndex.from_device_codes(pq, codes, a.rows)
    queries = mix.rows(0, a.queries, stream_seed=1)
    ref = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for cfg in a.cfgs or ["30:0", "15:32"]:
        parts = [int(v) for v in cfg.split(":")]
        lb, sdiv = parts[0], parts[1]
        boot = parts[2] if len(parts) > 2 else 0
        g.set_option("query_batch", parts[3] if len(parts) > 3 else 0)
        g.set_option("pruned_lb_quantizers", lb)
        g.set_option("pruned_stage_div", sdiv)
        g.set_option("boot_rows", boot)
        ix.batch_query_dev(a.k, queries)
        if lb == 0:
            for _ in range(3):
                ix.batch_query_dev(a.k, queries)
        torch.cuda.synchronize()
        g.set_option("profile", 0)
        e0.record()
        for _ in range(a.reps):
            out = ix.batch_query_dev(a.k, queries)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        g.set_option("profile", 1)
        ix.batch_query_dev(a.k, queries)
        torch.cuda.synchronize()
        st_ = {nm: N.counter("pscan_" + nm) for nm in ("survivors", "candidates", "slow_items", "pairs")}
        pns, pl = N.counter("pscan_kernel_ns"), N.counter("pscan_kernel_launches")
        sns = N.counter("scan_kernel_ns")
        ml = N.counter("pscan_lb_quantizers")
        g.set_option("profile", 0)
        ids = out[0].cpu()
        ds = out[1].cpu()
        if ref is None:
            ref = (ids, ds)
        same = bool(torch.equal(ids, ref[0]) and torch.equal(ds.view(torch.int32), ref[1].view(torch.int32)))
        print("cfg %-12s ML=%-3d qps=%9.0f ms=%8.2f pscan_ms=%8.2f boot_ms=%6.2f launches=%d surv_rate=%.3e cand=%d slow=%d same=%s"
              % (cfg, ml, a.queries / (ms * 1e-3), ms, pns * 1e-6, sns * 1e-6, pl,
                 st_["survivors"] / max(st_["pairs"], 1), st_["candidates"], st_["slow_items"], same),
              flush=True)


if __name__ == "__main__":
    main()
