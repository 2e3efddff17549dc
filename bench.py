#!/usr/bin/env python
"""bench.py -- headline benchmark of the PQ hot path (BASELINE.json: queries/sec at fixed
recall@10 on a 10M x 300-d PQ index; encode vectors/sec as a secondary figure).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference ...                    (the reference algorithm on host cores)

A "step" is one PQIndex.batchQuery of the whole query batch (100k queries, top-10) over the
index: ADC lookup-table build + uint8 code scan + top-k (+ all-gather merge when sharded).
`value` is timed with the queries and the index resident in HBM; `e2e` goes through the host
C-ABI call (gulon_pq_query) with host query/result buffers.  Data are synthetic
(word-embedding-shaped Gaussian mixture, gulon_b200/synth.py), codebooks are trained by the
library itself.  Between timed steps nothing is cached on purpose: the code planes (300 MB) and the
lookup tables of a query batch (> 70 MB per 2368-query tile) exceed the 126 MB L2 together, see
config.l2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec at fixed recall@10 on 10Mx300-d PQ index"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=300)
    ap.add_argument("--m", "--subquantizers", dest="m", type=int, default=30)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--train-rows", type=int, default=262_144)
    ap.add_argument("--train-iters", type=int, default=8)
    ap.add_argument("--recall-queries", type=int, default=200)
    ap.add_argument("--cpu-queries", type=int, default=0,
                    help="queries of the CPU sample (0 = about 10 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-recall", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    ap.add_argument("--row-shards", type=int, default=0,
                    help="row shards R of the code planes (0 = gulon_b200.sharded.shard_plan); the other "
                         "factor of the world size splits the query batch")
    return ap.parse_args()


def workload(a):
    return {"workload": "configs[1]: %dx%d-d word-embedding-shaped synthetic vectors, PQ m=%dx256, "
                        "top-%d query batch of %d" % (a.rows, a.dim, a.m, a.k, a.queries),
            "rows": a.rows, "dim": a.dim, "m": a.m, "clusters": 256, "k": a.k, "queries": a.queries}


# ---- clocks -------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- reference arm: the reference algorithm (C restatement; no JVM in the image) on host cores ---
def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as o
    o.build()
    T = o.num_threads()
    rng = np.random.default_rng(20261018)
    M, K, D = a.m, 256, a.dim
    dmax = -(-D // M)
    cb = rng.normal(size=(M, K, dmax)).astype(np.float32)
    codes = rng.integers(0, K, (M, a.rows), dtype=np.uint8)
    nq = a.cpu_queries or 16 * T      # ~2 s of work per step on 16 cores at the c2 shape
    Q = rng.normal(size=(nq, D)).astype(np.float32)

    def step():
        # Index.prepareQuery + PQIndex.batchQuery; one task per query as G/Tests.scala:109-121
        return o.pq_query(Q, cb, codes, a.k, topk_mode=o.TOPK_LITERAL, nthreads=T)

    for _ in range(min(a.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    v = nq * a.steps / dt
    sample = "%d queries x full %d-row index per step, literal TopKHeap" % (nq, a.rows)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload(a),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": T, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/); the "
                                 "reference is Scala/JVM and no JDK exists in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    return 0


# ---- own arm ---------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    import gulon_b200 as g
    from gulon_b200 import _native as N
    from gulon_b200.sharded import ShardedPQIndex, shard_bounds, shard_plan
    from gulon_b200.synth import Mixture

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if g.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: gulon_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    N.check(N.lib().gulon_set_device(local))
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for kv in a.opt:
        name, val = kv.split("=")
        g.set_option(name, int(val))
    D, M, K, k, Q = a.dim, a.m, 256, a.k, a.queries
    mix = Mixture(D, device=dev)

    # codebooks: trained by the library on the first rows of the data set (same on every rank)
    xt = mix.rows(0, min(a.train_rows, a.rows))
    t0 = time.perf_counter()
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt),
                                  g.ProductQuantizerConfig(K, M, a.train_iters))
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    del xt

    # (R row shards) x (C query groups) over the ranks: gulon_b200.sharded.shard_plan
    n_shards, n_groups = shard_plan(a.rows, world) if a.row_shards <= 0 else \
        (a.row_shards, world // a.row_shards)
    # this rank's row shard of the database, encoded chunk by chunk (rows independent: no exchange)
    lo, hi = shard_bounds(a.rows, n_shards)[rank % n_shards]
    n_local = hi - lo
    stride = (max(n_local, 1) + 15) // 16 * 16
    codes = torch.zeros((M, stride), dtype=torch.uint8, device=dev)
    keep_x = world == 1 and not a.no_recall
    X = torch.empty((n_local, D), dtype=torch.float32, device=dev) if keep_x else None
    CH = 1 << 20
    enc_ns, enc_rows = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    for r0 in range(0, n_local, CH):
        n = min(CH, n_local - r0)
        x = mix.rows(lo + r0, lo + r0 + n, out=X[r0:r0 + n] if keep_x else None)
        e0.record()
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST,
                                            codes.data_ptr() + r0, stride, st))
        e1.record()
        torch.cuda.synchronize()
        if r0 > 0 or n_local <= CH:          # first chunk is the warm-up
            enc_ns += e0.elapsed_time(e1) * 1e6
            enc_rows += n
        del x
    # candidate statistics of the tensor-core encode (one untimed chunk, profile counters on)
    enc_stats = None
    if n_local > 0:
        n = min(CH, n_local)
        x = X[:n] if keep_x else mix.rows(lo, lo + n)
        g.set_option("profile", 1)
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST,
                                            codes.data_ptr(), stride, st))
        torch.cuda.synchronize()
        tc_rows = N.counter("assign_tc_rows")
        if tc_rows:
            enc_stats = {"kernel": "tca::tc_assign_kernel (tcgen05 filter + exact fp32 recheck)",
                         "candidate_chunks_per_row_window": N.counter("assign_tc_pairs") / tc_rows}
        else:
            enc_stats = {"kernel": "assign_exact_kernel (CUDA cores)"}
        g.set_option("profile", 0)
        del x
    ix = g.PQIndex.from_device_codes(pq, codes, n_local)
    sh = ShardedPQIndex(ix, lo, plan=(n_shards, n_groups)) if world > 1 else ShardedPQIndex(ix, lo)
    queries = mix.rows(0, Q, stream_seed=1)

    def step_dev():
        return sh.batch_query(k, queries) if world > 1 else ix.batch_query_dev(k, queries)

    # warm-up, then the timed region
    for _ in range(a.warmup):
        step_dev()
    barrier()
    launches0 = g.kernel_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0.record()
    for _ in range(a.steps):
        out = step_dev()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    launches = g.kernel_launches() - launches0
    # the same K steps once more with the library's kernel timers and counters on (they add an event pair
    # per kernel and a host sync per scan stage, so they stay out of the timed region above); the
    # roofline's kernel time, launch count and survivor statistics come from this pass
    g.set_option("profile", 1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(a.steps):
        step_dev()
    p1.record()
    barrier()
    prof_ms = p0.elapsed_time(p1)
    scan_ns = N.counter("scan_kernel_ns")
    scan_launches = N.counter("scan_kernel_launches")
    pscan_ns = N.counter("pscan_kernel_ns")
    pscan_launches = N.counter("pscan_kernel_launches")
    pstats = {nm: N.counter("pscan_" + nm) for nm in ("survivors", "candidates", "slow_items", "pairs")}
    lb_quantizers = N.counter("pscan_lb_quantizers")
    g.set_option("profile", 0)
    value = Q * a.steps / (ms * 1e-3)

    # e2e: the host-facing call, host query buffer in, host (ids, distances) out, every step
    q_host = torch.empty((Q, D), dtype=torch.float32, pin_memory=True)
    q_host.copy_(queries)
    torch.cuda.synchronize()
    q_np = q_host.numpy()

    def step_e2e():
        if world == 1:
            r = ix.batch_query(k, q_np)          # gulon_pq_query: H2D + scan + D2H inside
            return r.keys, r.values
        ids, ds, _ = sh.batch_query(k, q_host)   # each rank copies its slice of the pinned batch to HBM
        return ids.cpu(), ds.cpu()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        res = step_e2e()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_value = Q * a.steps / (e2e_ms * 1e-3)

    # results of both paths must agree (same kernels, same inputs)
    ids_dev = out[0].cpu().numpy()
    same = bool(np.array_equal(ids_dev, np.asarray(res[0])))

    extra = {}
    # recall@10 per G/Tests.scala:18-41 on a query sample (single GPU: raw vectors are resident)
    if keep_x and rank == 0:
        from gulon_b200.index import rerank
        R = min(a.recall_queries, Q)
        qs = q_np[:R]
        pts = g.DevicePoints.from_torch(X)
        gt = g.exact_nearest_neighbours(pts, qs, k)
        ex = rerank(pts, qs, ids_dev[:R], k)            # exact distances of the returned keys
        kth = gt.values[:, k - 1]
        tp = [(ex.values[i][ex.keys[i] >= 0] <= kth[i]).sum() for i in range(R)]
        extra["recall_at_10"] = {"mean": float(np.mean(tp)) / k, "queries": R,
                                 "definition": "G/Tests.scala:18-41, eps=0"}
        del pts

    # roofline of the dominant kernel (fused scan): algorithmic bytes = one pass over this rank's
    # code planes per tile of Qt = 4 queries whose lookup tables share shared memory
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roof = None
    use_p = pscan_launches > 0 and pscan_ns >= scan_ns
    k_ns, k_launches = (pscan_ns, pscan_launches) if use_p else (scan_ns, scan_launches)
    if k_launches > 0:
        # algorithmic bytes: one pass over the scanned code planes per tile of Qt
        # queries whose tables share shared memory (Qt = 8 pruned kernel, 4 exact kernel)
        ml = lb_quantizers or M                  # quantizers summed by the lower bound (main stage)
        fb = 8 if 127 // ml >= 3 else 16         # the library's automatic field width
        QT = (N.counter("pscan_qt") or 128 // fb) if use_p else 4
        rows_scanned = (pstats["pairs"] / (a.steps * Q)) if use_p else n_local
        alg_bytes = a.steps * -(-Q // QT) * rows_scanned * M / k_launches
        sec = k_ns * 1e-9 / k_launches
        ach = alg_bytes / sec / 1e9
        # table reads actually issued: the bound pass reads ml of the M quantizers
        gathers = a.steps * Q * rows_scanned * (ml if use_p else M) / (k_ns * 1e-9)
        clk_mhz = (clk or {}).get("sm_mhz") or 1965.0
        smem_peak_bytes = 148 * 128 * clk_mhz * 1e6
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "kernel": "pruned_scan_kernel" if use_p else "fused_scan_kernel",
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_seconds": sec,
                "launches": k_launches, "query_tile": QT,
                "lower_bound_quantizers": ml if use_p else None,
                "note": "algorithmic bytes count all M code planes per pass (SURVEY 8d); the pruned kernel "
                        "streams only the planes its lower bound sums, so `achieved` is work done per "
                        "second, not bytes moved" if use_p else None,
                "kernel_share_of_step": k_ns * 1e-6 / prof_ms,
                "other_scan_kernel_share_of_step": (scan_ns if use_p else pscan_ns) * 1e-6 / prof_ms,
                "timed_in": "a repeat of the K steps with the kernel timers on (%.1f ms per step; the "
                            "timed region runs without them)" % (prof_ms / a.steps),
                "smem_gather": {"bytes_per_entry": (fb // 8) if use_p else 4,
                                "achieved_GBps": gathers * ((fb // 8) if use_p else 4) / 1e9,
                                "peak_GBps": smem_peak_bytes / 1e9,
                                "frac": gathers * ((fb // 8) if use_p else 4) / smem_peak_bytes,
                                "note": "table reads from shared memory: the binding resource "
                                        "(128 B/clk/SM crossbar)"}}
        if use_p and world == 1 and (a.rows, a.dim, a.m, a.k) == (10_000_000, 300, 30, 10):
            # one `ncu --set full` capture of the main-stage launch of this workload (2368 queries x 9.6M
            # rows, bound over 15 quantizers): profiles/r01e_pruned_scan_lb15_ncu_full.csv.  That launch's
            # algorithmic bytes are 148 tiles x 9.6M rows x 30 planes = 42.8 GB; it streams 15 planes.
            roof["traffic"] = 16.401e9 + 0.006e9
            roof["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the main-stage launch, "
                                      "profiles/r01e_pruned_scan_lb15_ncu_full.csv (not measured in this run)")
        if use_p and pstats["pairs"]:
            roof["pruning"] = {"survivor_rate": pstats["survivors"] / pstats["pairs"],
                               "list_candidates": pstats["candidates"],
                               "slow_path_items": pstats["slow_items"]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as o
        T = o.num_threads()
        # ~10 s of host work at the c2 shape (about 100 queries/s on 16 cores); every sampled query's ids and
        # distances are also compared with the GPU's
        nq = min(Q, a.cpu_queries or max(64, int(64 * T * (3e8 / max(1.0, float(n_local) * M)))))
        cb = pq.codebook()
        hc = codes[:, :n_local].cpu().numpy()
        t0 = time.perf_counter()
        ci, cd, _ = o.pq_query(q_np[:nq], cb, hc, k, topk_mode=o.TOPK_CANONICAL, nthreads=T)
        dt = time.perf_counter() - t0
        cpu = {"value": nq / dt, "unit": UNIT, "cores": T, "kind": "port",
               "sample": "%d of the %d queries x the full %d-row index" % (nq, Q, n_local),
               "matches_gpu": bool(np.array_equal(ci, ids_dev[:nq]) and
                                   np.array_equal(cd, out[1][:nq].cpu().numpy()))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload(a), sharding=("%d row shard(s) of the code planes x %d query group(s) "
                           "(gulon_b200.sharded.shard_plan: shards keep >= 8M rows); all-gather + "
                           "(distance,id) merge inside a group, all-gather of the slices across groups"
                           % (n_shards, n_groups)) if world > 1 else "single GPU",
                           l2="inputs larger than L2: 300 MB code planes + per-tile lookup tables "
                              "> 126 MB; no flush needed"),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            # whole job: every query group copies its slice of the batch to each of its R row shards;
            # every rank reads the assembled (ids, distances) back
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_shards * Q * D * 4,
                    "d2h_bytes_per_step": world * Q * k * 8 + (Q * 4 if world == 1 else 0),
                    "ms_per_step": e2e_ms / a.steps, "ids_equal_device_path": same},
            "gpu_launches": launches,
            "encode": dict({"value": enc_rows / (enc_ns * 1e-9) if enc_ns else None, "unit": "vectors/s",
                            "rows": enc_rows, "resident": True,
                            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                                         "achieved": enc_rows * (D * 4 + M) / enc_ns if enc_ns else None,
                                         "frac": enc_rows * (D * 4 + M) / enc_ns / peak if enc_ns else None,
                                         "algorithmic_bytes_per_vector": D * 4 + M}},
                           **(enc_stats or {})),
            "train": {"rows": min(a.train_rows, a.rows), "iters": a.train_iters, "seconds": train_s},
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
