#!/usr/bin/env python
"""bench.py -- headline benchmark of the PQ hot path (BASELINE.json: queries/sec at fixed
recall@10 on a 10M x 300-d PQ index; encode vectors/sec as a secondary figure).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference ...                    (the reference algorithm on host cores)

A "step" is one PQIndex.batchQuery of the whole query batch (100k queries, top-10) over the
index: ADC lookup-table build + uint8 code scan + top-k (+ all-gather merge when sharded).
`value` is timed with the queries and the index resident in HBM; `e2e` goes through the host
C-ABI call (gulon_pq_query / gulon_pq_query_sharded) with host query/result buffers.  Data are
synthetic (gulon_b200/csrc/synth_spec.h: a pure function of seed, row and column, identical on the
device and on host cores), codebooks are trained by the library itself.  Between timed steps
nothing is cached on purpose: the code planes (300 MB) and the lookup tables of a query batch
(> 70 MB per 2368-query tile) exceed the 126 MB L2 together, see config.l2.

Besides the headline the line carries, each with its own roofline / CPU baseline:
  encode      resident and end-to-end (host buffers) encode rates, oracle encode on host cores
  train_c3    configs[2]: k-means codebook training over all rows, 25 Lloyd iterations, rows sharded
              over the ranks with one all-reduce of sums / counts per iteration (world > 1)
  row_sharded configs[3]-shaped (world > 1): 12.5M x 128-d rows PER GPU, m = 16, code planes sharded by
              rows, NCCL all-gather + (distance, id) merge of every rank's k candidates, checked
              against the oracle on a query sample
  rerank_c5   configs[4]-shaped: 1M x 1000-d, m = 100, PQ top-R candidates + exact fp32 re-rank
  grouped_ivf the reference's partitioned index at its CLI defaults (rows / 1000 partitions, 5 % probed)
  other_shapes (1 GPU) configs[0] and one configs[3] shard: queries/s, encode rate, oracle check

The reference arm (`--impl reference`) rebuilds THE SAME index on host cores -- same synthetic rows,
the oracle's k-means and encode (bit-identical to the GPU's by the parity tests; `index_digest` in
both lines proves it for the run) -- and times the reference's prepareQuery + batchQuery on it.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec at fixed recall@10 on 10Mx300-d PQ index"
UNIT = "queries/s"
SEED = 20261018


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
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="headline only: skip train_c3 / row_sharded / rerank_c5 and the encode e2e/cpu legs")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    ap.add_argument("--row-shards", type=int, default=0,
                    help="row shards R of the code planes (0 = shard_plan); the other factor of the "
                         "world size splits the query batch")
    ap.add_argument("--c3-iters", type=int, default=25)
    ap.add_argument("--rs-rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--rs-queries", type=int, default=100_000)
    ap.add_argument("--c5-rows", type=int, default=1_000_000)
    ap.add_argument("--c5-queries", type=int, default=10_000)
    ap.add_argument("--c5-candidates", type=int, default=100)
    ap.add_argument("--gp-rows", type=int, default=1_000_000)
    ap.add_argument("--gp-queries", type=int, default=4_000)
    return ap.parse_args()


def shard_plan(n_rows, world, min_rows_per_shard=8_000_000):
    """(R row shards, C query groups), R * C == world -- the rule of gulon_b200.sharded.shard_plan,
    restated so that the reference arm prints the same `config` without importing the package."""
    R = 1
    for r in range(1, world + 1):
        if world % r == 0 and n_rows // r >= min_rows_per_shard:
            R = r
    return R, world // R


def config_of(a, world):
    """Identical in both arms (the driver compares them)."""
    R, Cq = shard_plan(a.rows, world) if a.row_shards <= 0 else (a.row_shards, world // a.row_shards)
    c2 = (a.rows, a.dim, a.m, a.k, a.queries) == (10_000_000, 300, 30, 10, 100_000)
    return {"workload": "%s: %dx%d-d word-embedding-shaped synthetic vectors, PQ m=%dx256, top-%d query "
                        "batch of %d" % ("configs[1]" if c2 else "custom shape (not a BASELINE config)",
                                         a.rows, a.dim, a.m, a.k, a.queries),
            "rows": a.rows, "dim": a.dim, "m": a.m, "clusters": 256, "k": a.k, "queries": a.queries,
            "data": "gulon_b200/csrc/synth_spec.h seed %d: 4096-centre mixture in a 32-d latent space mapped "
                    "to %d-d; queries = stream 1 of the same mixture" % (SEED, a.dim),
            "codebook": "trained on rows [0, %d), %d Lloyd iterations, running-mean update, ties to the "
                        "lowest index" % (min(a.train_rows, a.rows), a.train_iters),
            "sharding": ("%d row shard(s) of the code planes x %d query group(s); all-gather + (distance,id) "
                         "merge inside a group, all-gather of the slices across groups" % (R, Cq))
            if world > 1 else "single GPU",
            "l2": "inputs larger than L2: code planes (%d MB), their decoded bf16 copy (%d MB, the tensor scan's "
                  "operand) and the per-query lookup tables (%d MB) all exceed 126 MB; no flush needed"
                  % (a.rows * a.m // 1_000_000, a.rows * ((a.dim + 4 + 15) // 16 * 16) * 2 // 1_000_000,
                     a.queries * a.m * 1024 // 1_000_000),
            "arithmetic": "every returned distance is the reference's sequential fp32 table sum (dtype f32); the "
                          "tensor scan's lower bound, which only discards pairs, contracts bf16 operands with fp32 "
                          "accumulation"}


def index_digest(codebook, codes):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(codebook, np.float32).tobytes())
    h.update(np.ascontiguousarray(codes).tobytes())
    return h.hexdigest()[:16]


def probe_jvm():
    """is there a JVM toolchain on this box?"""
    out = {}
    for tool, arg in (("java", "-version"), ("javac", "-version"), ("scala", "-version"), ("sbt", "--version")):
        try:
            r = subprocess.run([tool, arg], capture_output=True, text=True, timeout=20)
            out[tool] = ((r.stderr or r.stdout).strip().splitlines() or ["?"])[0][:80]
        except FileNotFoundError:
            out[tool] = "not found"
        except Exception as e:  # pragma: no cover
            out[tool] = "error: %s" % type(e).__name__
    return out


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
def build_index_cpu(a, o, T):
    """The index of the GPU arm, rebuilt on host cores: same synthetic rows, the oracle's
    ProductQuantizer.apply (seed = quantizer index, running mean, lowest-index ties) and #encode.
    Cached under the temp directory: the driver runs this arm once per GPU count on the same box."""
    rows, D, M = a.rows, a.dim, a.m
    # (v2: bump when csrc/synth_spec.h or the oracle's training / encode change -- a cached index must be THE index)
    tag = "gulon_ref_index_v2_%dx%d_m%d_t%dx%d_s%d" % (rows, D, M, a.train_rows, a.train_iters, SEED)
    for base in ("/dev/shm", tempfile.gettempdir()):
        path = os.path.join(base, tag + ".npz")
        if os.path.exists(path):
            try:
                z = np.load(path)
                return z["cb"], z["codes"], "cache hit (%s)" % path, 0.0
            except Exception:
                pass
    t0 = time.perf_counter()
    mix = o.SynthMixture(D, seed=SEED)
    xt = mix.rows(0, min(a.train_rows, rows), nthreads=T)
    cb, _, _ = o.pq_train(xt, M, 256, a.train_iters, tie_mode=o.TIE_LOWEST, nthreads=T)
    del xt
    codes = np.empty((M, rows), np.uint8)
    CH = 250_000
    buf = np.empty((CH, D), np.float32)
    for r0 in range(0, rows, CH):
        n = min(CH, rows - r0)
        x = mix.rows(r0, r0 + n, out=buf[:n], nthreads=T)
        codes[:, r0:r0 + n] = o.pq_encode(x, cb, tie_mode=o.TIE_LOWEST, nthreads=T)
    dt = time.perf_counter() - t0
    note = "built in %.0f s" % dt
    for base in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(base)
            if st.f_bavail * st.f_frsize < codes.nbytes + cb.nbytes + (64 << 20):
                continue
            tmp = os.path.join(base, tag + ".%d.tmp.npz" % os.getpid())
            np.savez(tmp, cb=cb, codes=codes)
            os.replace(tmp, os.path.join(base, tag + ".npz"))
            note += ", cached in %s" % base
            break
        except Exception:
            continue
    return cb, codes, note, dt


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as o
    o.build()
    T = o.host_cores()
    cb, codes, how, build_s = build_index_cpu(a, o, T)
    nq = min(a.queries, a.cpu_queries or 8 * T)     # ~1 s of work per step on 16 cores at the c2 shape
    Q = o.SynthMixture(a.dim, seed=SEED).rows(0, nq, stream_seed=1, nthreads=T)   # the first nq queries

    def step():
        # Index.prepareQuery + PQIndex.batchQuery; one task per query as G/Tests.scala:109-121
        return o.pq_query(Q, cb, codes, a.k, topk_mode=o.TOPK_LITERAL, nthreads=T)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ids, ds, _ = step()
    dt = time.perf_counter() - t0
    v = nq * a.steps / dt
    sample = ("the first %d of the %d queries x the full %d-row index per step, literal TopKHeap"
              % (nq, a.queries, a.rows))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(a, a.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": T, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/); the "
                                 "reference is Scala/JVM and no JDK exists in this image",
                         "threads_from": "os.sched_getaffinity (not OMP_NUM_THREADS)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "index": {"digest": index_digest(cb, codes), "how": how,
                  "note": "same rows, codebook and codes as the GPU arm (compare `index.digest`)"},
        "jvm_probe": probe_jvm(),
        "gpu_launches": 0}))
    return 0


# ---- own arm ---------------------------------------------------------------------------------
class Ctx:
    """What every leg needs: ranks, device, timing helpers."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        import gulon_b200 as g
        from gulon_b200 import _native as N
        self.a, self.torch, self.dist, self.g, self.N = a, torch, dist, g, N
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if g.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: gulon_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        N.check(N.lib().gulon_set_device(self.local))
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_source = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        self.cores = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """W untimed + K timed calls bracketed by barrier + synchronize; ms for the K calls, max over ranks."""
        torch = self.torch
        out = None
        for _ in range(warmup):
            out = fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks(max(e0.elapsed_time(e1), 0.0)), self.max_over_ranks(wall), out

    def gather_codes(self, codes, n_local, bounds):
        """The row shards of the code planes, concatenated on rank 0 (host uint8 [M][N]); None elsewhere."""
        torch, dist = self.torch, self.dist
        if self.world == 1 or len(bounds) == 1:      # one shard: rank 0 already holds every row
            return codes[:, :n_local].cpu().numpy() if self.rank == 0 else None
        per = max(hi - lo for lo, hi in bounds)
        M = codes.shape[0]
        mine = torch.zeros((M, per), dtype=torch.uint8, device=self.dev)
        mine[:, :n_local] = codes[:, :n_local]
        parts = [torch.empty_like(mine) for _ in range(len(bounds))] if self.rank == 0 else None
        if len(bounds) == self.world:
            dist.gather(mine, parts, dst=0)
        else:      # R shards replicated over C groups: ranks 0..R-1 hold one full copy
            grp_ranks = list(range(len(bounds)))
            grp = dist.new_group(ranks=grp_ranks)
            if self.rank in grp_ranks:
                dist.gather(mine, parts, dst=0, group=grp)
        if self.rank != 0:
            return None
        return np.concatenate([p[:, :hi - lo].cpu().numpy() for p, (lo, hi) in zip(parts, bounds)], axis=1)


def encode_shard(cx, pq, mix, lo, hi, keep=None, time_it=True):
    """Encodes global rows [lo, hi) chunk by chunk into device code planes (rows independent: no
    exchange between ranks).  Returns (codes [M][stride], stride, ns, rows_timed)."""
    torch, N = cx.torch, cx.N
    M, D = len(pq.quantizers), pq.dimension
    n_local = hi - lo
    stride = (max(n_local, 1) + 15) // 16 * 16
    codes = torch.zeros((M, stride), dtype=torch.uint8, device=cx.dev)
    CH = 1 << 20
    ns, rows = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    buf = None if keep is not None else torch.empty((min(CH, max(n_local, 1)), D), dtype=torch.float32,
                                                    device=cx.dev)
    for r0 in range(0, n_local, CH):
        n = min(CH, n_local - r0)
        x = mix.rows(lo + r0, lo + r0 + n, out=keep[r0:r0 + n] if keep is not None else buf[:n])
        e0.record()
        N.check(N.lib().gulon_pq_encode_dev(pq.handle, x.data_ptr(), n, D, N.TIE_LOWEST,
                                            codes.data_ptr() + r0, stride, st))
        e1.record()
        torch.cuda.synchronize()
        if time_it and (r0 > 0 or n_local <= CH):          # first chunk is the warm-up
            ns += e0.elapsed_time(e1) * 1e6
            rows += n
    return codes, stride, ns, rows


def tensor_roofline(cx, D, M, Q, steps, n_local, prof_ms, shape_key, t_ns, t_l, boot_ns):
    """Roofline of tscan::filter2_kernel (the tensor scan's lower-bound contraction): the tensor pipe.
    FLOPs = 2 * 128 * 256 * KP per accumulator tile (counted by the kernel), against the SUSTAINED cuBLAS bf16
    figure of MEASURED_PEAKS.json (the kernel runs for hundreds of milliseconds under the power cap)."""
    N = cx.N
    KP = (D + 4 + 15) // 16 * 16
    st = {nm: N.counter("tscan_" + nm) for nm in ("tiles", "slow_paths", "survivors", "candidates", "pairs", "fallbacks",
                                                   "batches", "stages")}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flop = 2.0 * 128 * 256 * KP * st["tiles"]
    sec = t_ns * 1e-9
    ach = flop / sec / 1e12
    # SURVEY 8d's byte figure for the same work, for continuity with the pruned kernel's line: one pass over all
    # M code planes per tile of 16 queries
    alg_bytes = -(-Q // 16) * float(n_local) * M * steps
    roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "kernel": "tscan::filter2_kernel (tcgen05 lower bound over the decoded rows, cta_group::2)",
            "peak_source": ("measured, sustained cuBLAS bf16 (MEASURED_PEAKS.json; burst %.0f)" % peaks.get("bf16_tflops", 0))
            if peaks else "fallback (B200_PROFILING.md: ~1400 sustained)",
            "flop_per_launch": flop / t_l, "launch_seconds": sec / t_l, "launches": t_l,
            "contraction": "M128 x N256 x K%d per tile (D = %d coordinates + 4 bound columns), bf16 in, fp32 accumulate" % (KP, D),
            "kernel_share_of_step": t_ns * 1e-6 / prof_ms,
            "other_kernels_share_of_step": {"exact_kernel_boot_rows": boot_ns * 1e-6 / prof_ms},
            "timed_in": "a repeat of the K steps with the kernel timers on (%.1f ms per step; the timed region "
                        "runs without them)" % (prof_ms / steps),
            "hbm_algorithmic": {"unit": "GB/s", "bytes_per_step": alg_bytes / steps,
                                "achieved": alg_bytes / (prof_ms * 1e-3) / 1e9, "peak": cx.peak,
                                "frac": alg_bytes / (prof_ms * 1e-3) / 1e9 / cx.peak,
                                "note": "SURVEY 8d bytes (ceil(Q/16) passes over the M code planes) / whole step time: what the "
                                        "byte-gather formulation of this scan would have to move per second to keep up; the "
                                        "tensor scan itself reads the decoded rows from L2, not the code planes"},
            "filter": {"survivor_rate": st["survivors"] / max(st["pairs"], 1), "survivors": st["survivors"],
                       "list_candidates": st["candidates"], "stages_per_batch": st["stages"] / max(st["batches"], 1),
                       "batches": st["batches"], "handed_back_batches": st["fallbacks"],
                       "handed_back_queries": N.counter("tscan_handed_back_queries")}}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        key = "%s_tensor" % shape_key
        if cx.world == 1 and shape_key and key in tr:
            t = tr[key]
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = t["source"]
            roof["traffic_note"] = t.get("note")
    except Exception:
        pass
    return roof


def scan_roofline(cx, M, Q, steps, n_local, clk, prof_ms, shape_key=None, D=None):
    """Roofline of the dominant scan kernel from the library's per-launch event timers (profile = 1)."""
    N = cx.N
    scan_ns, scan_l = N.counter("scan_kernel_ns"), N.counter("scan_kernel_launches")
    p_ns, p_l = N.counter("pscan_kernel_ns"), N.counter("pscan_kernel_launches")
    f_ns, f_l = N.counter("pscan_first_kernel_ns"), N.counter("pscan_first_kernel_launches")
    pst = {nm: N.counter("pscan_" + nm) for nm in ("survivors", "candidates", "slow_items", "pairs", "main_pairs")}
    ml = N.counter("pscan_lb_quantizers") or M
    t_ns, t_l = N.counter("tscan_kernel_ns"), N.counter("tscan_kernel_launches")
    if t_l > 0 and t_ns >= max(p_ns, scan_ns):
        return tensor_roofline(cx, D or cx.a.dim, M, Q, steps, n_local, prof_ms, shape_key, t_ns, t_l, scan_ns)
    use_p = p_l > 0 and p_ns >= scan_ns
    k_ns, k_l = (p_ns, p_l) if use_p else (scan_ns, scan_l)
    if k_l <= 0:
        return None
    fb = 8 if 127 // ml >= 3 else 16
    QT = (N.counter("pscan_qt") or 128 // fb) if use_p else 4
    rows_main = (pst["main_pairs"] / (steps * Q)) if use_p else n_local
    # algorithmic bytes: one pass over ALL M code planes of the scanned rows per tile of QT
    # queries whose tables share shared memory
    alg = steps * -(-Q // QT) * rows_main * M / k_l
    sec = k_ns * 1e-9 / k_l
    ach = alg / sec / 1e9
    lookups = steps * Q * rows_main * (ml if use_p else M) / (k_ns * 1e-9)
    bpe = (fb // 8) if use_p else 4
    clk_mhz = (clk or {}).get("sm_mhz") or 1965.0
    smem_peak = 148 * 128 * clk_mhz * 1e6
    roof = {"bound": "hbm", "achieved": ach, "peak": cx.peak, "unit": "GB/s", "frac": ach / cx.peak,
            "traffic": None, "kernel": "pscan::pruned_scan_kernel (main stage)" if use_p else "fscan::fused_scan_kernel",
            "peak_source": cx.peak_source,
            "algorithmic_bytes_per_launch": alg, "launch_seconds": sec, "launches": k_l, "query_tile": QT,
            "lower_bound_quantizers": ml if use_p else None,
            "note": ("algorithmic bytes count all M code planes per pass (SURVEY 8d); the pruned kernel streams only "
                     "the planes its lower bound sums, so `achieved` is work done per second, not bytes moved; the "
                     "binding resource is the shared-memory crossbar, see smem_gather") if use_p else None,
            "kernel_share_of_step": k_ns * 1e-6 / prof_ms,
            "other_kernels_share_of_step": {"first_stage_pruned": f_ns * 1e-6 / prof_ms,
                                            "exact_kernel_boot_rows" if use_p else "pruned": (scan_ns if use_p else p_ns) * 1e-6 / prof_ms},
            "timed_in": "a repeat of the K steps with the kernel timers on (%.1f ms per step; the timed region "
                        "runs without them)" % (prof_ms / steps),
            "smem_gather": {"bytes_per_entry": bpe, "achieved_GBps": lookups * bpe / 1e9,
                            "peak_GBps": smem_peak / 1e9, "frac": lookups * bpe / smem_peak,
                            "note": "table reads from shared memory: the binding resource (128 B/clk/SM crossbar)"}}
    # dram bytes of the main-stage launch from the committed `ncu --set full` capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        key = "%s_q%d" % (shape_key, QT * 148)
        if use_p and cx.world == 1 and shape_key and key in tr:
            t = tr[key]
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = t["source"]
            roof["traffic_algorithmic_bytes_same_launch"] = t["algorithmic_bytes_per_launch"]
            roof["dram_frac"] = t["dram_bytes_per_launch"] / t["launch_seconds"] / 1e9 / cx.peak
    except Exception:
        pass
    if use_p and pst["pairs"]:
        roof["pruning"] = {"survivor_rate": pst["survivors"] / pst["pairs"], "list_candidates": pst["candidates"],
                           "slow_path_items": pst["slow_items"]}
    return roof


def leg_train_c3(cx, mix, X_full):
    """configs[2]: ProductQuantizer.apply over ALL rows, 25 Lloyd iterations, sum-mode update; rows
    sharded over the ranks, one all-reduce of the fixed-point sums / counts / changed-count per
    iteration through the gulon_comm_t hooks (NCCL).  At world > 1 rank 0 repeats the training alone
    and compares the centroids bit for bit."""
    torch, g, N, a = cx.torch, cx.g, cx.N, cx.a
    from gulon_b200.sharded import TorchComm, shard_bounds
    rows, D, M = a.rows, a.dim, a.m
    lo, hi = shard_bounds(rows, cx.world)[cx.rank]
    X = X_full if (X_full is not None and cx.world == 1) else mix.rows(lo, hi)
    pts = g.DevicePoints.from_torch(X)
    cfg = g.ProductQuantizerConfig(256, M, a.c3_iters, update_mode=g.UPDATE_SUM)
    comm = TorchComm(device=cx.dev) if cx.world > 1 else None

    def run():
        if comm is None:
            return g.ProductQuantizer.train(pts, cfg)
        return g.ProductQuantizer.train(pts, cfg, comm=comm.struct, n_total=rows, row_offset=lo)

    run()                                   # warm-up: allocations, tensor maps, NCCL channels
    secs = []
    for _ in range(3):                      # three timed trainings (each: barrier, train, barrier; max over ranks)
        calls0 = dict(comm.calls) if comm else {}
        bytes0 = dict(comm.bytes) if comm else {}
        ms, wall, pq = cx.timed(run, 1, 0)
        secs.append(max(ms, wall) * 1e-3)
    sec = float(np.median(secs))
    updates = N.counter("train_updates")
    alg = (updates + 1) * rows * D * 4.0      # the matrix is read at least once per Lloyd iteration
    out = {"config": "configs[2]: %dx%d-d, m=%dx256, max %d Lloyd iterations, sum-mode update" % (rows, D, M, a.c3_iters),
           "seconds": sec, "seconds_each": secs, "n_gpus": cx.world, "centroid_updates": updates,
           "assignment_passes": updates + 1,
           "kernel": "tca::tc_assign_kernel<.., FUSE> (assignment + fixed-point sums + changed count, one pass)",
           "row_iterations_per_s": rows * (updates + 1) / sec,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": cx.peak * cx.world,
                        "achieved": alg / sec / 1e9, "frac": alg / sec / 1e9 / (cx.peak * cx.world),
                        "algorithmic_bytes_per_pass": rows * D * 4,
                        "note": "N*D*4 bytes per Lloyd iteration (one read of the matrix); whole training time"}}
    if comm:
        d = {k: comm.calls[k] - calls0.get(k, 0) for k in comm.calls}
        b = {k: comm.bytes[k] - bytes0.get(k, 0) for k in comm.bytes}
        out["allreduce_calls"] = sum(v for k, v in d.items() if k.startswith("allreduce"))
        out["allreduce_calls_by_kind"] = d
        out["bytes_per_allreduce"] = {k: (b[k] // d[k] if d[k] else 0) for k in d if k.startswith("allreduce")}
        cb = pq.codebook()
        # every rank must hold the same centroids
        t = torch.from_numpy(cb.view(np.int32).astype(np.int64)).to(cx.dev)
        mn, mx = t.clone(), t.clone()
        cx.dist.all_reduce(mn, op=cx.dist.ReduceOp.MIN)
        cx.dist.all_reduce(mx, op=cx.dist.ReduceOp.MAX)
        out["centroids_equal_on_all_ranks"] = bool(torch.equal(mn, mx))
        if cx.rank == 0:
            del pts
            Xa = X_full if X_full is not None else mix.rows(0, rows)
            single = g.ProductQuantizer.train(g.DevicePoints.from_torch(Xa), cfg)
            out["centroids_equal_single_rank"] = bool(np.array_equal(cb.view(np.uint32),
                                                                     single.codebook().view(np.uint32)))
            del Xa
        cx.barrier()
    return out


def leg_train_cpu(cx, o, T):
    """The reference's ProductQuantizer.apply on host cores over a bounded sample (oracle port)."""
    a = cx.a
    n, it = 65_536, 2
    x = o.SynthMixture(a.dim, seed=SEED).rows(0, n, nthreads=T)
    t0 = time.perf_counter()
    _, nu, _ = o.pq_train(x, a.m, 256, it, tie_mode=o.TIE_LOWEST, nthreads=T)
    dt = time.perf_counter() - t0
    passes = float(np.mean(nu)) + 1
    return {"value": n * passes / dt, "unit": "row-iterations/s", "cores": T, "kind": "port",
            "sample": "%d rows x %d-d, m=%d, max %d iterations (%.1f assignment passes on average), running mean"
                      % (n, a.dim, a.m, it, passes)}


def leg_row_sharded(cx, clk_unused):
    """configs[3]-shaped: rs_rows_per_gpu x 128-d SIFT-shaped rows PER GPU, m = 16 x 256, the code planes
    sharded by rows, every rank's k candidates merged by one NCCL all-gather + (distance, id) merge
    (gulon_pq_query_sharded_dev).  Weak scaling in the index size; the query batch is fixed."""
    torch, g, N, a, dist = cx.torch, cx.g, cx.N, cx.a, cx.dist
    from gulon_b200.sharded import ShardedPQIndex, shard_bounds
    from gulon_b200.synth import Mixture
    D, M, k, Q = 128, 16, a.k, a.rs_queries
    rows = a.rs_rows_per_gpu * cx.world
    mix = Mixture(D, centres=16384, nonneg=True, span=40.0, seed=SEED + 4, device=cx.dev)
    xt = mix.rows(0, min(a.train_rows, rows))
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, a.train_iters))
    del xt
    bounds = shard_bounds(rows, cx.world)
    lo, hi = bounds[cx.rank]
    codes, stride, enc_ns, enc_rows = encode_shard(cx, pq, mix, lo, hi)
    ix = g.PQIndex.from_device_codes(pq, codes, hi - lo)
    sh = ShardedPQIndex(ix, lo, plan=(cx.world, 1))
    queries = mix.rows(0, Q, stream_seed=1)
    W, K = min(a.warmup, 2), min(a.steps, 3)
    ms, _, out = cx.timed(lambda: sh.batch_query(k, queries), K, W)
    value = Q * K / (ms * 1e-3)
    # the same steps without the exchange: local scan of this rank's shard only
    ms_local, _, _ = cx.timed(lambda: ix.batch_query_dev(k, queries, id_offset=lo), K, 1)
    # the exchange alone: all-gather of [ids | dists] + merge of `world` candidate lists
    send = torch.zeros((2 * Q * k,), dtype=torch.int32, device=cx.dev)
    recv = torch.zeros((cx.world * 2 * Q * k,), dtype=torch.int32, device=cx.dev)
    ms_ag, _, _ = cx.timed(lambda: dist.all_gather_into_tensor(recv, send), K, 1)
    ids_all = torch.zeros((cx.world, Q, k), dtype=torch.int32, device=cx.dev)
    ds_all = torch.rand((cx.world, Q, k), dtype=torch.float32, device=cx.dev).sort(dim=2).values
    oi = torch.empty((Q, k), dtype=torch.int32, device=cx.dev)
    od = torch.empty((Q, k), dtype=torch.float32, device=cx.dev)
    st = torch.cuda.current_stream().cuda_stream
    ms_mg, _, _ = cx.timed(lambda: N.check(N.lib().gulon_topk_merge_dev(
        ids_all.data_ptr(), ds_all.data_ptr(), cx.world, Q, k, oi.data_ptr(), od.data_ptr(), None, st)), K, 1)
    del send, recv, ids_all, ds_all
    # roofline of the main-stage kernel on this rank (every rank runs the same shapes)
    g.set_option("profile", 1)
    pms, _, _ = cx.timed(lambda: sh.batch_query(k, queries), K, 0)
    roof = scan_roofline(cx, M, Q, K, hi - lo, None, pms, D=D)
    g.set_option("profile", 0)
    # end to end: host queries in, host answers out on every rank
    q_host = torch.empty((Q, D), dtype=torch.float32, pin_memory=True)
    q_host.copy_(queries)
    torch.cuda.synchronize()
    q_np = q_host.numpy()
    ms_e, wall_e, res = cx.timed(lambda: sh.batch_query(k, q_np), K, 1)
    e2e = Q * K / (max(ms_e, wall_e) * 1e-3)
    same = bool(np.array_equal(out[0].cpu().numpy(), res[0]))
    line = {"config": "configs[3]-shaped: %d x %d-d SIFT-shaped rows (%d per GPU), m=%dx256, top-%d batch of %d queries, "
                      "%d row shards, NCCL all-gather + (distance,id) merge" % (rows, D, a.rs_rows_per_gpu, M, k, Q, cx.world),
            "value": value, "unit": UNIT, "ms_per_step": ms / K, "steps": K, "warmup": W, "scaling": "weak (rows per GPU fixed)",
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": cx.world * Q * D * 4,
                    "d2h_bytes_per_step": cx.world * Q * (k * 8 + 4), "ids_equal_device_path": same},
            "local_scan_ms": ms_local / K, "exchange_ms": (ms_ag + ms_mg) / K,
            "allgather_ms": ms_ag / K, "allgather_bytes_per_rank": 2 * Q * k * 4, "merge_ms": ms_mg / K,
            "exchange_note": "all-gather and merge timed in isolation on the same shapes; the step itself is "
                             "%.1f ms against %.1f ms for the local scan alone (separate runs)" % (ms / K, ms_local / K),
            "roofline": roof,
            "encode": {"value": enc_rows / (enc_ns * 1e-9) if enc_ns else None, "unit": "vectors/s per GPU",
                       "aggregate": cx.world * enc_rows / (enc_ns * 1e-9) if enc_ns else None}}
    # oracle check of a query sample over the CONCATENATED shards (rank 0, all host cores)
    nqc = 8
    full = cx.gather_codes(codes, hi - lo, bounds)
    if cx.rank == 0:
        from oracle import oracle as o
        T = o.host_cores()
        t0 = time.perf_counter()
        ci, cd, _ = o.pq_query(q_np[:nqc], pq.codebook(), full, k, topk_mode=o.TOPK_CANONICAL, nthreads=T)
        dt = time.perf_counter() - t0
        gi, gd = out[0][:nqc].cpu().numpy(), out[1][:nqc].cpu().numpy()
        line["matches_oracle"] = bool(np.array_equal(ci, gi) and np.array_equal(cd.view(np.uint32), gd.view(np.uint32)))
        line["cpu_baseline"] = {"value": nqc / dt, "unit": UNIT, "cores": T, "kind": "port",
                                "sample": "%d queries x the concatenated %d-row index" % (nqc, rows)}
        del full
    cx.barrier()
    return line


def leg_rerank_c5(cx):
    """configs[4]-shaped: 1M x 1000-d entity embeddings, m = 100 x 256: batched encode, PQ top-R
    candidates, exact fp32 re-rank of the candidates against the raw vectors (rows sharded over the
    ranks; candidates are re-ranked by the rank that owns the row, then merged)."""
    torch, g, N, a, dist = cx.torch, cx.g, cx.N, cx.a, cx.dist
    from gulon_b200.sharded import ShardedPQIndex, shard_bounds
    from gulon_b200.synth import Mixture
    from gulon_b200.pipeline import RerankPipeline
    D, M, k, Q, R = 1000, 100, a.k, a.c5_queries, a.c5_candidates
    rows = a.c5_rows
    mix = Mixture(D, seed=SEED + 5, device=cx.dev)
    xt = mix.rows(0, min(65_536, rows))
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, 4))
    del xt
    bounds = shard_bounds(rows, cx.world)
    lo, hi = bounds[cx.rank]
    X = torch.empty((hi - lo, D), dtype=torch.float32, device=cx.dev)
    codes, stride, enc_ns, enc_rows = encode_shard(cx, pq, mix, lo, hi, keep=X)
    pipe = RerankPipeline(pq, codes, X, lo, hi - lo, world=cx.world)
    queries = mix.rows(0, Q, stream_seed=1)
    W, K = min(a.warmup, 2), min(a.steps, 3)
    ms, _, out = cx.timed(lambda: pipe.query(k, R, queries), K, W)
    value = Q * K / (ms * 1e-3)
    ms_r, _, _ = cx.timed(lambda: pipe.rerank_only(k, R, queries), K, 1)
    line = {"config": "configs[4]-shaped: %dx%d-d, m=%dx256, %d queries: PQ top-%d -> exact fp32 re-rank -> top-%d"
                      % (rows, D, M, Q, R, k),
            "value": value, "unit": UNIT, "ms_per_step": ms / K, "n_gpus": cx.world,
            "encode": {"value": cx.world * enc_rows / (enc_ns * 1e-9) if enc_ns else None, "unit": "vectors/s",
                       "roofline": {"bound": "hbm", "unit": "GB/s", "peak": cx.peak,
                                    "achieved": enc_rows * (D * 4 + M) / enc_ns if enc_ns else None,
                                    "frac": enc_rows * (D * 4 + M) / enc_ns / cx.peak if enc_ns else None}},
            "rerank": {"ms_per_step": ms_r / K,
                       "roofline": {"bound": "hbm", "unit": "GB/s", "peak": cx.peak * cx.world,
                                    "algorithmic_bytes_per_query": R * D * 4,
                                    "achieved": cx.world * Q * R * D * 4 / (ms_r / K * 1e-3) / 1e9,
                                    "frac": Q * R * D * 4 / (ms_r / K * 1e-3) / 1e9 / cx.peak,
                                    "note": "the exact-distance kernel alone on a fixed candidate list: every rank "
                                            "scores R candidates of its own rows for all Q queries"}}}
    nqc = 16
    if cx.world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as o
        T = o.host_cores()
        hc = codes[:, :hi - lo].cpu().numpy()
        qn = queries[:nqc].cpu().numpy()
        cbh = pq.codebook()
        t0 = time.perf_counter()
        ci, _, _ = o.pq_query(qn, cbh, hc, R, topk_mode=o.TOPK_CANONICAL, nthreads=T)
        dt = time.perf_counter() - t0
        want_i = np.full((nqc, k), -1, np.int32)
        want_d = np.full((nqc, k), np.inf, np.float32)
        for q in range(nqc):
            cand = np.sort(ci[q][ci[q] >= 0])
            xc = X[torch.from_numpy(cand.astype(np.int64)).to(cx.dev)].cpu().numpy()   # only the candidate rows
            t1 = time.perf_counter()
            ei, ed, _ = o.exact_nn(xc, qn[q:q + 1], k, topk_mode=o.TOPK_CANONICAL)
            dt += time.perf_counter() - t1
            want_i[q, :len(ei[0])] = cand[ei[0]]
            want_d[q, :len(ed[0])] = ed[0]
        gi, gd = out[0][:nqc].cpu().numpy(), out[1][:nqc].cpu().numpy()
        line["matches_oracle"] = bool(np.array_equal(want_i, gi) and np.array_equal(want_d.view(np.uint32), gd.view(np.uint32)))
        line["cpu_baseline"] = {"value": nqc / dt, "unit": UNIT, "cores": T, "kind": "port",
                                "sample": "%d queries: oracle PQ top-%d over %d rows + exact distances of the candidates"
                                          % (nqc, R, rows)}
    return line


def leg_shapes(cx):
    """The other BASELINE shapes on one GPU, each with its own oracle check: configs[0] (1M x 100-d iid rows,
    m = 10: no cluster structure, the lower bound prunes least) and one configs[3] shard (12.5M x 128-d
    SIFT-shaped, m = 16).  Resident queries, 2 timed steps after one warm-up step."""
    torch, g, N, a = cx.torch, cx.g, cx.N, cx.a
    from gulon_b200.synth import Mixture
    out = {}
    for name, rows, D, M, kw, Q in (("c1_1Mx100_m10_iid", 1_000_000, 100, 10, dict(centres=0), 10_000),
                                    ("c4_shard_12.5Mx128_m16", a.rs_rows_per_gpu, 128, 16,
                                     dict(centres=16384, nonneg=True, span=40.0), 20_000)):
        mix = Mixture(D, seed=SEED + (1 if D == 100 else 4), device=cx.dev, **kw)
        xt = mix.rows(0, min(65_536, rows))
        pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt), g.ProductQuantizerConfig(256, M, 4))
        del xt
        codes, stride, enc_ns, enc_rows = encode_shard(cx, pq, mix, 0, rows)
        ix = g.PQIndex.from_device_codes(pq, codes, rows)
        queries = mix.rows(0, Q, stream_seed=1)
        ms, _, res = cx.timed(lambda: ix.batch_query_dev(a.k, queries), 2, 1)
        line = {"rows": rows, "dim": D, "m": M, "queries": Q, "value": Q * 2 / (ms * 1e-3), "unit": UNIT,
                "scan": {4: "tensor", 3: "pruned", 2: "fused", 1: "simple"}.get(N.counter("scan_last_impl"), "?"),
                "lower_bound_quantizers": N.counter("pscan_lb_quantizers") if N.counter("scan_last_impl") == 3 else None,
                "encode_vectors_per_s": enc_rows / (enc_ns * 1e-9) if enc_ns else None}
        if not a.no_cpu_baseline:
            from oracle import oracle as o
            T = o.host_cores()
            nqc = 8
            t0 = time.perf_counter()
            ci, cd, _ = o.pq_query(queries[:nqc].cpu().numpy(), pq.codebook(), codes[:, :rows].cpu().numpy(), a.k,
                                   topk_mode=o.TOPK_CANONICAL, nthreads=T)
            dt = time.perf_counter() - t0
            line["matches_oracle"] = bool(np.array_equal(ci, res[0][:nqc].cpu().numpy()) and
                                          np.array_equal(cd.view(np.uint32), res[1][:nqc].cpu().numpy().view(np.uint32)))
            line["cpu_baseline"] = {"value": nqc / dt, "unit": UNIT, "cores": T, "kind": "port",
                                    "sample": "%d queries x the full index" % nqc}
        out[name] = line
        del ix, codes, queries
        torch.cuda.empty_cache()
    return out


def leg_grouped(cx):
    """The reference's partitioned index at its CLI defaults (C/BuildIndex.scala:98-108: rows / 1000
    partitions, probe 5 % of them): WordVectors#grouped + Index.grouped + GroupedIndex#batchQuery on
    gp_rows x 300-d rows.  One launch covers every probed (query, partition) pair
    (gulon_grouped_query_dev); the per-pair lookup-table rebuild dominates, as it does in the
    reference (M * 256 * dsub * 3 flops against M gathers per row of a ~1000-row partition)."""
    torch, g, N, a = cx.torch, cx.g, cx.N, cx.a
    from gulon_b200.synth import Mixture
    D, M, k = a.dim, a.m, a.k
    rows, Q = a.gp_rows, a.gp_queries
    P = max(rows // 1000, 1)
    limit = max(int(P * 0.05), 5)
    mix = Mixture(D, seed=SEED, device=cx.dev)
    X = mix.rows(0, rows)
    xh = X.cpu().numpy()
    t0 = time.perf_counter()
    coarse = g.KMeans.compute_clusters(g.Vectors(g.Matrix(xh[:min(rows, 200_000)])), g.KMeansConfig(P, 3, seed=0))
    gv = g.GroupedVectors.group(g.Matrix(xh), coarse, device=cx.dev)
    res = gv.residuals_dev()
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(res[:min(rows, a.train_rows)].contiguous()),
                                  g.ProductQuantizerConfig(256, M, 4))
    ix = g.GroupedIndex.build(gv, pq, strategy=g.LimitGroups(limit))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    qh = mix.rows(0, Q, stream_seed=1).cpu().numpy()
    ix.batch_query(k, qh[:256])                      # warm-up
    t0 = time.perf_counter()
    r = ix.batch_query(k, qh)
    dt = time.perf_counter() - t0
    line = {"config": "%d x %d-d rows, %d partitions (rows / 1000), LimitGroups(%d) = 5 %% of the partitions, residual "
                      "PQ m=%dx256, %d queries, top-%d" % (rows, D, P, limit, M, Q, k),
            "value": Q / dt, "unit": UNIT, "e2e": True, "build_seconds": build_s,
            "pairs_per_query": limit, "note": "host queries in, host answers out (search space + one device launch)"}
    if not a.no_cpu_baseline:
        from oracle import oracle as o
        nqc = 8
        codes = ix.vector_index._keepalive[:, :rows].cpu().numpy()
        t0 = time.perf_counter()
        wi, wd, ws = o.grouped_query(qh[:nqc], gv.centroids, gv.offsets, gv.size, pq.codebook(), codes, k,
                                     ("groups", limit))
        dtc = time.perf_counter() - t0
        line["matches_oracle"] = bool(np.array_equal(wi, r.keys[:nqc]) and
                                      np.array_equal(wd.view(np.uint32), r.values[:nqc].view(np.uint32)))
        line["cpu_baseline"] = {"value": nqc / dtc, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d queries, GroupedIndex#query restated (one thread, as the reference's "
                                          "query path is sequential per query)" % nqc}
    return line


def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    cx = Ctx(a)
    torch, dist, g, N = cx.torch, cx.dist, cx.g, cx.N
    from gulon_b200.sharded import ShardedPQIndex, shard_bounds
    from gulon_b200.synth import Mixture
    world, rank, dev = cx.world, cx.rank, cx.dev

    for kv in a.opt:
        name, val = kv.split("=")
        g.set_option(name, int(val))
    D, M, K, k, Q = a.dim, a.m, 256, a.k, a.queries
    mix = Mixture(D, seed=SEED, device=dev)

    # codebooks: trained by the library on the first rows of the data set (same on every rank)
    xt = mix.rows(0, min(a.train_rows, a.rows))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pq = g.ProductQuantizer.train(g.DevicePoints.from_torch(xt),
                                  g.ProductQuantizerConfig(K, M, a.train_iters))
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    train_passes = N.counter("train_updates") + 1
    del xt

    # (R row shards) x (C query groups) over the ranks
    n_shards, n_groups = shard_plan(a.rows, world) if a.row_shards <= 0 else (a.row_shards, world // a.row_shards)
    bounds = shard_bounds(a.rows, n_shards)
    lo, hi = bounds[rank % n_shards]
    n_local = hi - lo
    keep_x = world == 1 and (not a.no_recall or not a.no_extra_legs)
    X = torch.empty((n_local, D), dtype=torch.float32, device=dev) if keep_x else None
    codes, stride, enc_ns, enc_rows = encode_shard(cx, pq, mix, lo, hi, keep=X)
    st = torch.cuda.current_stream().cuda_stream
    # candidate statistics of the tensor-core encode (one untimed chunk, profile counters on)
    enc_stats = None
    if n_local > 0:
        n = min(1 << 20, n_local)
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
    sh = ShardedPQIndex(ix, lo, plan=(n_shards, n_groups)) if world > 1 else None
    queries = mix.rows(0, Q, stream_seed=1)

    def step_dev():
        return sh.batch_query(k, queries) if world > 1 else ix.batch_query_dev(k, queries)

    # warm-up, then the timed region
    for _ in range(a.warmup):
        step_dev()
    cx.barrier()
    launches0 = g.kernel_launches()
    clocks = ClockSampler(cx.local)
    if rank == 0:
        clocks.start()
    ms, _, out = cx.timed(step_dev, a.steps, 0)
    clk = clocks.stop() if rank == 0 else None
    launches = g.kernel_launches() - launches0
    value = Q * a.steps / (ms * 1e-3)
    # the same K steps once more with the library's kernel timers and counters on (they add an event pair
    # per kernel and a host sync per scan stage, so they stay out of the timed region above)
    g.set_option("profile", 1)
    prof_ms, _, _ = cx.timed(step_dev, a.steps, 0)
    roof = scan_roofline(cx, M, Q, a.steps, n_local, clk, prof_ms, shape_key="%dx%d_m%d" % (a.rows, D, M))
    g.set_option("profile", 0)

    # e2e: the host-facing call, host query buffer in, host (ids, distances) out, every step
    q_host = torch.empty((Q, D), dtype=torch.float32, pin_memory=True)
    q_host.copy_(queries)
    torch.cuda.synchronize()
    q_np = q_host.numpy()

    def step_e2e():
        if world == 1:
            r = ix.batch_query(k, q_np)          # gulon_pq_query: H2D + scan + D2H inside
            return r.keys, r.values
        ids, ds, _ = sh.batch_query(k, q_np)     # gulon_pq_query_sharded: this rank's slice H2D, whole answer D2H
        return ids, ds

    e_ms, e_wall, res = cx.timed(step_e2e, a.steps, 1)
    e2e_ms = max(e_ms, e_wall)
    e2e_value = Q * a.steps / (e2e_ms * 1e-3)

    ids_dev = out[0].cpu().numpy()
    dist_dev = out[1].cpu().numpy()
    same = bool(np.array_equal(ids_dev, np.asarray(res[0])))

    extra = {}
    # recall@10 per G/Tests.scala:18-41 on a query sample (single GPU: raw vectors are resident)
    if keep_x and rank == 0 and not a.no_recall:
        from gulon_b200.index import rerank
        Rq = min(a.recall_queries, Q)
        qs = q_np[:Rq]
        pts = g.DevicePoints.from_torch(X)
        gt = g.exact_nearest_neighbours(pts, qs, k)
        ex = rerank(pts, qs, ids_dev[:Rq], k)            # exact distances of the returned keys
        kth = gt.values[:, k - 1]
        tp = [(ex.values[i][ex.keys[i] >= 0] <= kth[i]).sum() for i in range(Rq)]
        extra["recall_at_10"] = {"mean": float(np.mean(tp)) / k, "queries": Rq,
                                 "definition": "G/Tests.scala:18-41, eps=0"}
        del pts

    # ---- CPU baseline + oracle check of the headline (rank 0) --------------------------------------
    cpu = None
    digest = None
    enc_cpu = None
    train_cpu = None
    o = None
    T = 0
    if rank == 0 and not a.no_cpu_baseline:
        from oracle import oracle as o
        o.build()
        T = o.host_cores()
    full_codes = cx.gather_codes(codes, n_local, bounds) if not a.no_cpu_baseline else None
    if o is not None:
        try:
            # ~10 s of host work at the c2 shape on one GPU (about 7 queries/s per core); 16 queries at world > 1
            per_s = 3e8 / max(1.0, float(a.rows) * M)
            nq = a.cpu_queries or (max(64, int(64 * T * per_s)) if world == 1 else 16)
            nq = min(Q, nq)
            cb = pq.codebook()
            digest = index_digest(cb, full_codes)
            t0 = time.perf_counter()
            ci, cd, _ = o.pq_query(q_np[:nq], cb, full_codes, k, topk_mode=o.TOPK_CANONICAL, nthreads=T)
            dt = time.perf_counter() - t0
            ok = bool(np.array_equal(ci, ids_dev[:nq]) and np.array_equal(cd.view(np.uint32), dist_dev[:nq].view(np.uint32)))
            cpu = {"value": nq / dt, "unit": UNIT, "cores": T, "kind": "port",
                   "sample": "the first %d of the %d queries x the full %d-row index" % (nq, Q, a.rows),
                   "matches_gpu": ok, "threads_from": "os.sched_getaffinity"}
            if world > 1:      # the CPU baseline proper is an N = 1 figure; at N > 1 this is the parity check
                extra["oracle_check"] = dict(cpu, note="sharded answer (NCCL path) against the oracle over the whole index")
                cpu = None
            if not a.no_extra_legs:
                # encode on host cores: ProductQuantizer#encode over a row sample, M subspace tasks in parallel
                ne = 100_000
                xs = o.SynthMixture(D, seed=SEED).rows(0, min(ne, a.rows), nthreads=T)
                t0 = time.perf_counter()
                hc = o.pq_encode(xs, cb, tie_mode=o.TIE_LOWEST, nthreads=T)
                dt = time.perf_counter() - t0
                enc_cpu = {"value": len(xs) / dt, "unit": "vectors/s", "cores": T, "kind": "port",
                           "sample": "rows [0, %d) of the data set" % len(xs),
                           "matches_gpu": bool(lo == 0 and np.array_equal(hc, full_codes[:, :len(xs)]))}
                train_cpu = leg_train_cpu(cx, o, T)
        except Exception as e:      # pragma: no cover -- a failing checker must not lose the measured line
            extra["cpu_baseline_error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
        del full_codes
    cx.barrier()

    # ---- encode end to end: host rows in, host codes out (gulon_pq_encode) ---------------------------
    enc_e2e = None
    if not a.no_extra_legs and rank == 0:
        ne = min(1 << 20, a.rows)
        xh = torch.empty((ne, D), dtype=torch.float32, pin_memory=True)
        xh.copy_(X[:ne] if keep_x else mix.rows(0, ne))
        torch.cuda.synchronize()
        xn = xh.numpy()
        hc = np.empty((M, ne), np.uint8)
        fn = N.lib().gulon_pq_encode
        N.check(fn(pq.handle, xn.ctypes.data, ne, D, N.TIE_LOWEST, hc.ctypes.data))
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            N.check(fn(pq.handle, xn.ctypes.data, ne, D, N.TIE_LOWEST, hc.ctypes.data))
        dt = (time.perf_counter() - t0) / reps
        enc_e2e = {"value": ne / dt, "unit": "vectors/s", "rows": ne, "h2d_bytes_per_step": ne * D * 4,
                   "d2h_bytes_per_step": ne * M, "pcie_GBps": ne * (D * 4 + M) / dt / 1e9,
                   "codes_equal_device_path": bool(lo == 0 and np.array_equal(
                       hc, codes[:, :ne].cpu().numpy()))}
        del xh
    cx.barrier()

    # ---- other legs ---------------------------------------------------------------------------------
    # The headline above is complete; a leg that fails must not lose it.  A failure is recorded in the leg's
    # slot and the remaining legs are skipped on every rank (the ranks agree on that through an all-reduce;
    # single-rank legs run last).
    legs = {}

    def run_leg(name, fn, collective=True):
        err = None
        try:
            legs[name] = fn()
        except Exception as e:      # pragma: no cover
            err = "%s: %s" % (type(e).__name__, str(e)[:300])
            legs[name] = {"error": err}
        bad = 1.0 if err else 0.0
        if collective and world > 1:
            t = torch.tensor([bad], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            bad = float(t.item())
            if bad and not err:
                legs[name] = {"error": "failed on another rank"}
        return bad == 0.0

    if not a.no_extra_legs:
        del ix, sh
        ok = run_leg("train_c3", lambda: leg_train_c3(cx, mix, X))
        if ok and train_cpu:
            legs["train_c3"]["cpu_baseline"] = train_cpu
        X = None
        torch.cuda.empty_cache()
        if ok and world > 1:
            ok = run_leg("row_sharded", lambda: leg_row_sharded(cx, clk))
        if ok:
            ok = run_leg("rerank_c5", lambda: leg_rerank_c5(cx))
        if rank == 0:
            run_leg("grouped_ivf", lambda: leg_grouped(cx), collective=False)
            if world == 1:
                run_leg("other_shapes", lambda: leg_shapes(cx), collective=False)
        cx.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(a, world),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            # whole job: every query group copies its slice of the batch to each of its R row shards;
            # every rank reads the assembled (ids, distances, sizes) back
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_shards * Q * D * 4,
                    "d2h_bytes_per_step": world * Q * (k * 8 + 4),
                    "ms_per_step": e2e_ms / a.steps, "ids_equal_device_path": same},
            "gpu_launches": launches,
            "index": {"digest": digest, "note": "sha256 of codebook + code planes; equals the reference arm's"},
            "encode": dict({"value": world * enc_rows / (enc_ns * 1e-9) if enc_ns else None, "unit": "vectors/s",
                            "rows": enc_rows, "resident": True, "n_gpus": world,
                            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": cx.peak,
                                         "achieved": enc_rows * (D * 4 + M) / enc_ns if enc_ns else None,
                                         "frac": enc_rows * (D * 4 + M) / enc_ns / cx.peak if enc_ns else None,
                                         "algorithmic_bytes_per_vector": D * 4 + M},
                            "e2e": enc_e2e, "cpu_baseline": enc_cpu},
                           **(enc_stats or {})),
            "train": {"rows": min(a.train_rows, a.rows), "iters": a.train_iters, "seconds": train_s,
                      "assignment_passes": train_passes, "note": "the codebook of the index (running-mean update)"},
            "jvm_probe": probe_jvm(),
        }
        line.update(extra)
        line.update(legs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
