"""bench.py host logic that the driver's comparison of the two arms relies on (no GPU): both arms print the
same `config`, the shard plan restated in bench.py is the package's, the reference arm runs end to end on a
small shape and reports the digest of the index it built."""
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def args(**kw):
    a = types.SimpleNamespace(rows=10_000_000, dim=300, m=30, k=10, queries=100_000, train_rows=262_144,
                              train_iters=8, row_shards=0, gpus=1)
    a.__dict__.update(kw)
    return a


def test_shard_plan_is_the_package_rule():
    from gulon_b200.sharded import shard_plan
    for rows in (1, 9_999_999, 10_000_000, 40_000_000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            assert bench.shard_plan(rows, world) == shard_plan(rows, world)
    assert bench.shard_plan(10_000_000, 8) == (1, 8) and bench.shard_plan(100_000_000, 8) == (8, 1)


def test_config_names_the_workload_and_is_arm_independent():
    c = bench.config_of(args(), 1)
    assert c["workload"].startswith("configs[1]:") and c["sharding"] == "single GPU"
    assert "custom shape" in bench.config_of(args(rows=100_000_000, dim=128, m=16), 8)["workload"]
    c8 = bench.config_of(args(gpus=8), 8)
    assert c8["sharding"].startswith("1 row shard(s) of the code planes x 8 query group(s)")
    assert json.dumps(bench.config_of(args(), 4), sort_keys=True) == json.dumps(bench.config_of(args(), 4), sort_keys=True)


def test_reference_arm_small_shape(tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path), OMP_NUM_THREADS="1")     # torchrun exports OMP_NUM_THREADS=1
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "60000", "--dim", "24",
           "--m", "4", "--train-rows", "8000", "--train-iters", "2", "--steps", "2", "--warmup", "1", "--queries", "500"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))           # not OMP's 1
    assert line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    assert len(line["index"]["digest"]) == 16 and set(line["jvm_probe"]) == {"java", "javac", "scala", "sbt"}
    # other ranks of a torchrun launch print nothing and exit 0
    out2 = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="3"), timeout=60, cwd=ROOT)
    assert out2.returncode == 0 and out2.stdout.strip() == ""


def test_index_digest_is_order_sensitive():
    cb = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    codes = np.arange(20, dtype=np.uint8).reshape(2, 10)
    d = bench.index_digest(cb, codes)
    assert d == bench.index_digest(cb.copy(), codes.copy()) and d != bench.index_digest(cb, codes[::-1])
