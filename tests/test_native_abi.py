"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/gulon_b200.h declares, refuses to compute without a device (no CPU fallback),
and the host logic that needs no device agrees with the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from gulon_b200 import _native
    _native.build()
    return _native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gulon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gulon_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(native):
    lib = C.CDLL(native.SO_PATH)
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_binding_covers_header(native):
    assert sorted(native.SIGNATURES) == declared_symbols()


def test_version_and_error_channel(native):
    lib = native.lib()
    assert lib.gulon_version() == 100
    with pytest.raises(ValueError):
        native.set_option("no_such_option", 1)
    assert b"no_such_option" in lib.gulon_last_error()


def test_split_rule_matches_oracle(native, oracle):
    from gulon_b200 import subvector_windows
    for D, M in [(100, 10), (300, 30), (128, 16), (1000, 100), (37, 5), (7, 7), (10, 3), (5, 1)]:
        f, d, dmax = subvector_windows(D, M)
        of, od, odmax = oracle.subvectors(D, M)
        assert np.array_equal(f, of) and np.array_equal(d, od) and dmax == odmax
        assert d.sum() == D and d.max() - d.min() <= 1      # T/VectorsSpec.scala:42-65


def test_coder_width():
    from gulon_b200 import coder_width
    # width of ProductQuantizer#coderFactory: 32 - nlz(K - 1) rounded up to a supported coder
    assert [coder_width(k) for k in (1, 2, 3, 4, 16, 17, 255, 256)] == [0, 2, 2, 2, 4, 8, 8, 8]
    with pytest.raises(ValueError):
        coder_width(65537)   # "too many clusters", G/ProductQuantizer.scala:13-15


def test_tensor_scan_options_and_counters(native):
    """The tensor scan's knobs (header: gulon_set_option) exist, validate their ranges and never need a device."""
    import gulon_b200 as g
    from gulon_b200 import _native as N
    assert g.SCAN_TENSOR == 4
    defaults = {"tensor_min_rows": 1 << 16, "tensor_min_queries": 256, "tensor_min_pairs": 1 << 27, "tensor_query_batch": 0, "tensor_stage_ratio": 0,
                "tensor_boot_rows": 0, "tensor_max_bytes": 64 << 30, "tensor_chunk_bytes": 16 << 20, "tensor_pair": 1,
                "tensor_epi_wait": 2}
    for name, v in defaults.items():
        g.set_option(name, v)
        with pytest.raises(ValueError):
            g.set_option(name, -1)
    with pytest.raises(ValueError):
        g.set_option("tensor_epi_wait", 8)
    g.set_option("scan_impl", g.SCAN_TENSOR)
    g.set_option("scan_impl", g.SCAN_AUTO)
    with pytest.raises(ValueError):
        g.set_option("scan_impl", 5)
    for c in ("tscan_kernel_ns", "tscan_kernel_launches", "tscan_tiles", "tscan_survivors", "tscan_candidates",
              "tscan_pairs", "tscan_fallbacks", "tscan_batches", "tscan_stages", "tscan_handed_back_queries", "scan_last_impl"):
        assert N.counter(c) >= 0


def test_no_cpu_fallback(native):
    import gulon_b200 as g
    if g.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(g.NoDeviceError):
        g.Matrix(np.zeros((4, 4), np.float32)).device()
    with pytest.raises(g.NoDeviceError):
        g.ProductQuantizer.from_codebook(np.zeros((2, 4, 2), np.float32), 4)


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gulon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # comments may NAME the checker (e.g. "CPU twin: oracle/synth.c"); nothing may import,
                # include, link or load it
                for pat in (r"^\s*(import|from)\s+oracle\b", r"#\s*include\s*[\"<][^\n]*oracle",
                            r"libgulon_oracle", r"(CDLL|dlopen|LoadLibrary)\([^\n]*oracle", r"\bgo_[a-z_0-9]+\s*\("):
                    assert not re.search(pat, text, re.M), (f, pat)
