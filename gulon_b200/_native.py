"""ctypes binding of libgulon_b200.so (the C ABI declared in include/gulon_b200.h).

The library is built in-tree with nvcc for sm_100a (`build()`); there is no CPU fallback: when the
shared object is missing the import fails loudly, and when no CUDA device is visible every compute
entry point returns GULON_ENODEVICE, raised here as `NoDeviceError`.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, "libgulon_b200.so")
_SRC_DIR = os.path.join(_HERE, "csrc")
_SOURCES = ["gulon_b200.cu"]
_HEADERS = ["common.cuh", "kmeans.cuh", "scan.cuh", "select.cuh", "pscan.cuh", "tcassign.cuh", "tscan.cuh", "kupdate.cuh", "mlctl.h",
            "synth.cuh", "synth_spec.h"]

OK, EINVAL, ECUDA, ENOMEM, ENODEVICE, ECOMM, EUNSUPPORTED, ESTATE = 0, -1, -2, -3, -4, -5, -6, -7
TIE_LOWEST = 1
UPDATE_RUNNING_MEAN, UPDATE_SUM = 0, 1
SCAN_AUTO, SCAN_SIMPLE, SCAN_FUSED, SCAN_PRUNED, SCAN_TENSOR = 0, 1, 2, 3, 4
ASSIGN_AUTO, ASSIGN_EXACT, ASSIGN_TENSOR = 0, 1, 2

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class GulonError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gulon_b200 error %d: %s" % (code, msg))
        self.code = code


class NoDeviceError(GulonError):
    pass


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(_SRC_DIR, f) for f in _SOURCES + _HEADERS]
    deps.append(os.path.join(_ROOT, "include", "gulon_b200.h"))
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into gulon_b200/libgulon_b200.so for sm_100a (nvcc cross-compiles).

    Several processes may get here at once (one rank per GPU under torchrun, each finding the library
    older than a source): the build runs under a file lock, writes to a temporary name and renames it
    into place, and a process that waited for the lock re-checks before compiling again."""
    if not force and not _stale():
        return SO_PATH
    import fcntl
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return SO_PATH          # another process built it while this one waited
            nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
            if not os.path.exists(nvcc):
                nvcc = "nvcc"
            tmp = SO_PATH + ".%d.tmp" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(_SRC_DIR, s) for s in _SOURCES]
            if verbose:
                cmd += ["-Xptxas", "-v"]
                print(" ".join(cmd), file=sys.stderr)
            try:
                subprocess.check_call(cmd, cwd=_SRC_DIR)
                os.replace(tmp, SO_PATH)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


i32, i64 = C.c_int32, C.c_int64
vp = C.c_void_p


class Comm(C.Structure):
    """gulon_comm_t: exchange hooks for the sharded paths."""
    ALLREDUCE_F32 = C.CFUNCTYPE(C.c_int, vp, vp, i64, vp)
    ALLREDUCE_I32 = C.CFUNCTYPE(C.c_int, vp, vp, i64, vp)
    ALLGATHER = C.CFUNCTYPE(C.c_int, vp, vp, vp, i64, vp)
    ALLREDUCE_I64 = C.CFUNCTYPE(C.c_int, vp, vp, i64, vp)
    ALLREDUCE_MAX_F32 = C.CFUNCTYPE(C.c_int, vp, vp, i64, vp)
    _fields_ = [("rank", i32), ("world", i32),
                ("allreduce_sum_f32", ALLREDUCE_F32),
                ("allreduce_sum_i32", ALLREDUCE_I32),
                ("allgather", ALLGATHER),
                ("user", vp),
                ("allreduce_sum_i64", ALLREDUCE_I64),
                ("allreduce_max_f32", ALLREDUCE_MAX_F32)]


class Progress(C.Structure):
    """gulon_progress_t == KMeans.ProgressReport (G/KMeans.scala:119-127)."""
    _fields_ = [("quantizer", i32), ("num_iterations", i32), ("max_iterations", i32),
                ("step_mean", C.c_float), ("step_stddev", C.c_float), ("converged", i32),
                ("step_count", i32), ("step_s", C.c_float)]


PROGRESS_FN = C.CFUNCTYPE(None, vp, C.POINTER(Progress))


class SynthParams(C.Structure):
    """gulon_synth_params_t == gs_params (csrc/synth_spec.h)."""
    _fields_ = [("seed", C.c_uint64), ("D", i32), ("centres", i32), ("latent", i32), ("nonneg", i32),
                ("noise", C.c_float), ("eps", C.c_float), ("span", C.c_float),
                ("inv_sqrt_latent", C.c_float)]


# every symbol include/gulon_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gulon_version": (C.c_int, []),
    "gulon_last_error": (C.c_char_p, []),
    "gulon_device_count": (C.c_int, [C.POINTER(i32)]),
    "gulon_set_device": (C.c_int, [i32]),
    "gulon_get_device": (C.c_int, [C.POINTER(i32)]),
    "gulon_device_sync": (C.c_int, []),
    "gulon_init": (C.c_int, [vp, i32]),
    "gulon_shutdown": (C.c_int, []),
    "gulon_set_option": (C.c_int, [C.c_char_p, i64]),
    "gulon_get_counter": (C.c_int, [C.c_char_p, C.POINTER(i64)]),
    "gulon_debug_tscan": (C.c_int, [vp, vp, i64, i64, vp, i64, i64, vp, vp, vp, C.POINTER(i32)]),
    "gulon_subvectors": (C.c_int, [i32, i32, vp, vp]),
    "gulon_points_create": (C.c_int, [vp, i64, i32, i64, C.POINTER(vp)]),
    "gulon_points_wrap_dev": (C.c_int, [vp, i64, i32, i64, C.POINTER(vp)]),
    "gulon_points_info": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i64),
                                    C.POINTER(vp)]),
    "gulon_points_destroy": (C.c_int, [vp]),
    "gulon_points_normalize": (C.c_int, [vp]),
    "gulon_normalize": (C.c_int, [vp, i64, i32, i64, vp, i64]),
    "gulon_subtract_rows_dev": (C.c_int, [vp, i64, vp, vp, i64, vp, i64, i32, vp, i64, vp]),
    "gulon_kmeans_assign": (C.c_int, [vp, i32, i32, vp, i32, i64, i32, vp]),
    "gulon_kmeans_update": (C.c_int, [vp, i32, i32, vp, i32, i32, vp, vp]),
    "gulon_kmeans_init": (C.c_int, [vp, i32, i32, i32, i32, vp, vp]),
    "gulon_kmeans_train": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(Comm), i64,
                                     i64, PROGRESS_FN, vp, vp, C.POINTER(i32), C.POINTER(i32)]),
    "gulon_pq_train": (C.c_int, [vp, i32, i32, i32, i32, i32, C.POINTER(Comm), i64, i64,
                                 PROGRESS_FN, vp, C.POINTER(vp)]),
    "gulon_codebook_create": (C.c_int, [i32, i32, i32, vp, C.POINTER(vp)]),
    "gulon_codebook_info": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                      C.POINTER(i32)]),
    "gulon_codebook_export": (C.c_int, [vp, vp]),
    "gulon_codebook_destroy": (C.c_int, [vp]),
    "gulon_pq_encode": (C.c_int, [vp, vp, i64, i64, i32, vp]),
    "gulon_pq_encode_dev": (C.c_int, [vp, vp, i64, i64, i32, vp, i64, vp]),
    "gulon_pq_decode": (C.c_int, [vp, vp, i64, i64, vp, i64]),
    "gulon_pq_encode16": (C.c_int, [vp, vp, i64, i64, i32, vp]),
    "gulon_pq_encode16_dev": (C.c_int, [vp, vp, i64, i64, i32, vp, i64, vp]),
    "gulon_pq_decode16": (C.c_int, [vp, vp, i64, i64, vp, i64]),
    "gulon_index_create16": (C.c_int, [vp, vp, i64, i64, C.POINTER(vp)]),
    "gulon_index_create16_dev": (C.c_int, [vp, vp, i64, i64, C.POINTER(vp)]),
    "gulon_index_create": (C.c_int, [vp, vp, i64, i64, C.POINTER(vp)]),
    "gulon_index_create_dev": (C.c_int, [vp, vp, i64, i64, C.POINTER(vp)]),
    "gulon_index_info": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32),
                                   C.POINTER(i32)]),
    "gulon_index_destroy": (C.c_int, [vp]),
    "gulon_prepare_query": (C.c_int, [vp, vp, i64, i64, vp]),
    "gulon_pq_query": (C.c_int, [vp, vp, i64, i64, i32, i64, i64, i32, i64, vp, vp, vp]),
    "gulon_pq_query_dev": (C.c_int, [vp, vp, i64, i64, i32, i64, i64, i32, i64, vp, vp, vp, vp]),
    "gulon_topk_merge_dev": (C.c_int, [vp, vp, i32, i64, i32, vp, vp, vp, vp]),
    "gulon_pq_query_sharded_dev": (C.c_int, [vp, C.POINTER(Comm), C.POINTER(Comm), vp, i64, i64, i32,
                                             i32, i64, vp, vp, vp, vp]),
    "gulon_pq_query_sharded": (C.c_int, [vp, C.POINTER(Comm), C.POINTER(Comm), vp, i64, i64, i32, i32,
                                         i64, vp, vp, vp]),
    "gulon_rerank_dev": (C.c_int, [vp, vp, i64, i64, vp, i32, i32, i64, vp, vp, vp, vp]),
    "gulon_pq_rerank_query_dev": (C.c_int, [vp, vp, C.POINTER(Comm), vp, i64, i64, i32, i32, i32, i64,
                                            vp, vp, vp, vp]),
    "gulon_pq_rerank_query": (C.c_int, [vp, vp, C.POINTER(Comm), vp, i64, i64, i32, i32, i32, i64,
                                        vp, vp, vp]),
    "gulon_synth_tables_dev": (C.c_int, [C.POINTER(SynthParams), vp, vp, vp]),
    "gulon_synth_rows_dev": (C.c_int, [C.POINTER(SynthParams), i64, i64, i64, vp, vp, vp, i64, vp]),
    "gulon_grouped_query_dev": (C.c_int, [vp, vp, i64, i64, vp, i32, vp, vp, vp, vp, i64, i32, i32,
                                          vp, vp, vp, vp]),
    "gulon_exact_topk": (C.c_int, [vp, vp, i64, i64, i32, i64, i64, vp, vp, vp]),
    "gulon_rerank": (C.c_int, [vp, vp, i64, i64, vp, i32, i32, vp, vp, vp]),
}

_lib = None


def lib():
    """Load the shared library (building it first if the sources are newer)."""
    global _lib
    if _lib is None:
        if _stale():
            build()
        if not os.path.exists(SO_PATH):
            raise ImportError("libgulon_b200.so is missing and could not be built; "
                              "gulon_b200 has no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Map a negative status to the exception class the reference would throw."""
    if rc >= 0:
        return rc
    msg = (lib().gulon_last_error() or b"").decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)            # IllegalArgumentException (require(...))
    if rc == ESTATE:
        raise RuntimeError(msg)          # IllegalStateException
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == ENODEVICE:
        raise NoDeviceError(rc, msg)
    raise GulonError(rc, msg)


def set_option(name, value):
    check(lib().gulon_set_option(name.encode(), int(value)))


def kernel_launches():
    v = i64(0)
    check(lib().gulon_get_counter(b"kernel_launches", C.byref(v)))
    return int(v.value)


def counter(name):
    v = i64(0)
    check(lib().gulon_get_counter(name.encode(), C.byref(v)))
    return int(v.value)


def device_count():
    n = i32(0)
    check(lib().gulon_device_count(C.byref(n)))
    return int(n.value)
