"""gulon_b200 -- B200-native (sm_100a) implementation of tixxit/gulon's product-quantization hot
path: k-means codebook training, PQ encoding, ADC scan + top-k.  The classes here mirror the
reference's Scala API (G/KMeans.scala, G/ProductQuantizer.scala, G/Index.scala) and call
hand-written CUDA kernels through the C ABI in include/gulon_b200.h.  No CPU fallback.
"""
from . import _native
from ._native import (GulonError, NoDeviceError, SCAN_AUTO, SCAN_FUSED, SCAN_PRUNED, SCAN_SIMPLE, SCAN_TENSOR, TIE_LOWEST,
                      UPDATE_RUNNING_MEAN, UPDATE_SUM, build, device_count, kernel_launches,
                      set_option)
from .grouped import GroupedIndex, GroupedVectors, LimitGroups, LimitVectors
from .index import PQIndex, TopK, exact_nearest_neighbours, prepare_query
from .kmeans import Config as KMeansConfig
from .kmeans import KMeans
from .kmeans import ProgressReport as KMeansProgressReport
from .coder import BytePlus, Coder0, Coder2, Coder4, factory_for
from .coder import coder as make_coder
from . import coder
from .quantizer import Coder8, EncodedMatrix, ProductQuantizer, Quantizer, coder_width
from .quantizer import Config as ProductQuantizerConfig
from .storage import SortedIndex
from .recall import SummaryStats, Tests
from .vectors import DevicePoints, Matrix, Vectors, normalize, subvector_windows

__all__ = [
    "GulonError", "NoDeviceError", "SCAN_AUTO", "SCAN_FUSED", "SCAN_PRUNED", "SCAN_SIMPLE", "SCAN_TENSOR", "TIE_LOWEST",
    "UPDATE_RUNNING_MEAN", "UPDATE_SUM", "build", "device_count", "kernel_launches", "set_option",
    "GroupedIndex", "GroupedVectors", "LimitGroups", "LimitVectors", "PQIndex", "TopK", "exact_nearest_neighbours", "prepare_query", "KMeans", "KMeansConfig",
    "KMeansProgressReport", "Coder8", "EncodedMatrix", "ProductQuantizer", "Quantizer",
    "coder_width", "ProductQuantizerConfig", "DevicePoints", "Matrix", "Vectors", "normalize",
    "subvector_windows", "SortedIndex", "BytePlus", "Coder0", "Coder2", "Coder4", "make_coder", "factory_for", "SummaryStats", "Tests",
]
