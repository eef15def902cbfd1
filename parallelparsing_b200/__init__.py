"""parallelparsing_b200 — B200-native (sm_100a) drop-in for the checkpointed
gzip-FASTQ decode path of Quantumzhao/ParallelParsing.

The product is libppb200.so (C ABI: include/ppb200.h) — hand-written CUDA kernels
plus a C++ host runtime.  This package is the host-side mirror of the reference's
interface for that path (see api.py) and the in-tree build driver (build.py).
"""
from .api import (BatchedFASTQ, Core, Device, FastqRecord, Index, IndexIO, Job, MultiGpuDecompressAll, PairedDecompressAll, PairedFASTQ, Parsing,
                  Point, ZException, partition_chunks,
                  fields_from_line_starts, pinned_copy)
from ._lib import LIB_PATH, SYMBOLS, check, lib

__all__ = ["BatchedFASTQ", "Core", "Device", "FastqRecord", "Index", "IndexIO", "Job", "MultiGpuDecompressAll", "PairedDecompressAll", "PairedFASTQ",
           "Parsing", "Point", "partition_chunks",
           "ZException", "fields_from_line_starts", "pinned_copy", "LIB_PATH", "SYMBOLS", "lib"]
