"""TEST SCAFFOLDING: builds and drives the host emulation of the inflate kernel
(tests/emu/emu_inflate.cpp compiles parallelparsing_b200/csrc/inflate_core.cuh with
PP_HOST_EMU).  Used by the CPU tests to check the decoder logic against zlib; never
used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_inflate.cpp")
CORE = os.path.join(ROOT, "parallelparsing_b200", "csrc", "inflate_core.cuh")
CORE2 = os.path.join(ROOT, "parallelparsing_b200", "csrc", "blockscan_core.cuh")
OUT = os.path.join(ROOT, "tests", "emu", "_build")

_libs = {}


def lib(subw=31):
    if subw in _libs:
        return _libs[subw]
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, f"emu_inflate_w{subw}.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(CORE2)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", f"-DPP_SUBW={subw}", "-o", so, SRC])
    L = C.CDLL(so)
    L.emu_inflate_chunk.restype = C.c_int
    L.emu_inflate_chunk.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    L.emu_stats.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    L.emu_scan_segment.restype = C.c_int
    L.emu_scan_segment.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32,
                                   C.POINTER(C.c_uint64)]
    L.emu_probe_dynamic_header.restype = C.c_int
    L.emu_probe_dynamic_header.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    _libs[subw] = L
    return L


def inflate_chunk(gz: np.ndarray, in_byte: int, bits: int, in_limit: int, window: np.ndarray, out_len: int,
                  T=64, subw=31):
    """Emulated Core.ExtractDeflateIndex: the stream starts `bits` bits before byte `in_byte`
    of gz; returns (status, bytes produced, newlines, min_byte, end_bit)."""
    L = lib(subw)
    pad = (-gz.size) % 16
    comp = np.concatenate([gz, np.zeros(pad + 64, np.uint8)])
    lead_len = window.size
    slot = np.zeros(lead_len + out_len + 256, np.uint8)
    res = (C.c_uint64 * 4)()
    st = L.emu_inflate_chunk(T, comp.ctypes.data, comp.size - 64 + 0, in_byte * 8 - bits, in_limit,
                             slot.ctypes.data, window.ctypes.data, lead_len, out_len, res)
    return st, slot[lead_len: lead_len + int(res[0])], int(res[1]), int(res[2]), int(res[3])


def inflate_chunk_pull(gz: np.ndarray, in_byte: int, bits: int, in_limit: int, window: np.ndarray, out_len: int,
                       T=64, subw=31, exact_extent=False):
    """inflate_chunk<PULL>: as inflate_chunk, every window re-using the staged bytes the window before it left.
    exact_extent: the compressed buffer ends where gz ends (no padding behind it, as a caller's pinned buffer)."""
    L = lib(subw)
    L.emu_inflate_chunk_pull.restype = C.c_int
    L.emu_inflate_chunk_pull.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    pad = (-gz.size) % 16
    comp = np.concatenate([gz, np.zeros(pad + 64, np.uint8)])
    extent = gz.size if exact_extent else comp.size - 64
    lead_len = window.size
    slot = np.zeros(lead_len + out_len + 256, np.uint8)
    res = (C.c_uint64 * 4)()
    st = L.emu_inflate_chunk_pull(T, comp.ctypes.data, extent, in_byte * 8 - bits, in_limit, slot.ctypes.data,
                                  window.ctypes.data, lead_len, out_len, res)
    return st, slot[lead_len: lead_len + int(res[0])], int(res[1]), int(res[2]), int(res[3])


def inflate_chunk_dual(gz: np.ndarray, in_byte: int, bits: int, in_limit: int, window_a: np.ndarray, window_b: np.ndarray,
                       out_len: int, T=64, subw=31):
    """One decode, two resolves (inflate_chunk<DUAL>): returns (status, bytes against window_a, bytes against window_b)."""
    L = lib(subw)
    L.emu_inflate_chunk_dual.restype = C.c_int
    L.emu_inflate_chunk_dual.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    pad = (-gz.size) % 16
    comp = np.concatenate([gz, np.zeros(pad + 64, np.uint8)])
    lead_len = window_a.size
    delta = (lead_len + out_len + 1 + 127) // 128 * 128
    slot = np.zeros(2 * delta + 256, np.uint8)
    lead = np.ascontiguousarray(np.concatenate([window_a, window_b]))
    res = (C.c_uint64 * 4)()
    st = L.emu_inflate_chunk_dual(T, comp.ctypes.data, comp.size - 64, in_byte * 8 - bits, in_limit, slot.ctypes.data, delta,
                                  lead.ctypes.data, lead_len, out_len, res)
    n = int(res[0])
    return st, slot[lead_len: lead_len + n], slot[delta + lead_len: delta + lead_len + n]


def stats(subw=31, reset=True):
    out = (C.c_uint64 * 8)()
    lib(subw).emu_stats(out, int(reset))
    return dict(windows=int(out[0]), rounds=int(out[1]), max_rounds=int(out[2]), blocks=int(out[3]))


def _padded(gz):
    return np.concatenate([gz, np.zeros((-gz.size) % 16 + 64, np.uint8)])


def scan_segment(gz: np.ndarray, start_bit: int, end_bit: int, search: bool, T=64, rec_cap=4096):
    """Emulated block scanner over one segment: (status, first_bit, land_bit, out_bytes, recs[n,2])."""
    L = lib()
    comp = _padded(gz)
    recs = np.zeros((rec_cap, 2), np.uint64)
    res = (C.c_uint64 * 5)()
    st = L.emu_scan_segment(T, comp.ctypes.data, comp.size - 64, start_bit, end_bit, int(search), recs.ctypes.data, rec_cap, res)
    return st, int(res[0]), int(res[1]), int(res[2]), recs[: int(res[3])].astype(np.int64)


def probe_dynamic_header(gz: np.ndarray, bit: int) -> bool:
    comp = _padded(gz)
    return bool(lib().emu_probe_dynamic_header(comp.ctypes.data, comp.size - 64, bit))
