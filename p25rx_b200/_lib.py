"""ctypes loader for libp25cu.so (the C ABI declared in include/p25cu.h).

There is no fallback: if the library is missing or no sm_100 GPU is usable, importing
callers get an exception that says so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P25CU_LIB") or os.path.join(HERE, "libp25cu.so")   # P25CU_LIB: A/B builds of the same library
ABI_VERSION = 1

FMT_U8_IQ, FMT_CF32_IQ = 0, 1

EVENT_DTYPE = np.dtype([("stream", "<u4"), ("kind", "<u4"), ("sample", "<u8"), ("len", "<u4"),
                        ("payload", "u1", (60,))], align=True)
assert EVENT_DTYPE.itemsize == 80

# names every build must export (checked by tests/test_abi.py against include/p25cu.h)
EXPORTS = ["p25cu_create", "p25cu_destroy", "p25cu_last_error", "p25cu_demod", "p25cu_decode", "p25cu_process",
           "p25cu_poll", "p25cu_poll_view", "p25cu_pending", "p25cu_poll_start", "p25cu_poll_packed", "p25cu_unpack_events",
           "p25cu_host_alloc", "p25cu_host_free", "p25cu_host_register", "p25cu_host_unregister", "p25cu_resync", "p25cu_get_stats", "p25cu_cuda_stream", "p25cu_sync", "p25cu_set_overlap",
           "p25cu_launch_count", "p25cu_demod_timing", "p25cu_device_baseband", "p25cu_read_baseband", "p25cu_channelizer_output", "p25cu_set_keep_spectra", "p25cu_fec_selftest"]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_streams", C.c_uint32), ("format", C.c_int32), ("decimation", C.c_int32),
                ("max_chunk_samples", C.c_uint64), ("max_baseband", C.c_uint64), ("abi_version", C.c_uint32),
                ("event_slots", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("code", C.c_uint64 * 4 * 12)]


class P25Error(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"p25cu status {status}: {text}")
        self.status = status


def build(verbose: bool = False) -> str:
    """Compile libp25cu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc")], stdout=out)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  p25rx_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.p25cu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.p25cu_destroy.argtypes = [vp]
    L.p25cu_destroy.restype = None
    L.p25cu_last_error.argtypes = [vp]
    L.p25cu_last_error.restype = C.c_char_p
    L.p25cu_demod.argtypes = [vp, vp, sz, i, vp, C.POINTER(sz), vp]
    L.p25cu_decode.argtypes = [vp, vp, sz]
    L.p25cu_process.argtypes = [vp, vp, sz, i]
    L.p25cu_poll.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.p25cu_poll_view.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.p25cu_pending.argtypes = [vp, C.POINTER(sz)]
    L.p25cu_poll_start.argtypes = [vp]
    L.p25cu_poll_packed.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz), C.POINTER(i)]
    L.p25cu_unpack_events.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.p25cu_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.p25cu_host_free.argtypes = [vp, vp]
    L.p25cu_host_register.argtypes = [vp, vp, sz]
    L.p25cu_host_unregister.argtypes = [vp, vp]
    L.p25cu_resync.argtypes = [vp, C.c_uint32]
    L.p25cu_get_stats.argtypes = [vp, C.c_uint32, C.POINTER(Stats), i]
    L.p25cu_cuda_stream.argtypes = [vp]
    L.p25cu_cuda_stream.restype = vp
    L.p25cu_sync.argtypes = [vp]
    L.p25cu_set_overlap.argtypes = [vp, i]
    L.p25cu_launch_count.argtypes = [vp]
    L.p25cu_launch_count.restype = C.c_uint64
    L.p25cu_demod_timing.argtypes = [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_uint)]
    L.p25cu_device_baseband.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz)]
    L.p25cu_read_baseband.argtypes = [vp, C.c_uint32, vp, sz]
    L.p25cu_set_keep_spectra.argtypes = [vp, i]
    L.p25cu_channelizer_output.argtypes = [vp, vp, C.POINTER(sz)]
    L.p25cu_fec_selftest.argtypes = [vp, i, vp, sz, i, i, vp, vp]
    _lib = L
    return L
