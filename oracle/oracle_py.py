"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp) — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm.
The product package (parca_agent_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from parca_agent_b200 import abi  # noqa: E402  (struct layouts only)

_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.PaAggConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_register_strings.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_register_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.orc_register_labelsets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_set_external_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_ingest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_flush.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_stacktraces.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_fix_truncation.restype = C.c_int64
        L.orc_fix_truncation.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
        L.orc_funcdict_new.restype = C.c_void_p
        L.orc_funcdict_append.restype = C.c_uint32
        L.orc_funcdict_append.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64]
        L.orc_funcdict_len.argtypes = [C.c_void_p]
        L.orc_funcdict_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


STAT_NAMES = ("rows", "unique_stacks", "locations", "functions", "location_indices", "empty_samples")


class Oracle:
    """One reporter instance: register the workload's tables once, then ingest/flush batches."""

    def __init__(self, w, stack_cache_entries=0):
        L = lib()
        cfg = abi.PaAggConfig(abi_version=abi.PA_ABI_VERSION, device=0, hash_mode=w.hash_mode, label_flags=w.label_flags,
                              samples_per_second=w.samples_per_second, schema=getattr(w, "schema", 0),
                              unknown_frame_type_sid=getattr(w, "unknown_frame_type_sid", 0), stack_cache_entries=stack_cache_entries)
        self.h = L.orc_create(C.byref(cfg))
        assert self.h
        blob, offs = abi.pack_strings(w.strings[1:])  # id 0 == "" is implicit
        first = C.c_uint32()
        L.orc_register_strings(self.h, blob, offs.ctypes.data, len(w.strings) - 1, C.byref(first))
        assert first.value == 1
        fr = np.ascontiguousarray(w.frames)
        L.orc_register_frames(self.h, fr.ctypes.data, len(fr), None)
        pairs, loffs = abi.pack_labelsets(w.labelsets)
        L.orc_register_labelsets(self.h, pairs.ctypes.data, loffs.ctypes.data, len(w.labelsets), None)
        if w.external_labels:
            ext = np.asarray(w.external_labels, dtype=np.uint32).reshape(-1, 2)
            L.orc_set_external_labels(self.h, ext.ctypes.data, len(ext))

    def ingest(self, hdrs, frame_ids):
        hdrs = np.ascontiguousarray(hdrs)
        frame_ids = np.ascontiguousarray(frame_ids, dtype=np.uint64)
        self._keep = (hdrs, frame_ids)  # the stacks map keeps pointers until flush (like the Go LRU keeps Frames)
        lib().orc_ingest(self.h, hdrs.ctypes.data, frame_ids.ctypes.data, len(hdrs))

    def flush(self):
        p = C.POINTER(C.c_uint8)()
        n = C.c_uint64()
        st = (C.c_uint64 * 6)()
        lib().orc_flush(self.h, C.byref(p), C.byref(n), st)
        data = C.string_at(p, n.value) if n.value else b""
        return data, dict(zip(STAT_NAMES, [int(x) for x in st]))

    def stacktraces(self, ids):
        """v1: buildStacktraceRecord for the concatenated 16-byte ids; returns (ipc bytes, n_locations)."""
        ids = bytes(ids)
        assert len(ids) % 16 == 0
        p, n, nl = C.POINTER(C.c_uint8)(), C.c_uint64(), C.c_uint64()
        lib().orc_stacktraces(self.h, ids, len(ids) // 16, C.byref(p), C.byref(n), C.byref(nl))
        return C.string_at(p, n.value), int(nl.value)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(w):
    """Whole-batch convenience: returns (ipc_bytes, stats)."""
    o = Oracle(w)
    o.ingest(w.hdrs, w.frame_ids)
    out = o.flush()
    o.close()
    return out


def xxh64(data, seed=0):
    return int(lib().orc_xxh64(data, len(data), seed))


def fix_truncation(s, max_len):
    n = lib().orc_fix_truncation(s, len(s), max_len)
    return (None, False) if n < 0 else (s[:n], True)
