"""TEST INFRASTRUCTURE: builds and binds tests/emu/libpgemu.so, the CPU lane emulator that compiles the
device source (paragraph_b200/csrc/pg_core.cuh) for the host."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "emu", "libpgemu.so")
SRC = [os.path.join(HERE, "emu", "pg_emu.cpp"), os.path.join(ROOT, "paragraph_b200", "csrc", "pg_core.cuh"),
       os.path.join(ROOT, "paragraph_b200", "csrc", "pg_host.hpp")]
CIGAR_STRIDE = 1024
_lib = None


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in SRC):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", SO, SRC[0]])


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(SO)
        l.pgemu_align_batch.restype = C.c_int
        l.pgemu_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_int, C.c_char_p, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_uint8), C.c_uint, C.POINTER(C.c_int32), C.c_char_p, C.c_char_p,
                                        C.c_int, C.POINTER(C.c_int64)]
        l.pgemu_set_geometry.argtypes = [C.c_int]
        _lib = l
    return _lib


def set_geometry(w):
    """lanes per task (32, 16 or 8) used by the emulated kernels"""
    lib().pgemu_set_geometry(int(w))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def emu_align_batch(node_seqs, edges, reads, is_rev=None, flags=0xFFFFFFFF):
    l = lib()
    blob = "".join(node_seqs).encode("latin-1")
    off = np.zeros(len(node_seqs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in node_seqs])
    ef = np.array([e[0] for e in edges], dtype=np.int32)
    et = np.array([e[1] for e in edges], dtype=np.int32)
    rblob = "".join(reads).encode("latin-1")
    roff = np.zeros(len(reads) + 1, dtype=np.int32)
    roff[1:] = np.cumsum([len(s) for s in reads])
    n = len(reads)
    out = np.zeros((n, 6), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
    tiles = np.zeros(1, dtype=np.int64)
    rc = l.pgemu_align_batch(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32), _p(et, C.c_int32),
                             n, rblob, _p(roff, C.c_int32), None if rv is None else _p(rv, C.c_uint8),
                             flags & 0xFFFFFFFF, _p(out, C.c_int32), ob, cg, CIGAR_STRIDE, _p(tiles, C.c_int64))
    if rc != 0:
        raise RuntimeError("pgemu_align_batch rc=%d" % rc)
    res, raw = [], ob.raw
    for i in range(n):
        c = cg.raw[i * CIGAR_STRIDE:(i + 1) * CIGAR_STRIDE].split(b"\0", 1)[0].decode()
        res.append(dict(pos=int(out[i, 0]), score=int(out[i, 1]), unique=bool(out[i, 2]), mapq=int(out[i, 3]),
                        graph_reverse=bool(out[i, 4]), bases=raw[roff[i]:roff[i + 1]].decode("latin-1"), cigar=c,
                        status=int(out[i, 5]) & 0xFF, clipped=int(out[i, 5]) >> 8))
    return res, int(tiles[0])
