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
       os.path.join(ROOT, "paragraph_b200", "csrc", "pg_kmer.cuh"), os.path.join(ROOT, "paragraph_b200", "csrc", "pg_path.cuh"),
       os.path.join(ROOT, "paragraph_b200", "csrc", "pg_host.hpp"),
       os.path.join(ROOT, "paragraph_b200", "csrc", "pg_count.cuh")]
CIGAR_STRIDE = 4096
_lib = None


def build():
    if os.environ.get("PG_EMU_SO"):  # a prebuilt variant (e.g. another -DPG_CK): tools / A-B only
        return
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in SRC):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", SO, SRC[0]])


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(os.environ.get("PG_EMU_SO") or SO)
        l.pgemu_align_batch.restype = C.c_int
        l.pgemu_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_int, C.c_char_p, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_uint8), C.c_uint, C.POINTER(C.c_int32), C.c_char_p, C.c_char_p,
                                        C.c_int, C.POINTER(C.c_int64)]
        l.pgemu_set_geometry.argtypes = [C.c_int]
        _lib = l
    return _lib


def set_spec(on):
    """speculative "no gap alive" blocks (pg_core.cuh: lane_step_dead), what the kernels do when built with PG_SPEC_DEAD=1"""
    lib().pgemu_set_spec(int(on))  # 2 = EXPERIMENT: plus upper-bound pruning of gaps that cannot reach the best score so far


def set_dead_boundary(on):
    """node-boundary sub-blocks may run with the collapsed recurrence as well (the kernels' PG_DEAD_BOUNDARY)"""
    lib().pgemu_set_dead_boundary(int(on))


def dead_boundary_stats():
    """boundary sub-blocks so far: [run dead, attempted and redone]"""
    o = (C.c_long * 2)()
    lib().pgemu_dead_boundary_stats(o)
    return list(o)


def spec_stats():
    """blocks of SPEC_STEPS steps so far: [run dead, redone, gaps alive at the start, node boundary inside]"""
    o = (C.c_long * 4)()
    lib().pgemu_spec_stats(o)
    return list(o)


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


def emu_count_site(node_seqs, edges, edge_labels, reads, is_rev=None, fragment=None, remove_nonuniq=True,
                   bad_align_frac=0.8, use_filters=True, family_slots=256):
    """Emulated alignment + counting stage (pg_count.cuh on the host) of one site.  Returns a dict shaped like
    oracle.refbind.oracle_count_site's plus the records' CIGAR strings ("cigars")."""
    from oracle import refbind as R
    l = lib()
    i32p, u8p, u32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    l.pgemu_count_site.restype = C.c_int
    l.pgemu_count_site.argtypes = [C.c_int, C.c_char_p, i32p, C.c_int, i32p, i32p, u64p, C.c_int, C.c_char_p, i32p, u8p,
                                   i32p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, u32p, u32p, C.c_int, i32p,
                                   u32p, u32p, u32p, C.c_int, i32p]
    blob = "".join(node_seqs).encode("latin-1")
    off = np.zeros(len(node_seqs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in node_seqs])
    ef = np.array([e[0] for e in edges], dtype=np.int32)
    et = np.array([e[1] for e in edges], dtype=np.int32)
    lab = np.ascontiguousarray(edge_labels if edge_labels is not None else np.zeros(len(edges)), dtype=np.uint64)
    rblob = "".join(reads).encode("latin-1")
    roff = np.zeros(len(reads) + 1, dtype=np.int32)
    roff[1:] = np.cumsum([len(s) for s in reads])
    n, nn, ne = len(reads), len(node_seqs), len(edges)
    rv = np.ascontiguousarray(is_rev if is_rev is not None else np.zeros(n), dtype=np.uint8)
    fr = np.ascontiguousarray(fragment if fragment is not None else np.arange(n), dtype=np.int32)
    sup = np.zeros(n, dtype=R.SUPPORT_DTYPE)
    cap = int(roff[-1]) + (2 * nn + 8) * n + 16
    pw = np.zeros(cap, dtype=np.uint32)
    ops = np.zeros(cap, dtype=np.uint32)
    nc = np.zeros((max(nn, 1), 4), dtype=np.uint32)
    ec = np.zeros((max(ne, 1), 4), dtype=np.uint32)
    fcap = (4 + 4 * (1 + nn + ne)) * family_slots
    fw = np.zeros(fcap, dtype=np.uint32)
    used = np.zeros(2, dtype=np.int32)
    rc = l.pgemu_count_site(nn, blob, _p(off, C.c_int32), ne, _p(ef, C.c_int32), _p(et, C.c_int32), _p(lab, C.c_uint64),
                            n, rblob, _p(roff, C.c_int32), _p(rv, C.c_uint8), _p(fr, C.c_int32), int(remove_nonuniq),
                            float(bad_align_frac), int(use_filters), family_slots, sup.ctypes.data, _p(pw, C.c_uint32),
                            _p(ops, C.c_uint32), cap, _p(used[0:], C.c_int32), _p(nc, C.c_uint32), _p(ec, C.c_uint32),
                            _p(fw, C.c_uint32), fcap, _p(used[1:], C.c_int32))
    if rc != 0:
        raise RuntimeError("pgemu_count_site rc=%d" % rc)
    return dict(support=sup, path_words=pw[:used[0]], ops=ops[:used[0]], node_counts=nc[:nn].astype(np.int64),
                edge_counts=ec[:ne].astype(np.int64), families=unpack_family_entries(fw[:used[1]]))


def unpack_family_entries(words):
    """C-ABI family words {site, n, mask_lo, mask_hi, n x 4} -> {mask: [n, 4]} (single-site use) """
    out, w = {}, 0
    while w < len(words):
        n = int(words[w + 1])
        out[int(words[w + 2]) | (int(words[w + 3]) << 32)] = np.array(words[w + 4:w + 4 + 4 * n], dtype=np.int64).reshape(n, 4)
        w += 4 + 4 * n
    return out


def emu_path_align_batch(node_seqs, edges, reads, kmer_len=32):
    """The exact-match stage (paragraph_b200/csrc/pg_path.cuh) on the host.  -> (dicts like refbind.ref_path_align_batch,
    (attempted, anchored, mapped))"""
    l = lib()
    l.pgemu_path_align_batch.restype = C.c_int
    l.pgemu_path_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
    blob = "".join(node_seqs).encode("latin-1")
    off = np.zeros(len(node_seqs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in node_seqs])
    ef = np.array([e[0] for e in edges], dtype=np.int32)
    et = np.array([e[1] for e in edges], dtype=np.int32)
    rblob = "".join(reads).encode("latin-1")
    roff = np.zeros(len(reads) + 1, dtype=np.int32)
    roff[1:] = np.cumsum([len(s) for s in reads])
    n = len(reads)
    out = np.zeros((n, 8), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    cnt = np.zeros(3, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = l.pgemu_path_align_batch(len(node_seqs), blob, p(off), len(edges), p(ef), p(et), int(kmer_len), n, rblob, p(roff),
                                  p(out), ob, cg, CIGAR_STRIDE, p(cnt))
    if rc != 0:
        raise RuntimeError("pgemu_path_align_batch rc=%d" % rc)
    raw = ob.raw
    res = []
    for i in range(n):
        c = cg.raw[i * CIGAR_STRIDE:(i + 1) * CIGAR_STRIDE].split(b"\0", 1)[0].decode()
        o = out[i]
        res.append(dict(mapped=bool(o[0]), pos=int(o[1]), score=int(o[2]), unique=bool(o[3]), mapq=int(o[4]),
                        graph_reverse=bool(o[5]), bases=raw[roff[i]:roff[i + 1]].decode("latin-1"), cigar=c))
    return res, tuple(int(x) for x in cnt)


def emu_kmer_align_batch(node_seqs, edges, paths, reads, kmer_len=16, is_rev=None):
    """grm::KmerAligner<k> through the device source (pg_kmer.cuh) -> result dicts like refbind.ref_kmer_align_batch"""
    from oracle import refbind as R
    l = lib()
    l.pgemu_kmer_align_batch.restype = C.c_int
    blob, off, ef, et = R.pack_graph(node_seqs, edges)
    pptr, pnodes = R.pack_paths(paths)
    rblob, roff = R.pack_reads(reads)
    n = len(reads)
    out = np.zeros((max(n, 1), 8), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
    rc = l.pgemu_kmer_align_batch(C.c_int(len(node_seqs)), blob, _p(off, C.c_int32), C.c_int(len(edges)), _p(ef, C.c_int32),
                                  _p(et, C.c_int32), C.c_int(len(paths)), _p(pptr, C.c_int32), _p(pnodes, C.c_int32),
                                  C.c_int(int(kmer_len)), C.c_int(n), rblob, _p(roff, C.c_int32),
                                  None if rv is None else _p(rv, C.c_uint8), _p(out, C.c_int32), ob, cg, C.c_int(CIGAR_STRIDE))
    if rc != 0:
        raise RuntimeError("emulator: pgemu_kmer_align_batch rc=%d" % rc)
    return R._kmer_collect(out, ob, cg, roff, n)
