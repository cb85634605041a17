"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref/libpgref.so (the unmodified
reference sources compiled by oracle/Makefile) and oracle/liboracle.so (the plain-C
restatement, oracle/pg_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpgref.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

AF_CIGAR, AF_BOTH_STRANDS, AF_REVERSE_GRAPH, AF_ALL = 1, 2, 4, 0xFFFFFFFF
CIGAR_STRIDE = 4096


def build(ref=True):
    """make liboracle.so (+ _ref/libpgref.so when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all" if ref else os.path.join(HERE, "liboracle.so")])


def have_ref():
    return os.path.exists(REF_SO)


def pack_graph(node_seqs, edges):
    """-> (seq_blob bytes, seq_off int32[n+1], efrom int32[], eto int32[])"""
    blob = "".join(node_seqs).encode("latin-1")
    off = np.zeros(len(node_seqs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in node_seqs])
    ef = np.array([e[0] for e in edges], dtype=np.int32)
    et = np.array([e[1] for e in edges], dtype=np.int32)
    return blob, off, ef, et


def pack_reads(reads):
    blob = "".join(reads).encode("latin-1")
    off = np.zeros(len(reads) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in reads])
    return blob, off


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.pgref_aligner_create.restype = C.c_void_p
        lib.pgref_aligner_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.pgref_aligner_destroy.argtypes = [C.c_void_p]
        lib.pgref_aligner_align.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_uint,
                                            C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int]
        lib.pgref_align_batch.restype = C.c_int
        lib.pgref_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                          C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                          C.c_uint, C.c_int, C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int]
        i32p = C.POINTER(C.c_int32)
        lib.pgref_align_sites.restype = C.c_int
        lib.pgref_align_sites.argtypes = [C.c_int, i32p, C.c_char_p, i32p, i32p, i32p, i32p, i32p, C.c_char_p, i32p,
                                          C.POINTER(C.c_uint8), C.c_uint, C.c_int, i32p, C.c_char_p, C.c_char_p, C.c_int]
        if hasattr(lib, "pgref_kmer_align_batch"):
            lib.pgref_kmer_align_batch.restype = C.c_int
            lib.pgref_kmer_align_batch.argtypes = [C.c_int, C.c_char_p, i32p, C.c_int, i32p, i32p, C.c_int, i32p, i32p, C.c_int,
                                                   C.c_int, C.c_char_p, i32p, C.POINTER(C.c_uint8), i32p, C.c_char_p,
                                                   C.c_char_p, C.c_int, i32p]
        lib.pgref_gssw_create.restype = C.c_void_p
        lib.pgref_gssw_create.argtypes = lib.pgref_aligner_create.argtypes
        lib.pgref_gssw_destroy.argtypes = [C.c_void_p]
        lib.pgref_gssw_fill_trace.restype = C.c_int
        lib.pgref_gssw_fill_trace.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                              C.POINTER(C.c_int32), C.c_char_p, C.c_int]
        lib.pgref_gssw_fill_trace16.restype = C.c_int
        lib.pgref_gssw_fill_trace16.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint16),
                                                C.POINTER(C.c_int32), C.c_char_p, C.c_int]
        _ref = lib
    return _ref


def _result_dict(o, bases, cigar):
    return dict(pos=int(o[0]), score=int(o[1]), unique=bool(o[2]), mapq=int(o[3]),
                graph_reverse=bool(o[4]), bases=bases, cigar=cigar)


def ref_align_batch(node_seqs, edges, reads, is_rev=None, flags=AF_ALL, threads=1):
    """Run the reference GraphAligner::alignRead over a batch. Returns list of dicts."""
    lib = ref_lib()
    blob, off, ef, et = pack_graph(node_seqs, edges)
    rblob, roff = pack_reads(reads)
    n = len(reads)
    out = np.zeros((n, 6), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
    lib.pgref_align_batch(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32), _p(et, C.c_int32),
                          n, rblob, _p(roff, C.c_int32), None if rv is None else _p(rv, C.c_uint8),
                          flags & 0xFFFFFFFF, threads, _p(out, C.c_int32), ob, cg, CIGAR_STRIDE)
    res = []
    raw = ob.raw
    craw = cg.raw  # ONE copy of the buffer (.raw copies: taking it per read made this loop quadratic in the batch size)
    for i in range(n):
        c = craw[i * CIGAR_STRIDE:i * CIGAR_STRIDE + int(out[i][5])].decode()
        res.append(_result_dict(out[i], raw[roff[i]:roff[i + 1]].decode("latin-1"), c))
    return res


class RefBatch:
    """A graph and a read batch packed once, for TIMING the compiled reference: run(lo, hi, threads) executes
    pgref_align_batch (GraphAligner::alignRead over reads [lo, hi), `threads` GraphAligners over contiguous chunks like
    grm::alignReads, Align.cpp:107-153) and returns the seconds spent inside the library call -- no Python per read."""

    def __init__(self, node_seqs, edges, reads):
        self.lib = ref_lib()
        self.blob, self.off, self.ef, self.et = pack_graph(node_seqs, edges)
        self.n_nodes, self.n_edges = len(node_seqs), len(edges)
        self.rblob, self.roff = pack_reads(reads)
        self.n = len(reads)
        self.out = np.zeros((self.n, 6), dtype=np.int32)

    def run(self, lo, hi, threads=1, flags=AF_ALL):
        import time
        off = np.ascontiguousarray(self.roff[lo:hi + 1])
        t0 = time.perf_counter()
        self.lib.pgref_align_batch(self.n_nodes, self.blob, _p(self.off, C.c_int32), self.n_edges, _p(self.ef, C.c_int32),
                                   _p(self.et, C.c_int32), hi - lo, self.rblob, _p(off, C.c_int32), None,
                                   flags & 0xFFFFFFFF, threads, _p(self.out[lo:hi], C.c_int32), None, None, 0)
        return time.perf_counter() - t0


def pack_sites(sites):
    """sites = [(nodes, edges, reads), ...] -> the flat arrays pgref_align_sites takes (kept by the caller)."""
    node_ptr, edge_ptr, read_ptr = [0], [0], [0]
    seqs, ef, et, reads = [], [], [], []
    for nodes, edges, rds in sites:
        seqs += list(nodes)
        ef += [e[0] for e in edges]
        et += [e[1] for e in edges]
        reads += list(rds)
        node_ptr.append(len(seqs))
        edge_ptr.append(len(ef))
        read_ptr.append(len(reads))
    blob = "".join(seqs).encode("latin-1")
    off = np.zeros(len(seqs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    rblob, roff = pack_reads(reads)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return dict(n_sites=len(sites), node_ptr=i32(node_ptr), blob=blob, off=off, edge_ptr=i32(edge_ptr), ef=i32(ef),
                et=i32(et), read_ptr=i32(read_ptr), rblob=rblob, roff=roff, n_reads=len(reads))


def ref_align_sites_packed(pk, threads=1, flags=AF_ALL, cigar_stride=192, want_bases=False):
    """The reference GraphAligner::alignRead over many sites: threads pull whole sites (Workflow.cpp:108-146).
    Returns (out6 int32[n,6] = pos, score, unique, mapq, graph_reverse, cigar length; cigar bytes[n, stride])."""
    lib = ref_lib()
    n = pk["n_reads"]
    out = np.zeros((n, 6), dtype=np.int32)
    cg = np.zeros((max(1, n), cigar_stride), dtype=np.uint8)
    ob = C.create_string_buffer(max(1, len(pk["rblob"]))) if want_bases else None
    bad = lib.pgref_align_sites(pk["n_sites"], _p(pk["node_ptr"], C.c_int32), pk["blob"], _p(pk["off"], C.c_int32),
                                _p(pk["edge_ptr"], C.c_int32), _p(pk["ef"], C.c_int32), _p(pk["et"], C.c_int32),
                                _p(pk["read_ptr"], C.c_int32), pk["rblob"], _p(pk["roff"], C.c_int32), None,
                                flags & 0xFFFFFFFF, threads, _p(out, C.c_int32), ob,
                                cg.ctypes.data_as(C.c_char_p), cigar_stride)
    if bad:
        raise RuntimeError("reference threw on %d sites" % bad)
    return out, cg


def ref_align_sites(sites, threads=1, flags=AF_ALL):
    """-> list of result dicts (pos, score, unique, mapq, graph_reverse, cigar) over all reads of all sites, in order."""
    pk = pack_sites(sites)
    out, cg = ref_align_sites_packed(pk, threads, flags, cigar_stride=1024)
    res = []
    for i in range(pk["n_reads"]):
        c = cg[i].tobytes().split(b"\0", 1)[0].decode()
        assert len(c) == out[i, 5], "cigar truncated"
        res.append(dict(pos=int(out[i, 0]), score=int(out[i, 1]), unique=bool(out[i, 2]), mapq=int(out[i, 3]),
                        graph_reverse=bool(out[i, 4]), cigar=c))
    return res


class RefGssw:
    """Raw gssw fill + traceback with matrices (reference external/gssw/gssw.c)."""

    def __init__(self, node_seqs, edges):
        self.lib = ref_lib()
        self.node_seqs = list(node_seqs)
        blob, off, ef, et = pack_graph(node_seqs, edges)
        self.h = self.lib.pgref_gssw_create(len(node_seqs), blob, _p(off, C.c_int32), len(edges),
                                            _p(ef, C.c_int32), _p(et, C.c_int32))

    def fill_trace(self, read, want_mats=True, wide=False):
        """wide: 16-bit matrices, valid in gssw's byte and word mode alike (stats[:, 3] = is_byte)."""
        L = len(read)
        n = len(self.node_seqs)
        stats = np.zeros((n, 4), dtype=np.int32)
        tot = sum(len(s) for s in self.node_seqs) * L * 3
        mats = np.zeros(max(1, tot), dtype=np.uint16 if wide else np.uint8) if want_mats else None
        res = np.zeros(3, dtype=np.int32)
        cg = C.create_string_buffer(8192)
        if wide:
            self.lib.pgref_gssw_fill_trace16(self.h, read.encode("latin-1"), _p(stats, C.c_int32),
                                             _p(mats, C.c_uint16) if want_mats else None, _p(res, C.c_int32), cg, 8192)
        else:
            self.lib.pgref_gssw_fill_trace(self.h, read.encode("latin-1"), _p(stats, C.c_int32),
                                           _p(mats, C.c_uint8) if want_mats else None, _p(res, C.c_int32), cg, 8192)
        out = dict(stats=stats, max_node=int(res[0]), pos=int(res[1]), score=int(res[2]), cigar=cg.value.decode())
        if want_mats:
            ms, o = [], 0
            for s in self.node_seqs:
                sz = len(s) * L
                ms.append(tuple(mats[o + k * sz:o + (k + 1) * sz].reshape(len(s), L) for k in range(3)))
                o += 3 * sz
            out["mats"] = ms
        return out

    def close(self):
        if self.h:
            self.lib.pgref_gssw_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------- plain-C restatement (oracle/pg_oracle.c)
_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(ORACLE_SO)
        lib.pgo_graph_create.restype = C.c_void_p
        lib.pgo_graph_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.pgo_graph_destroy.argtypes = [C.c_void_p]
        lib.pgo_align_read.restype = C.c_int
        lib.pgo_align_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_uint,
                                       C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int]
        lib.pgo_align_batch.restype = C.c_int
        lib.pgo_align_batch.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                        C.c_uint, C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int]
        lib.pgo_fill_trace.restype = C.c_int
        lib.pgo_fill_trace.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.c_char_p, C.c_int]
        lib.pgo_fill_trace16.restype = C.c_int
        lib.pgo_fill_trace16.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_uint16), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.c_char_p, C.c_int]
        lib.pgo_set_fill_variant.argtypes = [C.c_int]
        _orc = lib
    return _orc


def oracle_bad_align(cigar, frac=0.8):
    """-> (filtered, clipped) per the oracle's restatement of readfilters::BadAlign"""
    lib = oracle_lib()
    lib.pgo_bad_align.restype = C.c_int
    lib.pgo_bad_align.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_int)]
    cl = C.c_int(0)
    bad = lib.pgo_bad_align(cigar.encode(), float(frac), C.byref(cl))
    return bool(bad), cl.value


def set_fill_variant(v):
    """0 = faithful restatement; 1 = textbook E; 2 = kernel recurrence (see pg_oracle.c)."""
    oracle_lib().pgo_set_fill_variant(int(v))


class OracleGraph:
    def __init__(self, node_seqs, edges):
        self.lib = oracle_lib()
        self.node_seqs = list(node_seqs)
        blob, off, ef, et = pack_graph(node_seqs, edges)
        self.h = self.lib.pgo_graph_create(len(node_seqs), blob, _p(off, C.c_int32), len(edges),
                                           _p(ef, C.c_int32), _p(et, C.c_int32))
        if not self.h:
            raise ValueError("oracle: bad graph")

    def align_batch(self, reads, is_rev=None, flags=AF_ALL):
        rblob, roff = pack_reads(reads)
        n = len(reads)
        out = np.zeros((n, 6), dtype=np.int32)
        ob = C.create_string_buffer(max(1, len(rblob)))
        cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
        rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
        rc = self.lib.pgo_align_batch(self.h, n, rblob, _p(roff, C.c_int32),
                                      None if rv is None else _p(rv, C.c_uint8), flags & 0xFFFFFFFF,
                                      _p(out, C.c_int32), ob, cg, CIGAR_STRIDE)
        if rc != 0:
            raise RuntimeError("oracle: pgo_align_batch rc=%d" % rc)
        res, raw = [], ob.raw
        craw = cg.raw  # one copy (see ref_align_batch)
        for i in range(n):
            c = craw[i * CIGAR_STRIDE:(i + 1) * CIGAR_STRIDE].split(b"\0", 1)[0].decode()
            res.append(_result_dict(out[i], raw[roff[i]:roff[i + 1]].decode("latin-1"), c))
        return res

    def fill_trace(self, read, reversed_graph=False, want_mats=True, wide=False):
        L = len(read)
        seqs = self.node_seqs[::-1] if reversed_graph else self.node_seqs
        n = len(seqs)
        stats = np.zeros((n, 4), dtype=np.int32)
        tot = sum(len(s) for s in seqs) * L * 3
        mats = np.zeros(max(1, tot), dtype=np.uint16 if wide else np.uint8) if want_mats else None
        res = np.zeros(3, dtype=np.int32)
        multi = np.zeros(1, dtype=np.int32)
        cg = C.create_string_buffer(8192)
        fn, ct = (self.lib.pgo_fill_trace16, C.c_uint16) if wide else (self.lib.pgo_fill_trace, C.c_uint8)
        rc = fn(self.h, 1 if reversed_graph else 0, read.encode("latin-1"), L,
                _p(stats, C.c_int32), _p(mats, ct) if want_mats else None,
                _p(res, C.c_int32), _p(multi, C.c_int32), cg, 8192)
        if rc < 0:
            raise RuntimeError("oracle: pgo_fill_trace rc=%d" % rc)
        out = dict(stats=stats, max_node=int(res[0]), pos=int(res[1]), score=int(res[2]), multi=bool(multi[0]),
                   cigar=cg.value.decode())
        if want_mats:
            ms, o = [], 0
            for s in seqs:
                sz = len(s) * L
                ms.append(tuple(mats[o + k * sz:o + (k + 1) * sz].reshape(len(s), L) for k in range(3)))
                o += 3 * sz
            out["mats"] = ms
        return out

    def close(self):
        if self.h:
            self.lib.pgo_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def ref_filter_batch(node_seqs, edges, read_lens, graph_pos, unique, cigars, bad_align_frac=0.8):
    """Reference decodeGraphAlignment + readfilters::NonUniq / BadAlign on given alignments.
    Returns int32 array [n][4] = {decode_ok, query_clipped, nonuniq_filtered, badalign_filtered}."""
    lib = ref_lib()
    lib.pgref_filter_batch.restype = C.c_int
    lib.pgref_filter_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_uint8), C.c_char_p, C.c_int, C.c_double, C.POINTER(C.c_int32)]
    blob, off, ef, et = pack_graph(node_seqs, edges)
    n = len(read_lens)
    rl = np.ascontiguousarray(read_lens, dtype=np.int32)
    gp = np.ascontiguousarray(graph_pos, dtype=np.int32)
    un = np.ascontiguousarray(unique, dtype=np.uint8)
    buf = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    for i, c in enumerate(cigars):
        b = c.encode()
        buf[i * CIGAR_STRIDE:i * CIGAR_STRIDE + len(b) + 1] = b + b"\0"
    out = np.zeros((n, 4), dtype=np.int32)
    lib.pgref_filter_batch(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32), _p(et, C.c_int32),
                           n, _p(rl, C.c_int32), _p(gp, C.c_int32), _p(un, C.c_uint8), buf, CIGAR_STRIDE,
                           float(bad_align_frac), _p(out, C.c_int32))
    return out


def _cigar_buffer(cigars, stride=None):
    stride = stride or max([CIGAR_STRIDE] + [len(c) + 1 for c in cigars])
    buf = C.create_string_buffer(max(1, len(cigars) * stride))
    for i, c in enumerate(cigars):
        b = c.encode()
        buf[i * stride:i * stride + len(b) + 1] = b + b"\0"
    return buf, stride


def ref_count_site(node_seqs, edges, edge_labels, read_lens, graph_pos, cigars, is_graph_reverse=None, fragment=None,
                   use_filters=True, detailed=True):
    """Disambiguation + counting of ONE site's MAPPED reads through the reference (oracle/ref_counts.cpp: the
    unmodified ReadCounting.cpp / Fragment.cpp / graph-tools, plus the restated filter lambdas).
    edge_labels: one uint64 bit mask per edge (bit k = label "L<k>") or None.  Returns the parsed JSON document
    (node i is "n<i>", edges "n<i>_n<j>"), or raises RuntimeError with the reference's exception text."""
    import json
    lib = ref_lib()
    lib.pgref_count_site.restype = C.c_int
    lib.pgref_count_site.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.c_char_p, C.c_int, C.POINTER(C.c_uint8),
                                     C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_char_p, C.c_int]
    blob, off, ef, et = pack_graph(node_seqs, edges)
    n = len(read_lens)
    rl = np.ascontiguousarray(read_lens, dtype=np.int32)
    gp = np.ascontiguousarray(graph_pos, dtype=np.int32)
    lab = np.ascontiguousarray(edge_labels if edge_labels is not None else np.zeros(len(edges)), dtype=np.uint64)
    rv = np.ascontiguousarray(is_graph_reverse if is_graph_reverse is not None else np.zeros(n), dtype=np.uint8)
    fr = np.ascontiguousarray(fragment if fragment is not None else np.arange(n), dtype=np.int32)
    buf, stride = _cigar_buffer(cigars)
    cap = 1 << 16
    while True:
        out = C.create_string_buffer(cap)
        rc = lib.pgref_count_site(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32),
                                  _p(et, C.c_int32), _p(lab, C.c_uint64), n, _p(rl, C.c_int32), _p(gp, C.c_int32),
                                  buf, stride, _p(rv, C.c_uint8), _p(fr, C.c_int32), int(use_filters), int(detailed),
                                  out, cap)
        if rc == -1:
            cap *= 4
            continue
        if rc == -2:
            raise RuntimeError(out.value.decode(errors="replace"))
        return json.loads(out.value.decode())


V_MAPPED, V_NONUNIQ, V_BAD_ALIGN, V_INVALID = 0, 1, 2, 3
SUP_NODE_MASK, SUP_NODE, SUP_EDGE = 0xFFFF, 0x40000000, 0x80000000
SUPPORT_DTYPE = np.dtype([("sequences", np.uint64), ("path_off", np.uint32), ("path_len", np.uint16),
                          ("verdict", np.uint8), ("graph_reverse", np.uint8)])


def unpack_families(words, n_nodes, n_edges):
    """family words -> {mask: int array [(1+n_nodes+n_edges), 4]}"""
    stride = 2 + 4 * (1 + n_nodes + n_edges)
    out = {}
    for q in range(len(words) // stride):
        h = words[q * stride:(q + 1) * stride]
        out[int(h[0]) | (int(h[1]) << 32)] = np.array(h[2:], dtype=np.int64).reshape(-1, 4)
    return out


def oracle_count_site(node_lens, edges, edge_labels, read_lens, graph_pos, unique, cigars, is_graph_reverse=None,
                      fragment=None, remove_nonuniq=True, bad_align_frac=0.8, use_filters=True):
    """oracle/pg_oracle_counts.c::pgo_count_site.  Returns dict(support, path_words, node_counts [n,4],
    edge_counts [m,4], families {mask: [(1+n+m),4]})."""
    lib = oracle_lib()
    lib.pgo_count_site.restype = C.c_int
    i32p, u8p, u32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    lib.pgo_count_site.argtypes = [C.c_int, i32p, C.c_int, i32p, i32p, u64p, C.c_int, i32p, i32p, u8p, C.c_char_p,
                                   C.c_int, u8p, i32p, C.c_int, C.c_double, C.c_int, C.c_void_p, u32p, C.c_int,
                                   i32p, C.c_void_p, C.c_void_p, u32p, C.c_int, i32p]
    n, nn, ne = len(read_lens), len(node_lens), len(edges)
    nl = np.ascontiguousarray(node_lens, dtype=np.int32)
    ef = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
    et = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
    lab = np.ascontiguousarray(edge_labels if edge_labels is not None else np.zeros(ne), dtype=np.uint64)
    rl = np.ascontiguousarray(read_lens, dtype=np.int32)
    gp = np.ascontiguousarray(graph_pos, dtype=np.int32)
    un = np.ascontiguousarray(unique, dtype=np.uint8)
    rv = np.ascontiguousarray(is_graph_reverse if is_graph_reverse is not None else np.zeros(n), dtype=np.uint8)
    fr = np.ascontiguousarray(fragment if fragment is not None else np.arange(n), dtype=np.int32)
    buf, stride = _cigar_buffer(cigars)
    sup = np.zeros(n, dtype=SUPPORT_DTYPE)
    path_cap = max(1, sum(c.count("[") for c in cigars))
    pw = np.zeros(path_cap, dtype=np.uint32)
    nc = np.zeros((max(nn, 1), 4), dtype=np.uint32)
    ec = np.zeros((max(ne, 1), 4), dtype=np.uint32)
    fam_cap = (2 + 4 * (1 + nn + ne)) * 256
    fw = np.zeros(fam_cap, dtype=np.uint32)
    used = np.zeros(2, dtype=np.int32)
    rc = lib.pgo_count_site(nn, _p(nl, C.c_int32), ne, _p(ef, C.c_int32), _p(et, C.c_int32), _p(lab, C.c_uint64), n,
                            _p(rl, C.c_int32), _p(gp, C.c_int32), _p(un, C.c_uint8), buf, stride, _p(rv, C.c_uint8),
                            _p(fr, C.c_int32), int(remove_nonuniq), float(bad_align_frac), int(use_filters),
                            sup.ctypes.data, _p(pw, C.c_uint32), path_cap, _p(used[0:], C.c_int32), nc.ctypes.data,
                            ec.ctypes.data, _p(fw, C.c_uint32), fam_cap, _p(used[1:], C.c_int32))
    if rc != 0:
        raise RuntimeError("pgo_count_site rc=%d" % rc)
    return dict(support=sup, path_words=pw[:used[0]], node_counts=nc[:nn].astype(np.int64),
                edge_counts=ec[:ne].astype(np.int64), families=unpack_families(fw[:used[1]], nn, ne))


def counts_from_ref_doc(doc, n_nodes, edges):
    """Convert ref_count_site()'s JSON (names n<i>, L<k>) into the oracle's array layout."""
    def four(d, key):
        return [d.get(key, 0), d.get(key + ":READS", 0), d.get(key + ":FWD", 0), d.get(key + ":REV", 0)]
    nc = np.array([four(doc["read_counts_by_node"], "n%d" % i) for i in range(n_nodes)], dtype=np.int64).reshape(-1, 4)
    ec = np.array([four(doc["read_counts_by_edge"], "n%d_n%d" % e) for e in edges], dtype=np.int64).reshape(-1, 4)
    fams = {}
    for key, d in doc["read_counts_by_sequence"].items():
        mask = sum(1 << int(x[1:]) for x in key.split(","))
        rows = [four(d, "total")] + [four(d, "n%d" % i) for i in range(n_nodes)] + [four(d, "n%d_n%d" % e) for e in edges]
        fams[mask] = np.array(rows, dtype=np.int64)
    return nc, ec, fams


def support_sets(sup, path_words, i, edges=None):
    """(nodes set, edges set of (a,b), sequences mask) of read i from the oracle/GPU arrays."""
    w = path_words[int(sup["path_off"][i]):int(sup["path_off"][i]) + int(sup["path_len"][i])]
    nodes = {int(x & SUP_NODE_MASK) for x in w if x & SUP_NODE}
    es = {(int(w[k - 1] & SUP_NODE_MASK), int(w[k] & SUP_NODE_MASK)) for k in range(1, len(w)) if w[k] & SUP_EDGE}
    return nodes, es, int(sup["sequences"][i])


# ---------------------------------------------------------------- grm::PathAligner (exact-match stage)
def _path_result(o, bases, cigar):
    return dict(mapped=bool(o[0]), pos=int(o[1]), score=int(o[2]), unique=bool(o[3]), mapq=int(o[4]),
                graph_reverse=bool(o[5]), bases=bases, cigar=cigar)


def ref_path_align_batch(node_seqs, edges, reads, kmer_len=32, is_rev=None):
    """The UNMODIFIED reference PathAligner (oracle/ref_path.cpp).  -> (list of dicts, (attempted, anchored, mapped))"""
    lib = ref_lib()
    lib.pgref_path_align_batch.restype = C.c_int
    lib.pgref_path_align_batch.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int,
                                           C.POINTER(C.c_int32)]
    blob, off, ef, et = pack_graph(node_seqs, edges)
    rblob, roff = pack_reads(reads)
    n = len(reads)
    out = np.zeros((n, 8), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    cnt = np.zeros(3, dtype=np.int32)
    rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
    rc = lib.pgref_path_align_batch(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32),
                                    _p(et, C.c_int32), int(kmer_len), n, rblob, _p(roff, C.c_int32),
                                    None if rv is None else _p(rv, C.c_uint8), _p(out, C.c_int32), ob, cg, CIGAR_STRIDE,
                                    _p(cnt, C.c_int32))
    if rc != 0:
        raise RuntimeError("reference PathAligner threw")
    raw = ob.raw
    craw = cg.raw
    res = []
    for i in range(n):
        c = craw[i * CIGAR_STRIDE:(i + 1) * CIGAR_STRIDE].split(b"\0", 1)[0].decode()
        res.append(_path_result(out[i], raw[roff[i]:roff[i + 1]].decode("latin-1"), c))
    return res, tuple(int(x) for x in cnt)


class OraclePathIndex:
    """oracle/pg_oracle_path.c: the C restatement of KmerIndex + PathAligner::alignRead."""

    def __init__(self, node_seqs, edges, kmer_len=32):
        self.lib = oracle_lib()
        self.lib.pgo_path_index_create.restype = C.c_void_p
        self.lib.pgo_path_index_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int,
                                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
        self.lib.pgo_path_index_destroy.argtypes = [C.c_void_p]
        self.lib.pgo_path_align_batch.restype = C.c_int
        self.lib.pgo_path_align_batch.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int,
                                                  C.POINTER(C.c_int32)]
        blob, off, ef, et = pack_graph(node_seqs, edges)
        self.h = self.lib.pgo_path_index_create(len(node_seqs), blob, _p(off, C.c_int32), len(edges),
                                                _p(ef, C.c_int32), _p(et, C.c_int32), int(kmer_len))
        if not self.h:
            raise ValueError("oracle: bad graph / k-mer length")

    def align_batch(self, reads):
        rblob, roff = pack_reads(reads)
        n = len(reads)
        out = np.zeros((n, 8), dtype=np.int32)
        ob = C.create_string_buffer(max(1, len(rblob)))
        cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
        cnt = np.zeros(3, dtype=np.int32)
        self.lib.pgo_path_align_batch(self.h, n, rblob, _p(roff, C.c_int32), _p(out, C.c_int32), ob, cg, CIGAR_STRIDE,
                                      _p(cnt, C.c_int32))
        raw = ob.raw
        craw = cg.raw
        res = []
        for i in range(n):
            c = craw[i * CIGAR_STRIDE:(i + 1) * CIGAR_STRIDE].split(b"\0", 1)[0].decode()
            res.append(_path_result(out[i], raw[roff[i]:roff[i + 1]].decode("latin-1"), c))
        return res, tuple(int(x) for x in cnt)

    def close(self):
        if self.h:
            self.lib.pgo_path_index_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------- grm::KmerAligner (second stage of the cascade)
KMER_STATUS = ("unmapped", "mapped", "bad_align")


def pack_paths(paths):
    """[[node ids], ...] -> (path_ptr int32[n+1], path_nodes int32[])"""
    ptr = np.zeros(len(paths) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(p) for p in paths])
    flat = np.ascontiguousarray([v for p in paths for v in p], dtype=np.int32)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.int32)
    return ptr, flat


def _kmer_result(o, bases, cigar):
    st = KMER_STATUS[int(o[0])]
    if st == "unmapped":  # nothing was written to the read
        return dict(status=st)
    return dict(status=st, pos=int(o[1]), score=int(o[2]), unique=bool(o[3]), mapq=int(o[4]), graph_reverse=bool(o[5]),
                bases=bases, cigar=cigar)


def _kmer_collect(out, ob, cg, roff, n):
    raw, craw = ob.raw, cg.raw
    res = []
    for i in range(n):
        c = craw[i * CIGAR_STRIDE:i * CIGAR_STRIDE + int(out[i][6])].decode()
        res.append(_kmer_result(out[i], raw[roff[i]:roff[i + 1]].decode("latin-1"), c))
    return res


def ref_kmer_align_batch(node_seqs, edges, paths, reads, kmer_len=16, is_rev=None):
    """The reference's KmerAligner<kmer_len> (kmer_len 16 or 10) over a batch -> (result dicts, (attempted, mapped))."""
    lib = ref_lib()
    blob, off, ef, et = pack_graph(node_seqs, edges)
    pptr, pnodes = pack_paths(paths)
    rblob, roff = pack_reads(reads)
    n = len(reads)
    out = np.zeros((max(n, 1), 8), dtype=np.int32)
    ob = C.create_string_buffer(max(1, len(rblob)))
    cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
    cnt = np.zeros(2, dtype=np.int32)
    rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
    rc = lib.pgref_kmer_align_batch(len(node_seqs), blob, _p(off, C.c_int32), len(edges), _p(ef, C.c_int32), _p(et, C.c_int32),
                                    len(paths), _p(pptr, C.c_int32), _p(pnodes, C.c_int32), int(kmer_len), n, rblob,
                                    _p(roff, C.c_int32), None if rv is None else _p(rv, C.c_uint8), _p(out, C.c_int32), ob, cg,
                                    CIGAR_STRIDE, _p(cnt, C.c_int32))
    if rc != 0:
        raise RuntimeError("reference KmerAligner: rc=%d" % rc)
    return _kmer_collect(out, ob, cg, roff, n), tuple(int(x) for x in cnt)


class OracleKmerIndex:
    """oracle/pg_oracle_kmer.c: the restatement of grm::KmerAligner<K> (any 2 <= K <= 16)."""

    def __init__(self, node_seqs, edges, paths, kmer_len=16):
        self.lib = oracle_lib()
        self.lib.pgo_kmer_index_create.restype = C.c_void_p
        self.lib.pgo_kmer_index_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                                   C.POINTER(C.c_int32), C.c_int]
        self.lib.pgo_kmer_index_destroy.argtypes = [C.c_void_p]
        self.lib.pgo_kmer_align_batch.restype = C.c_int
        self.lib.pgo_kmer_align_batch.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                                  C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int]
        blob, off, _, _ = pack_graph(node_seqs, edges)
        pptr, pnodes = pack_paths(paths)
        self.h = self.lib.pgo_kmer_index_create(len(node_seqs), blob, _p(off, C.c_int32), len(paths), _p(pptr, C.c_int32),
                                                _p(pnodes, C.c_int32), int(kmer_len))
        if not self.h:
            raise RuntimeError("oracle: pgo_kmer_index_create failed")

    def align_batch(self, reads, is_rev=None):
        rblob, roff = pack_reads(reads)
        n = len(reads)
        out = np.zeros((max(n, 1), 8), dtype=np.int32)
        ob = C.create_string_buffer(max(1, len(rblob)))
        cg = C.create_string_buffer(max(1, n * CIGAR_STRIDE))
        rv = None if is_rev is None else np.asarray(is_rev, dtype=np.uint8)
        rc = self.lib.pgo_kmer_align_batch(self.h, n, rblob, _p(roff, C.c_int32), None if rv is None else _p(rv, C.c_uint8),
                                           _p(out, C.c_int32), ob, cg, CIGAR_STRIDE)
        if rc != 0:
            raise RuntimeError("oracle: pgo_kmer_align_batch rc=%d" % rc)
        return _kmer_collect(out, ob, cg, roff, n)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pgo_kmer_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
