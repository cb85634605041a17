"""ctypes binding of the C-ABI (include/pg_align.h) exported by paragraph_b200/libpgalign.so.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU
fallback anywhere in this package: importing works without a GPU (so that symbols can be checked), but
creating a context raises ``PgError`` unless a B200-class device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpgalign.so")

PG_OK = 0
AF_CIGAR, AF_BOTH_STRANDS, AF_REVERSE_GRAPH, AF_ALL = 1, 2, 4, 0xFFFFFFFF
MAX_READ_LEN = 1024
OPS = "MXNIDS??"

RECORD_DTYPE = np.dtype([("graph_pos", "<i4"), ("score", "<i2"), ("query_clipped", "<u2"), ("unique", "u1"),
                         ("chose_reverse", "u1"), ("status", "u1"), ("mapped_by", "u1"), ("cigar_off", "<u4"),
                         ("cigar_len", "<u4")])

# every symbol include/pg_align.h declares
SYMBOLS = ["pg_create", "pg_destroy", "pg_last_error", "pg_set_stream", "pg_set_scratch_limit", "pg_add_graph",
           "pg_add_graphs", "pg_clear_graphs", "pg_align_batch", "pg_batch_upload", "pg_batch_run", "pg_batch_download",
           "pg_format_cigar", "pg_stats", "pg_version", "pg_host_alloc", "pg_host_free", "pg_set_edge_labels",
           "pg_batch_import", "pg_batch_count", "pg_count_stats", "pg_set_stages", "pg_path_stats", "pg_set_paths",
           "pg_set_kmer_stage", "pg_kmer_stats"]

# counting stage (include/pg_align.h, "Counting stage")
V_MAPPED, V_NONUNIQ, V_BAD_ALIGN, V_INVALID = 0, 1, 2, 3
SUP_NODE_MASK, SUP_NODE, SUP_EDGE = 0xFFFF, 0x40000000, 0x80000000
SUPPORT_DTYPE = np.dtype([("sequences", "<u8"), ("path_off", "<u4"), ("path_len", "<u2"), ("verdict", "u1"),
                          ("graph_reverse", "u1")])


# pg_record.mapped_by -> (stage of the cascade that mapped the read, reverse complements applied by the stages before it)
STAGES = {0: ("gssw", 0), 1: ("path", 0), 2: ("gssw", 1), 3: ("kmer", 0), 4: ("kmer", 1), 5: ("gssw", 2)}


class CountParams(C.Structure):
    _fields_ = [("remove_nonuniq", C.c_int32), ("use_support_filters", C.c_int32), ("bad_align_frac", C.c_double),
                ("family_slots", C.c_int32), ("reserved", C.c_int32)]


class PgError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libpgalign.so and declare prototypes.  Fails loudly when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PG_LIB", LIB_PATH)  # A/B builds of the same ABI (tools/ab_variants.py); default = in-tree library
    if not os.path.exists(path):
        raise PgError("paragraph_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, i32p, u32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    lib.pg_create.restype = C.c_int
    lib.pg_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.pg_destroy.restype = None
    lib.pg_destroy.argtypes = [vp]
    lib.pg_last_error.restype = C.c_char_p
    lib.pg_last_error.argtypes = [vp]
    lib.pg_set_stream.restype = C.c_int
    lib.pg_set_stream.argtypes = [vp, vp]
    lib.pg_set_scratch_limit.restype = C.c_int
    lib.pg_set_scratch_limit.argtypes = [vp, C.c_uint64]
    lib.pg_add_graph.restype = C.c_int
    lib.pg_add_graph.argtypes = [vp, C.c_int32, C.c_char_p, i32p, C.c_int32, i32p, i32p, i32p]
    lib.pg_add_graphs.restype = C.c_int
    lib.pg_add_graphs.argtypes = [vp, C.c_int32, i32p, C.c_char_p, i32p, i32p, i32p, i32p, i32p]
    lib.pg_clear_graphs.restype = C.c_int
    lib.pg_clear_graphs.argtypes = [vp]
    lib.pg_set_paths.restype = C.c_int
    lib.pg_set_paths.argtypes = [vp, C.c_int32, C.c_int32, i32p, i32p]
    lib.pg_set_kmer_stage.restype = C.c_int
    lib.pg_set_kmer_stage.argtypes = [vp, C.c_int32]
    lib.pg_kmer_stats.restype = C.c_int
    lib.pg_kmer_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    lib.pg_align_batch.restype = C.c_int
    lib.pg_align_batch.argtypes = [vp, C.c_int32, vp, i32p, i32p, C.c_uint32, vp, u32p, C.c_uint64,
                                   C.POINTER(C.c_uint64)]
    lib.pg_batch_upload.restype = C.c_int
    lib.pg_batch_upload.argtypes = [vp, C.c_int32, vp, i32p, i32p]
    lib.pg_batch_run.restype = C.c_int
    lib.pg_batch_run.argtypes = [vp, C.c_uint32]
    lib.pg_batch_download.restype = C.c_int
    lib.pg_batch_download.argtypes = [vp, vp, u32p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.pg_format_cigar.restype = C.c_int
    lib.pg_format_cigar.argtypes = [vp, u32p, C.c_char_p, C.c_int]
    lib.pg_stats.restype = C.c_int
    lib.pg_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.pg_version.restype = C.c_char_p
    lib.pg_version.argtypes = []
    lib.pg_host_alloc.restype = C.c_int
    lib.pg_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    lib.pg_host_free.restype = None
    lib.pg_host_free.argtypes = [vp]
    u64p = C.POINTER(C.c_uint64)
    lib.pg_set_edge_labels.restype = C.c_int
    lib.pg_set_edge_labels.argtypes = [vp, C.c_int32, u64p]
    lib.pg_batch_import.restype = C.c_int
    lib.pg_batch_import.argtypes = [vp, C.c_int32, i32p, i32p, vp, u32p, C.c_uint64]
    lib.pg_count_stats.restype = C.c_int
    lib.pg_count_stats.argtypes = [vp, u64p, C.POINTER(C.c_float)]
    lib.pg_set_stages.restype = C.c_int
    lib.pg_set_stages.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.pg_path_stats.restype = C.c_int
    lib.pg_path_stats.argtypes = [vp, u64p, C.POINTER(C.c_float)]
    lib.pg_batch_count.restype = C.c_int
    lib.pg_batch_count.argtypes = [vp, i32p, C.POINTER(C.c_uint8), C.POINTER(CountParams), vp, u32p, C.c_uint64, u64p,
                                   vp, C.c_uint64, vp, C.c_uint64, u32p, C.c_uint64, u64p]
    _lib = lib
    return lib


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def format_cigar(rec, ops):
    """'<node>[<len><op>...]...' exactly as GraphAlignerImpl::extractCigar (GraphAligner.cpp:88-108)."""
    out, cur = [], -1
    o0, n = int(rec["cigar_off"]), int(rec["cigar_len"])
    for w in ops[o0:o0 + n]:
        w = int(w)
        node = w >> 16
        if node != cur:
            if cur >= 0:
                out.append("]")
            out.append("%d[" % node)
            cur = node
        if (w & 7) == 7:
            continue
        out.append("%d%s" % ((w >> 3) & 0x1FFF, OPS[w & 7]))
    if cur >= 0:
        out.append("]")
    return "".join(out)


def parse_cigar(cigar):
    """'<node>[<len><op>...]...' -> op words (node << 16 | len << 3 | op), the inverse of format_cigar; an empty
    node group "id[]" becomes the OP_NONE marker."""
    import re
    words = []
    for node, body in re.findall(r"(\d+)\[([^\]]*)\]", cigar):
        ops = re.findall(r"(\d+)([MXNIDS])", body)
        if not ops:
            words.append((int(node) << 16) | 7)
        for ln, op in ops:
            words.append((int(node) << 16) | (int(ln) << 3) | OPS.index(op))
    return words


class PinnedArray:
    """numpy view of page-locked host memory from pg_host_alloc (copied to / from the device without staging)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        if lib.pg_host_alloc(max(n, 16), C.byref(p)) != PG_OK:
            raise PgError("pg_host_alloc(%d) failed" % n)
        self._lib, self._p = lib, p
        buf = (C.c_uint8 * max(n, 16)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            self._lib.pg_host_free(self._p)
        except Exception:
            pass


class Context:
    """One engine context (one CUDA stream).  Mirrors a grm::GraphAligner that can hold many graphs."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.pg_create(int(device), C.byref(h))
        if rc != PG_OK:
            raise PgError("pg_create(device=%d) failed with %d: no usable sm_100 CUDA device; "
                          "paragraph_b200 has no CPU fallback" % (device, rc))
        self.h = h
        self._keep = None
        self._shape = []   # (n_nodes, n_edges) of every registered site, for sizing the count tables
        self._n = 0
        self._rec = None   # output buffers are reused between calls of the same size (no per-call allocation)
        self._ops = None
        if stream is not None:
            self.set_stream(stream)

    def _check(self, rc):
        if rc != PG_OK:
            raise PgError("paragraph_b200 error %d: %s" % (rc, self.lib.pg_last_error(self.h).decode()))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.pg_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else None)))

    def set_scratch_limit(self, nbytes):
        self._check(self.lib.pg_set_scratch_limit(self.h, int(nbytes)))

    def set_stages(self, path_kmer_len=0, graph_matching=True, nonuniq_second_chance=False):
        """grm::CompositeAligner's cascade: exact-match stage (PathAligner, k-mer length; 0 = off) in front of the DP;
        nonuniq_second_chance: a non-unique exact match goes on to the DP (what the NonUniq read filter causes)."""
        self._check(self.lib.pg_set_stages(self.h, int(path_kmer_len), 1 if graph_matching else 0,
                                           1 if nonuniq_second_chance else 0))
        self._path_k = int(path_kmer_len)

    def set_paths(self, site, paths):
        """The paths of the site's graph JSON ([[node ids], ...]): what the k-mer stage aligns to."""
        ptr = np.zeros(len(paths) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(p) for p in paths])
        flat = np.ascontiguousarray([v for p in paths for v in p] or [0], dtype=np.int32)
        self._check(self.lib.pg_set_paths(self.h, int(site), len(paths), _i32(ptr), _i32(flat)))

    def set_kmer_stage(self, kmer_len=16):
        """grm::KmerAligner<kmer_len> between the exact-match stage and gssw (0 = off)."""
        self._check(self.lib.pg_set_kmer_stage(self.h, int(kmer_len)))
        self._kmer_k = int(kmer_len)

    def kmer_stats(self):
        cnt = (C.c_uint64 * 2)()
        ms = C.c_float(0)
        self._check(self.lib.pg_kmer_stats(self.h, cnt, C.byref(ms)))
        return dict(attempted=int(cnt[0]), mapped=int(cnt[1]), kmer_ms=float(ms.value))

    def path_stats(self):
        """-> dict(attempted, anchored, mapped, path_ms) of the last batch (PathAligner.hh:66-68)"""
        cnt = (C.c_uint64 * 4)()
        ms = C.c_float(0)
        self._check(self.lib.pg_path_stats(self.h, cnt, C.byref(ms)))
        return dict(attempted=int(cnt[0]), anchored=int(cnt[1]), mapped=int(cnt[2]), path_ms=float(ms.value),
                    index_build_ms=cnt[3] / 1e3)

    def add_graph(self, node_seqs, edges):
        blob = "".join(node_seqs).encode("latin-1")
        off = np.zeros(len(node_seqs) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(s) for s in node_seqs])
        ef = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
        et = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
        sid = C.c_int32(-1)
        self._check(self.lib.pg_add_graph(self.h, len(node_seqs), blob, _i32(off), len(edges), _i32(ef), _i32(et),
                                          C.byref(sid)))
        self._shape.append((len(node_seqs), len(edges)))
        return sid.value

    @staticmethod
    def pack_graphs(graphs):
        """[(node sequences, edges), ...] -> the flat arrays pg_add_graphs takes (node_ptr, blob, off, edge_ptr, ef, et)"""
        node_ptr, edge_ptr, lens, ef, et, seqs = [0], [0], [], [], [], []
        for nodes, edges in graphs:
            seqs.extend(nodes)
            lens.extend(len(x) for x in nodes)
            ef.extend(e[0] for e in edges)
            et.extend(e[1] for e in edges)
            node_ptr.append(len(lens))
            edge_ptr.append(len(ef))
        off = np.zeros(len(lens) + 1, dtype=np.int32)
        off[1:] = np.cumsum(lens)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        return i32(node_ptr), "".join(seqs).encode("latin-1"), off, i32(edge_ptr), i32(ef), i32(et)

    def add_graphs(self, graphs=None, packed=None):
        """Register many sites with one call (pg_add_graphs); returns the id of the first one."""
        node_ptr, blob, off, edge_ptr, ef, et = packed if packed is not None else self.pack_graphs(graphs)
        first = C.c_int32(-1)
        self._check(self.lib.pg_add_graphs(self.h, len(node_ptr) - 1, _i32(node_ptr), blob, _i32(off), _i32(edge_ptr),
                                           _i32(ef), _i32(et), C.byref(first)))
        return first.value

    def clear_graphs(self):
        self._check(self.lib.pg_clear_graphs(self.h))
        self._shape = []

    # ---- counting stage ------------------------------------------------------------------------
    def set_edge_labels(self, site, masks):
        """Path-family labels ("sequences") of a site's edges: one uint64 bit mask per edge, in add_graph's edge order."""
        m = None if masks is None else np.ascontiguousarray(masks, dtype=np.uint64)
        self._check(self.lib.pg_set_edge_labels(self.h, int(site),
                                                None if m is None else m.ctypes.data_as(C.POINTER(C.c_uint64))))

    def import_alignments(self, read_lens, records, ops, sites=None):
        """Make alignments produced elsewhere (records + op words) the context's current batch (pg_batch_import)."""
        rl = np.ascontiguousarray(read_lens, dtype=np.int32)
        rec = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        op = np.ascontiguousarray(ops, dtype=np.uint32)
        st = None if sites is None else np.ascontiguousarray(sites, dtype=np.int32)
        self._n = len(rl)
        self._check(self.lib.pg_batch_import(self.h, len(rl), _i32(rl), _i32(st) if st is not None else None,
                                             C.c_void_p(rec.ctypes.data), op.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             len(op)))

    def count(self, fragment=None, is_rev=None, remove_nonuniq=True, bad_align_frac=0.8, use_filters=True,
              family_slots=0, want_support=True):
        """Read filters + disambiguation + fragment counts of the batch last run (pg_batch_count).
        Returns dict(support, path_words, node_counts [sum n_nodes, 4], edge_counts [sum n_edges, 4],
        families {(site, mask): int array [1 + n_nodes + n_edges, 4]}); rows are site-major in add_graph order."""
        n = self._n
        prm = CountParams(int(remove_nonuniq), int(use_filters), float(bad_align_frac), int(family_slots), 0)
        fr = None if fragment is None else np.ascontiguousarray(fragment, dtype=np.int32)
        rv = None if is_rev is None else np.ascontiguousarray(is_rev, dtype=np.uint8)
        tn, te = sum(s[0] for s in self._shape), sum(s[1] for s in self._shape)
        nc = np.zeros((max(tn, 1), 4), dtype=np.uint32)
        ec = np.zeros((max(te, 1), 4), dtype=np.uint32)
        sup = np.zeros(max(n, 1), dtype=SUPPORT_DTYPE)
        path_cap = int(n * 64 + 4096)
        fam_cap = 1 << 16
        for attempt in range(3):
            pw = np.zeros(path_cap if want_support else 1, dtype=np.uint32)
            fw = np.zeros(fam_cap, dtype=np.uint32)
            pu, fu = C.c_uint64(0), C.c_uint64(0)
            rc = self.lib.pg_batch_count(
                self.h, _i32(fr) if fr is not None else None,
                rv.ctypes.data_as(C.POINTER(C.c_uint8)) if rv is not None else None, C.byref(prm),
                C.c_void_p(sup.ctypes.data) if want_support else None,
                pw.ctypes.data_as(C.POINTER(C.c_uint32)) if want_support else None, path_cap, C.byref(pu),
                C.c_void_p(nc.ctypes.data), len(nc), C.c_void_p(ec.ctypes.data), len(ec),
                fw.ctypes.data_as(C.POINTER(C.c_uint32)), fam_cap, C.byref(fu))
            if rc == -5 and attempt < 2 and b"needs" in self.lib.pg_last_error(self.h):
                path_cap = max(path_cap, int(n) * 520 + 4096)
                fam_cap = max(fam_cap * 16, int(fu.value) + 16)
                continue
            self._check(rc)
            break
        fams, w, words = {}, 0, fw[:fu.value]
        while w < len(words):
            k = int(words[w + 1])
            fams[(int(words[w]), int(words[w + 2]) | (int(words[w + 3]) << 32))] = \
                words[w + 4:w + 4 + 4 * k].astype(np.int64).reshape(k, 4)
            w += 4 + 4 * k
        return dict(support=sup[:n], path_words=pw[:pu.value], node_counts=nc[:tn].astype(np.int64),
                    edge_counts=ec[:te].astype(np.int64), families=fams)

    @staticmethod
    def pack_reads(reads, pinned=False):
        raw = np.frombuffer("".join(reads).encode("latin-1"), dtype=np.uint8)
        offs = np.zeros(len(reads) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(r) for r in reads])
        if not pinned:
            return raw.copy(), offs
        b, o = PinnedArray(raw.shape, np.uint8), PinnedArray(offs.shape, np.int32)
        b.array[:] = raw
        o.array[:] = offs
        b.array.flags.writeable = True
        keep = (b, o)
        blob, off = b.array, o.array
        Context._pinned_keep.append(keep)  # keep the page-locked buffers alive as long as the process
        return blob, off

    _pinned_keep = []

    # ---- staged API (bench) -------------------------------------------------------------------
    def upload(self, blob, off, sites=None):
        self._keep = (blob, off, sites)
        self._n = len(off) - 1
        self._check(self.lib.pg_batch_upload(self.h, self._n, C.c_void_p(blob.ctypes.data), _i32(off),
                                             _i32(sites) if sites is not None else None))

    def run(self, flags=AF_ALL):
        self._check(self.lib.pg_batch_run(self.h, flags & 0xFFFFFFFF))

    def _out_buffers(self, n, cigar_cap):
        cap = int(cigar_cap or n * 64 + 4096)
        if self._rec is None or len(self._rec) != n:
            self._rec_pin = PinnedArray((n,), RECORD_DTYPE)  # page-locked: D2H lands here without staging
            self._rec = self._rec_pin.array
        if self._ops is None or len(self._ops) < cap:
            self._ops_pin = PinnedArray((cap,), np.uint32)
            self._ops = self._ops_pin.array
        return self._rec, self._ops, len(self._ops)

    def download(self, cigar_cap=None):
        """Returns (records, ops) as views of buffers owned by the context: valid until the next call."""
        n = self._n
        for attempt in range(2):
            rec, ops, cap = self._out_buffers(n, cigar_cap)
            used = C.c_uint64(0)
            rc = self.lib.pg_batch_download(self.h, C.c_void_p(rec.ctypes.data),
                                            ops.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(used))
            if rc == -5 and attempt == 0 and cigar_cap is None:  # PG_E_CAPACITY: retry with the worst-case arena
                cigar_cap = n * 520 + 4096
                continue
            self._check(rc)
            return rec, ops[:used.value]

    # ---- one-call API (host buffers in, host buffers out) -----------------------------------------
    def align_packed(self, blob, off, sites=None, flags=AF_ALL, cigar_cap=None):
        """One pg_align_batch call.  Returns (records, ops) as views of buffers owned by the context (valid until
        the next call).  If the default CIGAR arena is too small the call is repeated once with a full-size one."""
        n = self._n = len(off) - 1
        for attempt in range(2):
            rec, ops, cap = self._out_buffers(n, cigar_cap)
            used = C.c_uint64(0)
            rc = self.lib.pg_align_batch(self.h, n, C.c_void_p(blob.ctypes.data), _i32(off),
                                         _i32(sites) if sites is not None else None, flags & 0xFFFFFFFF,
                                         C.c_void_p(rec.ctypes.data), ops.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         cap, C.byref(used))
            if rc == -5 and attempt == 0 and cigar_cap is None:  # PG_E_CAPACITY: adversarial CIGARs, take the worst case
                cigar_cap = n * 520 + 4096
                continue
            self._check(rc)
            return rec, ops[:used.value]

    def align(self, reads, sites=None, is_rev=None, flags=AF_ALL):
        """Align python strings; returns dicts with the fields GraphAligner::alignRead sets on common::Read."""
        from .synth import revcomp_exact
        if not reads:
            return []
        blob, off = self.pack_reads(reads)
        st = None if sites is None else np.ascontiguousarray(sites, dtype=np.int32)
        rec, ops = self.align_packed(blob, off, st, flags)
        out = []
        for i, r in enumerate(reads):
            x = rec[i]
            rv = bool(x["chose_reverse"])
            stage, flips = STAGES[int(x["mapped_by"])]  # which aligner of the cascade, after how many reverse complements
            by_path = stage == "path"  # PathAligner sets the strand itself and always writes a CIGAR
            for _ in range(flips):  # second chance(s): the stage saw bases earlier stages had reverse-complemented
                r = revcomp_exact(r)
            d = dict(pos=int(x["graph_pos"]), score=int(x["score"]), unique=bool(x["unique"]),
                     mapq=60 if x["unique"] else 0,
                     graph_reverse=rv if by_path else (bool(is_rev[i] if is_rev is not None else 0) != rv),
                     bases=revcomp_exact(r) if rv else r,
                     cigar=format_cigar(x, ops) if (flags & AF_CIGAR or stage != "gssw") else "", status=int(x["status"]),
                     clipped=int(x["query_clipped"]))
            if getattr(self, "_path_k", 0) or getattr(self, "_kmer_k", 0):
                d["stage"] = stage + ("2" if flips == 1 else "3" if flips == 2 else "")
            out.append(d)
        return out

    @staticmethod
    def read_filter(rec, read_lens, remove_nonuniq=True, bad_align_frac=0.8):
        """The reference's default read filter chain (createReadFilter, src/c++/lib/paragraph/ReadFilter.cpp:73-90:
        NonUniq then BadAlign) evaluated on the records, vectorised.  Returns (nonuniq, bad_align) boolean arrays;
        a read is filtered when either is set (NonUniq is tested first, like the chain does)."""
        L = np.asarray(read_lens, dtype=np.int64)
        nonuniq = (rec["unique"] == 0) if remove_nonuniq else np.zeros(len(rec), dtype=bool)
        aligned = L - rec["query_clipped"].astype(np.int64)
        thr = np.floor(bad_align_frac * L + 0.5)  # C round(): half away from zero, arguments are non-negative
        return nonuniq, aligned < thr

    def stats(self):
        n, a, b = C.c_uint64(0), C.c_float(0), C.c_float(0)
        self.lib.pg_stats(self.h, C.byref(n), C.byref(a), C.byref(b))
        cn, cm = C.c_uint64(0), C.c_float(0)
        self.lib.pg_count_stats(self.h, C.byref(cn), C.byref(cm))
        return dict(kernel_launches=n.value, fill_ms=a.value, trace_ms=b.value, count_launches=cn.value,
                    count_ms=cm.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
