"""Synthetic SV-site workloads (graphs + reads) for the parity tests and bench.py.

Shapes follow SURVEY.md section 8(d): config 2 = 3-node DEL graph (500 bp flanks, D=300),
150 bp reads, 1 % substitutions, 1 % reads with a short indel, half reverse-complemented,
even reads from the ALT haplotype and odd ones from REF.  Graph shapes for DEL/INS/DUP/INV
follow what the reference's vcf2paragraph emits (src/python/lib/grm/vcfgraph/vcfgraph.py:174-204).
Everything is seeded numpy; nothing here touches the GPU.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def random_seq(rng, n):
    return ACGT[rng.integers(0, 4, size=int(n))].tobytes().decode()


def revcomp(s):
    return "".join(_COMP.get(c, "N") for c in reversed(s))


# ----------------------------------------------------------------------------- graphs

def del_graph(rng, flank=500, d=300):
    """LF -> D -> RF plus bypass LF -> RF."""
    return [random_seq(rng, flank), random_seq(rng, d), random_seq(rng, flank)], [(0, 1), (1, 2), (0, 2)]


def ins_graph(rng, flank=500, ins=100):
    """LF -> INS -> RF plus LF -> RF."""
    return [random_seq(rng, flank), random_seq(rng, ins), random_seq(rng, flank)], [(0, 1), (1, 2), (0, 2)]


def dup_graph(rng, flank=500, seg=200):
    """INS-shaped graph whose inserted node copies the reference segment that follows it."""
    lf = random_seq(rng, flank)
    refseg = random_seq(rng, seg)
    rf = refseg + random_seq(rng, max(1, flank - seg))
    return [lf, refseg, rf], [(0, 1), (1, 2), (0, 2)]


def inv_graph(rng, flank=500, seg=200):
    """Swap bubble LF -> {REFSEG | revcomp(REFSEG)} -> RF."""
    refseg = random_seq(rng, seg)
    return ([random_seq(rng, flank), refseg, revcomp(refseg), random_seq(rng, flank)],
            [(0, 1), (0, 2), (1, 3), (2, 3)])


def long_del_graph(rng, flank=150, keep=150):
    """vcf2paragraph-shaped long deletion: source 'X', LF, ref-start, ref-end (middle dropped, so
    ref-end is a second source branch), RF, sink 'X' -- cf. share/test-data/paragraph/long-del."""
    nodes = ["X", random_seq(rng, flank), random_seq(rng, keep), random_seq(rng, keep), random_seq(rng, flank), "X"]
    edges = [(0, 1), (0, 3), (1, 2), (1, 4), (3, 4), (4, 5), (2, 5)]
    return nodes, sorted(edges)


def bubble_graph(rng, n_nodes=None, max_len=60, p_edge=0.5, alphabet="ACGT"):
    """Random topologically sorted DAG for fuzzing (every non-first node gets >= 1 predecessor
    with probability 0.9, so there can be several sources)."""
    n = int(n_nodes or rng.integers(1, 7))
    a = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    nodes = [a[rng.integers(0, len(a), size=int(rng.integers(1, max_len + 1)))].tobytes().decode() for _ in range(n)]
    edges = set()
    for t in range(1, n):
        for f in range(t):
            if rng.random() < p_edge:
                edges.add((f, t))
        if not any(e[1] == t for e in edges) and rng.random() < 0.9:
            edges.add((int(rng.integers(0, t)), t))
    return nodes, sorted(edges)


def short_node_graphs(rng, n_graphs=6):
    """Graphs that exercise every branch of the fill's node events (pg_core.cuh: entry_word / seed_prefetch / node_event_pre):
    runs of 1-3 bp nodes (a lane crosses two boundaries within one 8-step sub-block), chain links, merges that include the
    node just finished, merges that do not, several sources, fan-in of up to five predecessors, between longer flanks."""
    a = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = lambda n: a[rng.integers(0, 4, size=int(n))].tobytes().decode()
    out = []
    for gi in range(n_graphs):
        lens = [int(rng.integers(40, 160))]
        for _ in range(int(rng.integers(6, 14))):
            lens.append(int(rng.choice([1, 1, 2, 3, 5, 9, 30])))
        lens.append(int(rng.integers(40, 160)))
        n = len(lens)
        edges = set()
        for t in range(1, n):
            mode = int(rng.integers(0, 5)) if t < n - 1 else 1
            if mode == 0:    # chain link
                edges.add((t - 1, t))
            elif mode == 1:  # merge including the node just finished
                edges.add((t - 1, t))
                for f in rng.choice(t, size=min(t, int(rng.integers(1, 5))), replace=False):
                    edges.add((int(f), t))
            elif mode == 2 and t >= 2:  # merge of older nodes only
                for f in rng.choice(t - 1, size=min(t - 1, int(rng.integers(1, 4))), replace=False):
                    edges.add((int(f), t))
            elif mode == 3 and gi % 2:  # another source
                pass
            else:
                edges.add((int(rng.integers(0, t)), t))
        out.append(([seq(l) for l in lens], sorted(edges)))
    return out


def site_graph(rng, kind, flank=None, sv_len=None):
    flank = int(flank if flank is not None else rng.integers(150, 501))
    sv_len = int(sv_len if sv_len is not None else rng.integers(20, 501))
    if kind == "DEL":
        return del_graph(rng, flank, sv_len)
    if kind == "INS":
        return ins_graph(rng, flank, sv_len)
    if kind == "DUP":
        return dup_graph(rng, flank, min(sv_len, flank - 1))
    if kind == "INV":
        return inv_graph(rng, flank, sv_len)
    raise ValueError(kind)


def _split(seq, max_len=300, pad=150):
    """graphUtils.split_ref_nodes / split_alt_nodes (src/python/lib/grm/vcfgraph/graphUtils.py:54-101): a node longer
    than max_len keeps its first and last `pad` bases as two nodes that are NOT connected to each other."""
    return [seq] if len(seq) <= max_len else [seq[:pad], seq[-pad:]]


def vcf_site_graph(rng, kind, sv_len, flank=150, max_len=300, pad=150):
    """One SV site shaped as the reference's vcf2paragraph / multigrmpy.py make it (SURVEY.md 8d "shape fidelity"):
    `flank` = read length bases of reference either side (multigrmpy.py:198), the padding base as its own 1-bp
    reference node, nodes longer than 300 bp cut to their first / last 150 bp with the middle dropped
    (graphUtils.py:54-101; the cut-off halves hang off source / sink), `source` / `sink` = the 1-base sequence "X"
    (GraphInput.cpp:81-89).  DEL / INS / DUP / INV as vcfgraph.py:174-204 (DUP = inserted copy of the reference
    segment, INV = swap bubble with the reverse complement).  cf. share/test-data/paragraph/long-del/*.json (7 nodes,
    8 edges, two source branches).  Node ids are topologically ordered."""
    lf, padb, rf = random_seq(rng, flank), random_seq(rng, 1), random_seq(rng, flank)
    seg = random_seq(rng, sv_len)
    if kind == "DEL":
        ref_parts, alt_parts = _split(seg, max_len, pad), None
    elif kind == "INS":
        ref_parts, alt_parts = None, _split(seg, max_len, pad)
    elif kind == "DUP":  # the reference segment follows the breakpoint; its copy is inserted in front of it
        rf = (seg + rf)[:max(flank, min(len(seg) + flank, max_len))]
        ref_parts, alt_parts = None, _split(seg, max_len, pad)
    elif kind == "INV":
        ref_parts, alt_parts = _split(seg, max_len, pad), _split(revcomp(seg), max_len, pad)
    else:
        raise ValueError(kind)
    # layout (topological): source, [tails of cut nodes: fed by source], LF, padding base, [ref heads], [alt heads], RF, sink
    nodes, edges = ["X"], []
    SRC = 0
    tails = []
    for parts in (ref_parts, alt_parts):
        if parts and len(parts) == 2:
            nodes.append(parts[1])
            tails.append(len(nodes) - 1)
            edges.append((SRC, tails[-1]))
        else:
            tails.append(None)
    nodes.append(lf)
    LF = len(nodes) - 1
    edges.append((SRC, LF))
    nodes.append(padb)
    PAD = len(nodes) - 1
    edges.append((LF, PAD))
    heads = []
    for parts in (ref_parts, alt_parts):
        if parts:
            nodes.append(parts[0])
            heads.append(len(nodes) - 1)
            edges.append((PAD, heads[-1]))
        else:
            heads.append(None)
    nodes.append(rf)
    RF = len(nodes) - 1
    nodes.append("X")
    SNK = len(nodes) - 1
    edges.append((RF, SNK))
    if ref_parts is None or alt_parts is None:
        edges.append((PAD, RF))  # the allele without sequence of its own: bypass edge
    for parts, hd, tl in ((ref_parts, heads[0], tails[0]), (alt_parts, heads[1], tails[1])):
        if not parts:
            continue
        if len(parts) == 1:
            edges.append((hd, RF))
        else:  # cut node: the head runs into the sink, the tail (fed by the source) continues into RF
            edges.append((hd, SNK))
            edges.append((tl, RF))
    # tails sit before LF in id order but their successor RF after: ids ascend along every edge
    assert all(f < t for f, t in edges)
    return nodes, sorted(set(edges))


def vcf_sites(seed, n_sites, kinds=("DEL", "INS", "DUP", "INV"), coverage=30, read_len=150, max_sv=1000, max_reads=None):
    """configs[3] shape: SV sites as vcf2paragraph shapes them (vcf_site_graph), ~coverage x reads over each site."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_sites):
        kind = kinds[i % len(kinds)]
        nodes, edges = vcf_site_graph(rng, kind, int(rng.integers(50, max_sv + 1)), flank=read_len)
        span = max(len(h) for h in haplotypes(nodes, edges))
        nr = max(8, int(coverage * span / read_len))
        if max_reads:
            nr = min(nr, max_reads)
        out.append((kind, nodes, edges, simulate_reads(rng, nodes, edges, nr, read_len, alternate=False)))
    return out


def long_node_sites(seed, n_sites=24, reads_per_site=1000, read_len=150, lo=1000, hi=10000):
    """configs[4]: INV / DUP graphs whose variant nodes are 1-10 kb, 1k reads per site."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_sites):
        kind = "INV" if i % 2 else "DUP"
        n = int(rng.integers(lo, hi + 1))
        nodes, edges = (inv_graph(rng, 500, n) if kind == "INV" else dup_graph(rng, n + 500, n))
        out.append((kind, nodes, edges, simulate_reads(rng, nodes, edges, reads_per_site, read_len, alternate=False)))
    return out


# ----------------------------------------------------------------------------- packed (vectorised) batches for bench.py
_COMP_LUT = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP_LUT[_a] = _b


def site_reads_packed(rng, nodes, edges, n_reads, read_len=150, sub=0.01, indel_frac=0.01, rc_frac=0.5):
    """simulate_reads(alternate=False) without the per-base Python loop: -> (uint8 bases of all reads back to back,
    int32 length per read).  Same model: random haplotype and start, 1 % substitutions, 1 % of the reads with one
    1-6 bp indel, half reverse-complemented."""
    haps = [np.frombuffer(h.encode("latin-1"), dtype=np.uint8) for h in haplotypes(nodes, edges)]
    pick = rng.integers(0, len(haps), size=n_reads)
    rows, lens = [None] * n_reads, np.zeros(n_reads, dtype=np.int32)
    for hi, H in enumerate(haps):
        idx = np.flatnonzero(pick == hi)
        if idx.size == 0:
            continue
        L = min(read_len, len(H))
        span = len(H) - L
        starts = rng.integers(0, max(1, span - 7), size=idx.size)
        win = H[starts[:, None] + np.arange(L)[None, :]].copy()
        m = rng.random(win.shape) < sub
        win[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
        for k in np.flatnonzero(rng.random(idx.size) < indel_frac):  # the rare read with an indel: serial
            frag = H[starts[k]:starts[k] + L + 8].tobytes().decode("latin-1")
            r = mutate(rng, frag, sub=sub, indel=1.0)[:L]
            if len(r) == L:
                win[k] = np.frombuffer(r.encode("latin-1"), dtype=np.uint8)
        rc = rng.random(idx.size) < rc_frac
        win[rc] = _COMP_LUT[win[rc]][:, ::-1]
        for j, i in enumerate(idx):
            rows[i] = win[j]
        lens[idx] = L
    return np.concatenate(rows), lens


def packed_sweep(seed, n_sites, kinds=("DEL", "INS", "DUP", "INV"), coverage=30, read_len=150, max_sv=1000, shaped=True):
    """configs[3]: the SV-site sweep as flat arrays, generated fast enough for 10k sites inside bench.py.
    -> dict(graphs=[(nodes, edges)], blob=uint8 bases, off=int32[n+1], site=int32[n], read_ptr=int32[n_sites+1],
    cost=int64 per site (4 L G summed over its reads), kinds=[...]).  `shaped`: vcf2paragraph-shaped graphs
    (vcf_site_graph), else the idealised 3- / 4-node ones (site_graph)."""
    rng = np.random.default_rng(seed)
    graphs, blobs, lens, read_ptr, cost, kd = [], [], [], [0], [], []
    for i in range(n_sites):
        kind = kinds[i % len(kinds)]
        if shaped:
            nodes, edges = vcf_site_graph(rng, kind, int(rng.integers(50, max_sv + 1)), flank=read_len)
            span = max(len(h) for h in haplotypes(nodes, edges))
        else:
            nodes, edges = site_graph(rng, kind)
            span = sum(len(x) for x in nodes)
        nr = max(8, int(coverage * span / read_len))
        b, l = site_reads_packed(rng, nodes, edges, nr, read_len)
        graphs.append((nodes, edges))
        blobs.append(b)
        lens.append(l)
        read_ptr.append(read_ptr[-1] + nr)
        cost.append(4 * int(l.sum()) * sum(len(x) for x in nodes))
        kd.append(kind)
    lens = np.concatenate(lens)
    off = np.zeros(len(lens) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    site = np.repeat(np.arange(n_sites, dtype=np.int32), np.diff(read_ptr))
    return dict(graphs=graphs, blob=np.concatenate(blobs), off=off, site=site, read_ptr=np.asarray(read_ptr, dtype=np.int32),
                cost=np.asarray(cost, dtype=np.int64), kinds=kd)


def sweep_subset(sw, site_ids):
    """The sites `site_ids` (ascending ids, e.g. one rank's LPT shard) of a packed sweep as a sweep of their own."""
    site_ids = np.asarray(site_ids, dtype=np.int64)
    rp = sw["read_ptr"]
    sel = np.concatenate([np.arange(rp[s], rp[s + 1]) for s in site_ids]) if len(site_ids) else np.zeros(0, dtype=np.int64)
    lens = np.diff(sw["off"])[sel]
    off = np.zeros(len(sel) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    src = sw["off"][:-1][sel]
    idx = np.repeat(src - off[:-1], lens) + np.arange(int(off[-1])) if len(sel) else np.zeros(0, dtype=np.int64)
    counts = np.diff(rp)[site_ids]
    return dict(graphs=[sw["graphs"][s] for s in site_ids], blob=sw["blob"][idx], off=off,
                site=np.repeat(np.arange(len(site_ids), dtype=np.int32), counts),
                read_ptr=np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), cost=sw["cost"][site_ids],
                kinds=[sw["kinds"][s] for s in site_ids])


def sweep_reads(sw, site_index):
    """python strings of one site's reads (tests, parity samples)"""
    a, b = sw["read_ptr"][site_index], sw["read_ptr"][site_index + 1]
    raw = sw["blob"].tobytes()
    return [raw[sw["off"][i]:sw["off"][i + 1]].decode("latin-1") for i in range(a, b)]


# The BASELINE.json workloads by name (shared by tests/ and bench.py so that both see the same batches).
def workload(name, scale=1.0):
    """-> list of (kind, nodes, edges, reads).  `scale` < 1 shrinks the number of sites (CPU-side tests)."""
    if name == "config2":
        nodes, edges, reads = config2(seed=42, n_reads=max(64, int(10000 * scale)))
        return [("DEL", nodes, edges, reads)]
    if name == "config3":  # 1k mixed DEL/INS sites <= 500 bp, 30x, idealised 3-node graphs
        return sites(seed=3, n_sites=max(2, int(1000 * scale)), kinds=("DEL", "INS"))
    if name == "config4_share":  # one GPU's eighth of the 10k-site DEL/INS/DUP/INV sweep, vcf2paragraph-shaped graphs
        return vcf_sites(seed=4, n_sites=max(4, int(1250 * scale)))
    if name == "config5":
        return long_node_sites(seed=5, n_sites=max(2, int(24 * scale)))
    raise ValueError(name)


def flatten_sites(site_list):
    """-> (reads, site id per read, cells = 4 L G summed) for a list of (kind, nodes, edges, reads) registered in order."""
    reads, sids, cells = [], [], 0
    for i, (_, nodes, _, rds) in enumerate(site_list):
        reads += rds
        sids += [i] * len(rds)
        cells += 4 * sum(len(r) for r in rds) * sum(len(n) for n in nodes)
    return reads, np.ascontiguousarray(sids, dtype=np.int32), cells


# ----------------------------------------------------------------------------- reads

def _successors(n, edges):
    succ = [[] for _ in range(n)]
    for f, t in edges:
        succ[f].append(t)
    return succ


def haplotypes(nodes, edges, limit=64):
    """All source->sink path sequences (up to `limit`)."""
    n = len(nodes)
    succ = _successors(n, edges)
    has_pred = {t for _, t in edges}
    out = []

    def walk(v, acc):
        if len(out) >= limit:
            return
        acc = acc + nodes[v]
        if not succ[v]:
            out.append(acc)
            return
        for w in succ[v]:
            walk(w, acc)

    for s in range(n):
        if s not in has_pred:
            walk(s, "")
    return out


def haplotype_paths(nodes, edges, limit=64):
    """Node-id lists of the source->sink paths, in the order haplotypes() enumerates them."""
    n = len(nodes)
    succ = _successors(n, edges)
    has_pred = {t for _, t in edges}
    out = []

    def walk(v, acc):
        if len(out) >= limit:
            return
        acc = acc + [v]
        if not succ[v]:
            out.append(acc)
            return
        for w in succ[v]:
            walk(w, acc)

    for s in range(n):
        if s not in has_pred:
            walk(s, [])
    return out


def haplotype_labels(nodes, edges, limit=8):
    """One uint64 label mask per edge, bit k = the edge lies on haplotype k (what vcf2paragraph writes as the
    edge's "sequences", e.g. REF / ALT)."""
    paths = haplotype_paths(nodes, edges, limit=limit)
    masks = []
    for f, t in edges:
        m = 0
        for k, p in enumerate(paths):
            if any(p[i] == f and p[i + 1] == t for i in range(len(p) - 1)):
                m |= 1 << k
        masks.append(m)
    return masks


def mutate(rng, s, sub=0.01, indel=0.0, max_indel=6, n_rate=0.0, alphabet="ACGT"):
    b = list(s)
    for i in range(len(b)):
        r = rng.random()
        if r < sub:
            b[i] = alphabet[int(rng.integers(0, len(alphabet)))]
        elif r < sub + n_rate:
            b[i] = "N"
    if indel > 0 and rng.random() < indel and len(b) > 2 * max_indel + 2:
        k = int(rng.integers(1, max_indel + 1))
        p = int(rng.integers(1, len(b) - k - 1))
        if rng.random() < 0.5:
            del b[p:p + k]
        else:
            b[p:p] = list(random_seq(rng, k))
    return "".join(b)


def simulate_reads(rng, nodes, edges, n_reads, read_len=150, sub=0.01, indel_frac=0.01, rc_frac=0.5,
                   n_rate=0.0, alternate=True):
    """Reads drawn from the graph's haplotypes.  With `alternate`, read i comes from haplotype
    i % n_haplotypes (config 2: even = ALT, odd = REF for the 3-node DEL graph whose
    haplotypes() order is [LF+D+RF, LF+RF] -> we flip so that even = shorter/ALT)."""
    haps = haplotypes(nodes, edges)
    haps.sort(key=len)
    reads = []
    for i in range(n_reads):
        h = haps[i % len(haps)] if alternate else haps[int(rng.integers(0, len(haps)))]
        if len(h) <= read_len + 8:
            frag = h
        else:
            st = int(rng.integers(0, len(h) - read_len - 7))
            frag = h[st:st + read_len + 8]
        r = mutate(rng, frag, sub=sub, indel=indel_frac, n_rate=n_rate)[:read_len]
        if rng.random() < rc_frac:
            r = revcomp(r)
        reads.append(r)
    return reads


def config2(seed=42, n_reads=10000, read_len=150, flank=500, d=300):
    """BASELINE.json configs[1]: single 3-node DEL graph (500 bp flanks), 10k synthetic 150 bp reads."""
    rng = np.random.default_rng(seed)
    nodes, edges = del_graph(rng, flank, d)
    reads = simulate_reads(rng, nodes, edges, n_reads, read_len)
    return nodes, edges, reads


def sites(seed, n_sites, kinds=("DEL", "INS"), coverage=30, read_len=150, max_reads=None):
    """configs[2]/[3]: many independent sites, ~coverage x (span + 2*read_len) / read_len reads each."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_sites):
        kind = kinds[i % len(kinds)]
        nodes, edges = site_graph(rng, kind)
        span = sum(len(s) for s in nodes)
        nr = max(8, int(coverage * span / read_len))
        if max_reads:
            nr = min(nr, max_reads)
        out.append((kind, nodes, edges, simulate_reads(rng, nodes, edges, nr, read_len, alternate=False)))
    return out


def fuzz_reads(rng, nodes, edges, n_reads, min_len=8, max_len=160, lower=0.02, iupac=0.02):
    """Adversarial reads for parity fuzzing: heavy substitutions, indels, Ns, IUPAC codes,
    lower-case bases, random (unrelated) reads, low-complexity repeats."""
    haps = haplotypes(nodes, edges) or [random_seq(rng, 50)]
    reads = []
    for _ in range(n_reads):
        L = int(rng.integers(min_len, max_len + 1))
        mode = rng.random()
        if mode < 0.1:
            r = random_seq(rng, L)
        elif mode < 0.2:
            unit = random_seq(rng, int(rng.integers(1, 4)))
            r = (unit * (L // len(unit) + 1))[:L]
        else:
            h = haps[int(rng.integers(0, len(haps)))]
            if len(h) > L:
                st = int(rng.integers(0, len(h) - L + 1))
                h = h[st:st + L + 6]
            r = mutate(rng, h, sub=float(rng.choice([0.0, 0.02, 0.1])), indel=float(rng.choice([0.0, 0.5, 1.0])),
                       max_indel=int(rng.integers(1, 9)), n_rate=float(rng.choice([0.0, 0.02])))[:L]
            if rng.random() < 0.3 and len(r) > 4:  # second indel
                r = mutate(rng, r, sub=0.0, indel=1.0, max_indel=3)
            if rng.random() < 0.5:
                r = revcomp(r)
        if not r:
            r = "A"
        b = list(r)
        for i in range(len(b)):
            x = rng.random()
            if x < lower:
                b[i] = b[i].lower()
            elif x < lower + iupac:
                b[i] = "RYKMSWBDHVNU="[int(rng.integers(0, 13))]
        reads.append("".join(b))
    return reads


def revcomp_exact(s):
    """graphtools::reverseComplement (SequenceOperations.cpp:66-89): case-sensitive, non-ACGT -> 'N'."""
    return revcomp(s)


def write_workload_file(path, site_list):
    """The text format tools/cpp/bench_mirror.cpp reads: sites = [(kind, nodes, edges, reads)]."""
    with open(path, "w") as f:
        f.write("SITES %d\n" % len(site_list))
        for _, nodes, edges, reads in site_list:
            f.write("SITE %d %d %d\n" % (len(nodes), len(edges), len(reads)))
            f.write("\n".join(nodes) + "\n")
            if edges:
                f.write("\n".join("%d %d" % (a, b) for a, b in edges) + "\n")
            if reads:
                f.write("\n".join(reads) + "\n")


def sweep_as_site_list(sw):
    return [(sw["kinds"][i], sw["graphs"][i][0], sw["graphs"][i][1], sweep_reads(sw, i)) for i in range(len(sw["graphs"]))]
