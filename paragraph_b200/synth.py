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
