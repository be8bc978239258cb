"""CPU oracle for the GAVISUNK hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain numpy / pure-Python restatement of the reference's algorithm for the
SUNK match + inter-SUNK validation path.  It exists so that the CUDA engine in
``gavisunk_b200`` can be checked for bit-exact parity.  It is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path (``gavisunk_b200``) never imports
anything from ``oracle/`` and has no CPU fallback.

Parity pinning (see DESIGN.md "Oracle"):
  * match / diag stages (a4-a8): pinned against the reference's own prebuilt ELF binaries
    (``/root/reference/workflow/scripts/{kmerpos_annot3,diag_filter_v3,diag_filter_step2,rlen}``)
    on the Appendix-B probe vectors and on seeded random inputs; the ELF outputs are committed
    under ``tests/golden/`` with the generating script ``tests/golden/make_golden.py``.
  * SUNK database build (a1-a3): pinned at the README gap coordinate
    (``README.md:36``  AMY_h1 284861/324276 adjacent group starts) -- otherwise the third-party
    tools (jellyfish 2.3.0, mrsfast 3.4.2, bedtools 2.30.0) are absent: "parity pinned at 2
    coordinates + published tool semantics".
  * Python stages (a10-a13, a15, a16): pinned against the reference's OWN scripts
    (``workflow/scripts/{badsunks_AR,split_locs,process-by-contig_lowmem_AR,get_gaps,covprob}.py``),
    run unmodified by ``tests/golden/make_golden_py.py`` on sunkpos rows produced by the reference
    executables; outputs committed as ``tests/golden/pystages_*.json.gz`` and checked by
    ``tests/test_oracle_pystages.py``.  graph-tool, pyranges, matplotlib and seaborn cannot be
    installed in the build container, so the scripts run against minimal stand-ins for exactly the
    calls they make (``tests/golden/refpy_stubs``), with pandas-1.3 behaviours restored by
    ``refpy_compat.py``; sympy / scipy / numpy / pandas are real.  What stays unpinned are tie rules
    that live INSIDE those third-party packages (pandas 1.3 ``value_counts`` tie order in the
    "multipos" clean-up, graph-tool's choice among equally large components, pyranges' row order):
    they follow SURVEY.md A.6 / the packages' documented behaviour.

All citations are file:line relative to /root/reference/.
"""
from __future__ import annotations

import math
from collections import OrderedDict, defaultdict

import numpy as np

# ------------------------------------------------------------------------------------------------
# a4: nim-kmer 0.2.6 encode / slide  (third-party, un-vendored; call sites
#     workflow/src/kmerpos_annot3.nim:24,68,88).  Mapping pinned by probing the ELF with every
#     byte value: A/a->0, C/c/0x01->1, G/g/0x02->2, T/t/U/u/0x03->3, everything else -> 0.
# ------------------------------------------------------------------------------------------------
BASE_LUT = np.zeros(256, dtype=np.uint8)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    BASE_LUT[ord(_ch)] = _v
    BASE_LUT[ord(_ch.lower())] = _v
BASE_LUT[1], BASE_LUT[2], BASE_LUT[3] = 1, 2, 3

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp_str(s: str) -> str:
    return "".join(_COMP[c] for c in reversed(s))


def encode(s) -> int:
    """2-bit big-endian pack, first base most significant (kmer.encode; kmerpos_annot3.nim:24)."""
    if isinstance(s, str):
        s = s.encode("latin-1")
    v = 0
    for b in s:
        v = (v << 2) | int(BASE_LUT[b])
    return v


def decode(v: int, k: int) -> str:
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def codes_of(seq: bytes) -> np.ndarray:
    return BASE_LUT[np.frombuffer(seq, dtype=np.uint8)]


def canonical_windows(codes: np.ndarray, k: int) -> np.ndarray:
    """canonical (min of fwd and reverse-complement) 2-bit value of every window
    (kmer.slide, kmerpos_annot3.nim:88).  len(codes)-k+1 values (0 if shorter)."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    c = codes.astype(np.uint64)
    f = np.zeros(n, dtype=np.uint64)
    r = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        f |= c[j:j + n] << np.uint64(2 * (k - 1 - j))
        r |= (np.uint64(3) - c[j:j + n]) << np.uint64(2 * j)
    return np.minimum(f, r)


# ------------------------------------------------------------------------------------------------
# a1-a3: SUNK database (defineSUNKs.smk:17,39,59-60,101,124-126)
# ------------------------------------------------------------------------------------------------
_ACGT_VALID = np.zeros(256, dtype=bool)
for _ch in "ACGTacgt":
    _ACGT_VALID[ord(_ch)] = True


def build_sunk_db(contigs, k):
    """contigs: list of (name, bytes) in ref.fa order (hap1 contigs then hap2, defineSUNKs.smk:17).
    Returns loc rows sorted by (contig order, start): dict of arrays
      contig (int32 index), start (int64), kmer (uint64 canonical), group (int64 merged-run start).
    jellyfish -C -U 1: canonical windows containing only ACGT (case-insensitive), count == 1
    over all contigs (defineSUNKs.smk:39); mrsfast -e 0 gives the single location
    (defineSUNKs.smk:101); bedtools merge merges overlapping or book-ended [start,start+k)
    (defineSUNKs.smk:125) and column 8 is the merged-run start (defineSUNKs.smk:126)."""
    all_k, all_c, all_p = [], [], []
    for ci, (_, seq) in enumerate(contigs):
        b = np.frombuffer(seq, dtype=np.uint8)
        n = len(b) - k + 1
        if n <= 0:
            continue
        valid = _ACGT_VALID[b]
        bad = np.concatenate([[0], np.cumsum(~valid)])
        ok = (bad[k:k + n] - bad[:n]) == 0
        canon = canonical_windows(BASE_LUT[b], k)
        pos = np.nonzero(ok)[0]
        all_k.append(canon[pos])
        all_c.append(np.full(len(pos), ci, dtype=np.int32))
        all_p.append(pos.astype(np.int64))
    if not all_k:
        z = np.zeros(0, dtype=np.int64)
        return dict(contig=z.astype(np.int32), start=z, kmer=z.astype(np.uint64), group=z)
    km = np.concatenate(all_k)
    cc = np.concatenate(all_c)
    pp = np.concatenate(all_p)
    order = np.argsort(km, kind="stable")
    kms = km[order]
    first = np.ones(len(kms), dtype=bool)
    first[1:] = kms[1:] != kms[:-1]
    last = np.ones(len(kms), dtype=bool)
    last[:-1] = kms[1:] != kms[:-1]
    uniq = order[first & last]
    uniq.sort()  # (contig, pos) order because windows were appended in that order
    cc, pp, km = cc[uniq], pp[uniq], km[uniq]
    grp = np.zeros(len(pp), dtype=np.int64)
    cur = -1
    # merged run: new group iff contig changes or start > previous start + k (book-end merges)
    newg = np.ones(len(pp), dtype=bool)
    if len(pp) > 1:
        newg[1:] = (cc[1:] != cc[:-1]) | (pp[1:] > pp[:-1] + k)
    starts_of_group = np.where(newg, pp, 0)
    idx = np.maximum.accumulate(np.where(newg, np.arange(len(pp)), 0))
    grp = pp[idx]
    del cur, starts_of_group
    return dict(contig=cc, start=pp, kmer=km, group=grp)


# ------------------------------------------------------------------------------------------------
# a5: kmerpos_annot3 (workflow/src/kmerpos_annot3.nim:12-97)
# ------------------------------------------------------------------------------------------------
def match_chunk(reads, db_kmers, loc_rows, k):
    """reads: list of (name, bytes) in file order (one chunk file).
    db_kmers: iterable of uint64 (encode() of each db line, kmerpos_annot3.nim:24).
    loc_rows: iterable of (contig_name, start, kmer_u64, group) in file order; later rows
    overwrite earlier ones (kmerpos_annot3.nim:68, Q7).
    Returns list of rows (read_name, pos, contig, start, group).
    Raises KeyError if a db k-mer that is hit has no loc row (kmerpos_annot3.nim:90, Q7)."""
    dbset = set(int(x) for x in db_kmers)
    coords = {}
    for (c, s, km, g) in loc_rows:
        coords[int(km)] = (c, int(s), int(g))
    out = []
    if k > 32:
        return out
    prev = ("", 0)  # kmerpos_annot3.nim:82-84, declared outside the read loop (Q4)
    # sorted key array for vectorised membership
    keys = np.array(sorted(dbset), dtype=np.uint64)
    for name, seq in reads:
        if len(seq) == k - 1:
            seq = seq + b"\x00"  # Q6: NUL terminator read as 'A' by the unchecked slice
        if len(seq) < k:
            continue  # shorter reads: reference reads heap garbage (undefined) -> no windows
        canon = canonical_windows(codes_of(seq), k)
        if k == 32 and len(canon) > 0:
            # Q2, pinned by tests/golden/kat_k32{,b} + ksweep[k=32] through the ELF.  nim-kmer's forward_add masks
            # with (1 shl 2k) - 1, which is 0 for k = 32: after the first window the forward value keeps only the
            # code of the incoming base, while the rolling reverse-complement value stays right.  In the first
            # window the forward value is exact and the reverse-complement value comes out with its last base
            # (the complement of the read's first base) read as T.  Hence window 0 yields min(fwd, rc | 3) and
            # window w >= 1 yields min(code(seq[w+31]), rc).
            c = codes_of(seq).astype(np.uint64)
            n = len(canon)
            f0 = np.uint64(0)
            for j in range(k):
                f0 |= c[j] << np.uint64(2 * (k - 1 - j))
            r = np.zeros(n, dtype=np.uint64)
            for j in range(k):
                r |= (np.uint64(3) - c[j:j + n]) << np.uint64(2 * j)
            canon = canon.copy()
            canon[0] = min(f0, r[0] | np.uint64(3))
            canon[1:] = np.minimum(c[k:k + n - 1], r[1:])
        if len(keys) == 0:
            continue
        idx = np.searchsorted(keys, canon)
        idx[idx == len(keys)] = 0
        hit = keys[idx] == canon
        i_adj = 0  # number of suppressed hits so far in this read (Q3)
        for w in np.nonzero(hit)[0]:
            c = int(canon[w])
            chrom, start, group = coords[c]  # KeyError == reference crash (Q7)
            cur = (chrom, group)
            if cur == prev:  # kmerpos_annot3.nim:92  `continue` skips `inc i`
                i_adj += 1
                continue
            out.append((name, int(w) - i_adj, chrom, start, group))
            prev = cur
    return out


# ------------------------------------------------------------------------------------------------
# Nim stdlib hashes/tables emulation for Q9 (diag_filter_v3.nim:54,74)
# ------------------------------------------------------------------------------------------------
def murmur3_32(data: bytes, seed: int = 0) -> int:
    """MurmurHash3_x86_32 (Nim >=1.4 `hash(string)`; pinned by SURVEY probe B.6)."""
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & 0xFFFFFFFF
    n = len(data)
    nblocks = n // 4
    for i in range(nblocks):
        kk = int.from_bytes(data[4 * i:4 * i + 4], "little")
        kk = (kk * c1) & 0xFFFFFFFF
        kk = ((kk << 15) | (kk >> 17)) & 0xFFFFFFFF
        kk = (kk * c2) & 0xFFFFFFFF
        h ^= kk
        h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    tail = data[4 * nblocks:]
    kk = 0
    if len(tail) >= 3:
        kk ^= tail[2] << 16
    if len(tail) >= 2:
        kk ^= tail[1] << 8
    if len(tail) >= 1:
        kk ^= tail[0]
        kk = (kk * c1) & 0xFFFFFFFF
        kk = ((kk << 15) | (kk >> 17)) & 0xFFFFFFFF
        kk = (kk * c2) & 0xFFFFFFFF
        h ^= kk
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def nim_hash(name: str) -> int:
    """Nim `hash(string)` as used by Table: murmur3 of the bytes; Table remaps hc==0."""
    h = murmur3_32(name.encode("utf-8"))
    # Nim Hash is a signed int (64-bit): the 32-bit value is zero-extended.  tables.nim: if hc == 0
    # the stored hash code becomes 314159265.
    return 314159265 if h == 0 else h


class NimTable:
    """Insertion/iteration-order model of Nim's `Table[string, T]` (open addressing, linear
    probing, power-of-two capacity; initTable() default -> 64 slots; growth checked before each
    insert as `cap*2 < count*3 or cap-count < 4`; `clear` keeps the capacity)."""

    def __init__(self, cap=64):
        self.cap = cap
        self.slots = [None] * cap  # (hash, key)
        self.count = 0

    def clear(self):
        self.slots = [None] * self.cap
        self.count = 0

    def _raw_insert(self, slots, hc, key):
        mask = len(slots) - 1
        i = hc & mask
        while slots[i] is not None:
            i = (i + 1) & mask
        slots[i] = (hc, key)

    def has(self, key):
        hc = nim_hash(key)
        mask = self.cap - 1
        i = hc & mask
        while self.slots[i] is not None:
            if self.slots[i][0] == hc and self.slots[i][1] == key:
                return True
            i = (i + 1) & mask
        return False

    def insert(self, key):
        if self.has(key):
            return
        if self.cap * 2 < self.count * 3 or self.cap - self.count < 4:
            new = [None] * (self.cap * 2)
            for s in self.slots:  # enlarge: re-insert in old slot order
                if s is not None:
                    self._raw_insert(new, s[0], s[1])
            self.slots = new
            self.cap *= 2
        self._raw_insert(self.slots, nim_hash(key), key)
        self.count += 1

    def keys(self):
        return [s[1] for s in self.slots if s is not None]


# ------------------------------------------------------------------------------------------------
# a7: diag_filter_v3 (workflow/src/diag_filter_v3.nim:18-229)
# ------------------------------------------------------------------------------------------------
def _trunc_median(vals):
    """arraymancer percentile(n, 50) (linear interpolation, float64) then int() truncation
    toward zero (diag_filter_v3.nim:86,113; Q8)."""
    a = sorted(vals)
    f = (len(a) - 1) / 2.0
    lo = int(math.floor(f))
    if f == lo:
        m = float(a[lo])
    else:
        m = a[lo] + (a[lo + 1] - a[lo]) * 0.5
    return int(m)  # Python int() truncates toward zero like Nim's int()


def diag_filter_v3(rows, hap_contigs, bandwidth=2500):
    """rows: list of (read, pos, contig, start, group) for one chunk, file order.
    hap_contigs: set of contig names in the haplotype .fai (diag_filter_v3.nim:30-37).
    Returns list of (read, contig, n, dir, n) (diag_filter_v3.nim:141)."""
    out = []
    table = NimTable()
    data = {}

    def evaluate(rname):
        maxgood, maxhitlen = 0, 0
        best, bestdir = "", ""
        for ctg in table.keys():  # Nim table slot order (Q9)
            p, s, g = data[ctg]
            hitlen = len(g)
            if len(p) == 1:  # diag_filter_v3.nim:82
                continue
            if len(p) == 0:
                # contig not in the hap fai: empty seqs; never wins (probe B.5)
                continue
            for sign, d in ((+1, "-"), (-1, "+")):  # reverse (p+s) first, then forward (s-p)
                n = [si + sign * pi for pi, si in zip(p, s)]
                med = _trunc_median(n)
                good = len({gi for ni, gi in zip(n, g) if abs(ni - med) < bandwidth})
                if good > maxgood:
                    maxgood, maxhitlen, best, bestdir = good, hitlen, ctg, d
                if good == maxgood and hitlen < maxhitlen:
                    maxgood, maxhitlen, best, bestdir = good, hitlen, ctg, d
        if maxgood > 1:
            out.append((rname, best, maxgood, bestdir, maxgood))

    prev = ""
    for (r, pos, ctg, start, grp) in rows:
        if r != prev:
            # NB the reference also evaluates once before the first row (empty table -> no output)
            evaluate(prev)
            table.clear()
            data = {}
        if not table.has(ctg):
            table.insert(ctg)
            data[ctg] = ([], [], [])
        if ctg in hap_contigs:
            data[ctg][0].append(int(pos))
            data[ctg][1].append(int(start))
            data[ctg][2].append(int(grp))
        prev = r
    evaluate(prev)
    return out


def diag_filter_step2(rows, diag_rows):
    """workflow/src/diag_filter_step2.nim:13-66: keep every row whose contig is the read's best
    contig; reads without a best contig vanish (getOrDefault -> "" never equals a contig)."""
    best = {}
    for d in diag_rows:
        best[d[0]] = d[1]
    return [r for r in rows if best.get(r[0], "") == r[2]]


# ------------------------------------------------------------------------------------------------
# a10: badsunks_AR.py (workflow/scripts/badsunks_AR.py:20-103)
# ------------------------------------------------------------------------------------------------
def bad_sunks_one_hap(rows, hap_contigs):
    cnt = OrderedDict()
    for (_, _, ctg, _, grp) in rows:
        key = (ctg, int(grp))
        cnt[key] = cnt.get(key, 0) + 1
    correct = {kk: v for kk, v in cnt.items() if kk[0] in hap_contigs}
    if not correct:
        raise IndexError("mode of empty series (badsunks_AR.py:43)")
    cc = defaultdict(int)
    for v in correct.values():
        cc[v] += 1
    top = max(cc.values())
    m = min(v for v, c in cc.items() if c == top)  # pandas mode() is sorted -> smallest
    limit = m + (m ** 0.5) * 4  # badsunks_AR.py:46
    bad = {kk for kk, v in correct.items() if v > limit or v < 2}
    bad |= {kk for kk, v in cnt.items() if kk[0] not in hap_contigs and v > limit}
    return bad, m


def bad_sunks(rows1, contigs1, rows2, contigs2):
    b1, _ = bad_sunks_one_hap(rows1, contigs1)
    b2, _ = bad_sunks_one_hap(rows2, contigs2)
    return b1 | b2  # badsunks_AR.py:97


# ------------------------------------------------------------------------------------------------
# a12-a13: process-by-contig_lowmem_AR.py (:50-260)
# ------------------------------------------------------------------------------------------------
def _uf_find(par, x):
    while par[x] != x:
        par[x] = par[par[x]]
        x = par[x]
    return x


def validate_read(rows):
    """rows: list of (pos, start, ID) for one read on one contig, already sorted by start (stable)
    and de-duplicated.  Returns the validated group IDs in graph-vertex order, or None if the
    read produces no output.  (process-by-contig_lowmem_AR.py:136-198)"""
    n = len(rows)
    edges = []  # (i, j) with mask2
    n_mask = [0, 0]
    cand = []
    for i in range(n):
        pi, si, _ = rows[i]
        for j in range(i + 1, n):
            pj, sj, _ = rows[j]
            ds = abs(si - sj)
            dp = abs(pi - pj)
            # float64 ratio test (:145-147) == integer predicate 9ds < 10dp < 11ds (Q12)
            if 9 * ds < 10 * dp < 11 * ds:
                sg = 1 if pi > pj else 0
                n_mask[sg] += 1
                cand.append((i, j, sg))
    if n_mask[0] + n_mask[1] < 1:
        return None  # :148
    orient = 1 if n_mask[1] > n_mask[0] else 0  # np.unique sorted + argmax -> tie = 0 (:151-152)
    edges = [(i, j) for (i, j, sg) in cand if sg == orient]
    # multipos (:161-181).  M = left endpoints in edge order, then right endpoints.
    M = [(rows[i][2], rows[i][0]) for (i, j) in edges] + [(rows[j][2], rows[j][0]) for (i, j) in edges]
    pos_by_id = OrderedDict()
    for (idv, p) in M:
        d = pos_by_id.setdefault(idv, OrderedDict())
        d[p] = d.get(p, 0) + 1
    bad_edges = set()
    for idv in sorted(pos_by_id):
        d = pos_by_id[idv]
        if len(d) <= 1:
            continue
        # value_counts().index[0]: highest count; ties -> first appearance in M (parity unpinned,
        # SURVEY A.6 step 4)
        best_c = max(d.values())
        goodpos = next(p for p, c in d.items() if c == best_c)
        for e, (i, j) in enumerate(edges):
            if rows[i][2] == idv and not (rows[i][0] == goodpos or rows[j][0] == goodpos):
                bad_edges.add(e)
    kept = [e for t, e in enumerate(edges) if t not in bad_edges]
    if not kept:
        return None  # reference would raise on an empty graph; counted as skipped (A.6 step 5)
    # graph on IDs; vertices by first appearance (source before target) (:189-192)
    vid = OrderedDict()
    for (i, j) in kept:
        for r in (i, j):
            idv = rows[r][2]
            if idv not in vid:
                vid[idv] = len(vid)
    par = list(range(len(vid)))
    for (i, j) in kept:
        a, b = _uf_find(par, vid[rows[i][2]]), _uf_find(par, vid[rows[j][2]])
        if a != b:
            par[max(a, b)] = min(a, b)
    roots = [_uf_find(par, v) for v in range(len(vid))]
    # label_components labels in order of lowest vertex; hist.argmax -> first largest
    size = defaultdict(int)
    for r in roots:
        size[r] += 1
    best_root = None
    for v in range(len(vid)):  # lowest-numbered vertex order == component label order
        r = roots[v]
        if best_root is None or size[r] > size[best_root]:
            best_root = r
    ids = list(vid.keys())
    return [ids[v] for v in range(len(vid)) if roots[v] == best_root]


def process_by_contig(rows, rlen, bad, contig, minlen=10000):
    """rows: list of (read, pos, contig, start, ID) of ONE contig (breaks/{contig}_{hap}.sunkpos).
    rlen: dict read -> length.  bad: set of (contig, ID).
    Returns (inter_rows [(ID, read)] or None, bed_rows [(contig, start, end)] or None).
    None for inter_rows == "contig name only" output; None for bed == no bed written."""
    rows = [r for r in rows if (r[2], int(r[4])) not in bad]  # :70-72
    ids_by_read = defaultdict(set)
    for r in rows:
        ids_by_read[r[0]].add(int(r[4]))
    multis = {r for r, s in ids_by_read.items() if len(s) > 1}  # :91
    if not multis:
        return None, None  # :92-94
    sub = [r for r in rows if r[0] in multis]
    sub.sort(key=lambda r: (r[0], int(r[3])))  # stable, (rname,start) (:100)
    seen = set()
    sub2 = []
    for r in sub:  # drop_duplicates (:104)
        t = (r[0], int(r[1]), r[2], int(r[3]), int(r[4]))
        if t in seen:
            continue
        seen.add(t)
        sub2.append(t)
    # the reference ignores its --minlen flag and hard-codes 10000 (:106-108, Q13); `minlen` here
    # only exists so that tests can exercise short synthetic reads
    sub2 = [r for r in sub2 if rlen.get(r[0], -1) >= minlen]
    by_read = OrderedDict()
    for r in sub2:
        by_read.setdefault(r[0], []).append((r[1], r[3], r[4]))
    inter = []
    for rname in sorted(by_read):  # groupby sorts by key (:135-136)
        ids = validate_read(by_read[rname])
        if ids is None:
            continue
        inter.extend((i, rname) for i in ids)
    if not inter:
        return None, None  # :202-204
    # contig-wide components (:219-252)
    par = {}

    def find(x):
        while par[x] != x:
            par[x] = par[par[x]]
            x = par[x]
        return x

    reads_ids = OrderedDict()
    for i, rname in inter:
        reads_ids.setdefault(rname, []).append(i)
    for rname, ids in reads_ids.items():
        # combinations(g['ID'], 2): a read with a single ID contributes no vertex at all
        if len(ids) < 2:
            continue
        for i in ids:
            par.setdefault(i, i)
        r0 = find(ids[0])
        for i in ids[1:]:
            ri = find(i)
            if ri != r0:
                par[ri] = r0
    comps = defaultdict(list)
    for i in par:
        comps[find(i)].append(i)
    regions = []
    for c in comps.values():
        if len(c) <= 2:  # :248
            continue
        regions.append((contig, min(c), max(c)))
    regions.sort(key=lambda t: t[1])
    return inter, regions


# ------------------------------------------------------------------------------------------------
# a15: get_gaps.py (:17-123)   a17: slop_gaps (tagONT.smk:249)
# ------------------------------------------------------------------------------------------------
def get_gaps(fai, beds):
    """fai: list of (contig, length) in .fai order.  beds: dict contig -> list of (start, end)
    (missing key == no bed file; empty list == empty bed).  Returns (gaps, nodata)."""
    gaps, nodata = [], []
    for ctg, ln in fai:
        iv = beds.get(ctg)
        if not iv:
            nodata.append((ctg, 0, ln))
            continue
        iv = sorted(iv)
        merged = []
        for s, e in iv:  # pyranges merge: overlapping (book-ended cannot occur, A.7)
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        for a, b in zip(merged[:-1], merged[1:]):
            gaps.append((ctg, a[1], b[0] - 1))  # :60-61
    return gaps, nodata


def slop_gaps(gaps, fai_len, b=200000):
    """bedtools slop -b 200000 clipped to [0, contig length] (tagONT.smk:249)."""
    return [(c, max(0, s - b), min(fai_len[c], e + b)) for c, s, e in gaps]


# ------------------------------------------------------------------------------------------------
# a16: covprob.py (:14-134)
# ------------------------------------------------------------------------------------------------
def covprob_root(r, p=0.94):
    """positive real root of 1 - x + q p^r x^(r+1) other than 1/p (covprob.py:68-73), by
    deflating the known root 1/p and Newton iteration in float64 / mpmath-free bisection."""
    q = 1 - p
    a = q * p ** r

    def f(x):
        return 1 - x + a * x ** (r + 1)

    # f is convex on x>0 with exactly two positive roots; the minimum is at x* = ((r+1)a)^(-1/r)
    xs = ((r + 1) * a) ** (-1.0 / r)
    pinv = 1 / p
    if pinv > xs:
        lo, hi = 1.0, xs  # other root left of the minimum
    else:
        lo, hi = xs, max(2.0, 4 * xs)
        while f(hi) < 0:
            hi *= 2
    # bisection on the bracket [lo,hi] where f changes sign
    flo = f(lo)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if (fm > 0) == (flo > 0):
            lo, flo = mid, fm
        else:
            hi = mid
    x = 0.5 * (lo + hi)
    for _ in range(4):  # polish
        x -= f(x) / (-1 + a * (r + 1) * x ** r)
    return x


def covprob_pn(r, p=0.94, n=30):
    q = 1 - p
    x = covprob_root(r, p)
    qn = ((1 - p * x) / (q * (r + 1 - r * x))) * (1 / (x ** (n + 1)))  # covprob.py:79
    return float(1 - qn)


def covprob_table(read_lens, genome_kbp, r):
    """read_lens: iterable of (name, len) rows (duplicates dropped, covprob.py:44).
    Returns list t[0..3499] with t[0] = 1.0 (covprob.py:86-100)."""
    pn = covprob_pn(r)
    seen = set()
    cnt = OrderedDict()
    lens = []
    for nm, ln in read_lens:
        if (nm, ln) in seen:
            continue
        seen.add((nm, ln))
        lens.append(ln)
    for ln in sorted(lens, reverse=True):  # sort_values(len, descending) then value_counts(sort=False)
        kbp = int(ln / 1000)
        cnt[kbp] = cnt.get(kbp, 0) + 1
    table = [1.0]
    for I in range(1, 3500):
        cum = 1.0
        for kbp, c in cnt.items():
            if I > kbp:
                prob = 1
            else:
                prob = (1 - (((kbp - I) / genome_kbp) * (pn ** 2))) ** c  # covprob.py:56-60
            cum = cum * prob
        table.append(1 - cum)
    return table


def covprob_gaps(gaps, loc_rows, table):
    """gaps: list of (contig, start, end).  loc_rows: (contig, start, kmer, ID) in kmer.loc file
    order.  Returns list of (contig, start, end, max_gap, covprob) (covprob.py:35-40,109-131)."""
    seen_k, seen_g = set(), set()
    groups = []  # (contig, ID, dist)
    prev_id = None
    for (c, s, km, g) in loc_rows:
        if km in seen_k:
            continue
        seen_k.add(km)
        if (c, g) in seen_g:
            continue
        seen_g.add((c, g))
        d = g if prev_id is None else g - prev_id  # diff over file order, across contigs
        groups.append((c, g, max(d, 0)))
        prev_id = g
    out = []
    for (c, s, e) in gaps:
        ds = [d for (cc, g, d) in groups if cc == c and (s - 2) < g < (e + 2)]
        mg = max(ds)  # ValueError if empty == reference error
        out.append((c, s, e, mg, table[int(mg / 1000)]))
    return out
