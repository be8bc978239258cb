"""Deterministic ONT-like read simulator for BASELINE.json configs 1-2 (the bundled AMY loci).

The reference's own ONT read files are absent from the repository snapshot (.MISSING_LARGE_BLOBS); what
survives is the read-length index of the pseudodiploid sample (.test/data/AMY_hap{1,2}.ONT.fa.gz.fai), so reads
are simulated from the bundled haplotype assemblies with exactly those lengths.
The same code runs in tests/golden/make_golden_j1.py (where the reference's executables and scripts produce the
golden outputs) and in the GPU test (which regenerates the reads and checks their digest), so only the
assemblies, the length lists and the outputs are committed -- not 125 Mbp of reads.

numpy's PCG64 stream is stable across numpy versions for integers()/random() -- the digest test guards it.
"""
import hashlib

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _x, _y in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_x] = _y
_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutate(rng, seq: np.ndarray, sub=0.03, dele=0.02, ins=0.01) -> np.ndarray:
    """3 % substitutions, 2 % deletions, 1 % insertions (vectorised)"""
    n = len(seq)
    u = rng.random((n, 3))
    out = seq.copy()
    s = u[:, 1] < sub
    out[s] = _ALPHA[rng.integers(0, 4, int(s.sum()))]
    keep = u[:, 0] >= dele
    reps = keep.astype(np.int64) + ((u[:, 2] < ins) & keep)
    res = np.repeat(out, reps)
    # the second copy of a repeated base is the inserted (random) base
    idx = np.cumsum(reps) - 1
    insm = reps == 2
    res[idx[insm]] = _ALPHA[rng.integers(0, 4, int(insm.sum()))]
    return res


def simulate(contigs, lengths, seed, prefix, forbid=None, skip=()):
    """contigs: [(name, bytes)]; lengths: read lengths to draw (clipped to the contig); forbid: optional
    (contig name, lo, hi): no read may cover both lo and hi of that contig (leaves the stretch unspanned);
    skip: contig names no read is drawn from (decoys).  Returns [(read name, bytes)] in generation order."""
    rng = np.random.default_rng(seed)
    contigs = [c for c in contigs if c[0] not in skip]
    arrs = [np.frombuffer(s.upper(), dtype=np.uint8) for _, s in contigs]
    clen = np.array([len(a) for a in arrs], dtype=np.int64)
    cum = np.cumsum(clen)
    reads = []
    for i, ln in enumerate(lengths):
        for _attempt in range(1000):
            g = int(rng.integers(0, int(cum[-1])))
            ci = int(np.searchsorted(cum, g, side="right"))
            a = arrs[ci]
            l = int(min(ln, len(a)))
            st = int(rng.integers(0, len(a) - l + 1))
            if forbid is not None and contigs[ci][0] == forbid[0] and st <= forbid[1] and st + l >= forbid[2]:
                continue
            break
        r = mutate(rng, a[st:st + l])
        if rng.random() < 0.5:
            r = _COMP[r[::-1]]
        reads.append((f"{prefix}{i:06d}", r.tobytes()))
    return reads


def split_round_robin(reads, nchunks):
    """read i goes to chunk i % nchunks (the record distribution of `rustybam fastq-split`, tagONT.smk:17)"""
    return [reads[c::nchunks] for c in range(nchunks)]


def fastq_text(reads) -> bytes:
    """what `seqtk seq -F '#'` hands to fastq-split: 4-line FASTQ, quality '#' per base (tagONT.smk:17)"""
    return b"".join(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"#" * len(s) + b"\n" for n, s in reads)


def digest(reads) -> str:
    h = hashlib.sha256()
    for n, s in reads:
        h.update(n.encode())
        h.update(b"\0")
        h.update(s)
        h.update(b"\n")
    return h.hexdigest()
