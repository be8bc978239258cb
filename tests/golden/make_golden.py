#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE'S OWN PREBUILT ELF
BINARIES (workflow/scripts/{kmerpos_annot3,rlen,diag_filter_v3,diag_filter_step2}) in this
container.  /root/reference does not exist on the GPU box, so the inputs and the ELF outputs are
committed as small gzip'd JSON fixtures; tests only read those.

Run from the repo root:   python tests/golden/make_golden.py
(The SUNK db / loc inputs are made with the oracle's DB restatement because jellyfish/mrsfast/
bedtools are not installed; the ELF outputs recorded here are what pins the oracle's match and
diag stages.)
"""
import gzip
import itertools
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gavisunk_oracle as O  # noqa: E402

REF = "/root/reference/workflow/scripts"


def run_kmerpos(reads_txt: bytes, db_txt: str, loc_txt: str, reads_name="reads.fa"):
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, reads_name), "wb").write(reads_txt)
        open(os.path.join(d, "db.txt"), "w").write(db_txt)
        open(os.path.join(d, "loc.txt"), "w").write(loc_txt)
        p = subprocess.run([f"{REF}/kmerpos_annot3", reads_name, "db.txt", "loc.txt", "out.txt"], cwd=d,
                           stderr=subprocess.PIPE)
        out = open(os.path.join(d, "out.txt")).read() if os.path.exists(os.path.join(d, "out.txt")) else ""
        return p.returncode, out


def run_rlen(reads_txt: bytes):
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "reads.fa"), "wb").write(reads_txt)
        subprocess.run([f"{REF}/rlen", "reads.fa", "out.txt"], cwd=d, check=True)
        return open(os.path.join(d, "out.txt")).read()


def run_diag(sunkpos_txt: str, fai_txt: str):
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.sunkpos"), "w").write(sunkpos_txt)
        open(os.path.join(d, "h.fai"), "w").write(fai_txt)
        diag = subprocess.run([f"{REF}/diag_filter_v3", "s.sunkpos", "h.fai"], cwd=d, check=True,
                              stdout=subprocess.PIPE).stdout.decode()
        open(os.path.join(d, "s.diag"), "w").write(diag)
        diag2 = subprocess.run([f"{REF}/diag_filter_step2", "s.sunkpos", "s.diag"], cwd=d, check=True,
                               stdout=subprocess.PIPE).stdout.decode()
        return diag, diag2


def save(name, obj):
    with gzip.open(os.path.join(HERE, name + ".json.gz"), "wt") as f:
        json.dump(obj, f)
    print("wrote", name, os.path.getsize(os.path.join(HERE, name + ".json.gz")), "bytes")


# ------------------------------------------------------------------------------------------------
def kat_b1():
    db = "AAAC\nAACT\nCCCG\n"
    loc = "c1\t100\tAAAC\t100\nc1\t101\tAACT\t100\nc1\t200\tCCCG\t200\n"
    reads = (b">r1\nTTAAACTTCCCGTT\n>r2\nCCCGTTAAAC\n>r3 revcomp of r1\nAACGGGAAGTTTAA\n"
             b">r4\nTTAAANCTTCCCGNTTAAAC\n")
    rc, out = run_kmerpos(reads, db, loc)
    return dict(k=4, db=db, loc=loc, reads=reads.decode("latin-1"), out=out, rc=rc)


def kat_bytes():
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rcs = lambda s: "".join(comp[c] for c in reversed(s))
    kms = sorted({min("".join(p), rcs("".join(p))) for p in itertools.product("ACGT", repeat=4)})
    db = "\n".join(kms) + "\n"
    loc = "".join(f"c1\t{100 * i}\t{km}\t{100 * i}\n" for i, km in enumerate(kms))
    reads = b""
    for b in range(1, 256):
        if b in (10, 13, 62, 64):
            continue
        reads += b">b%d\nCC" % b + bytes([b]) + b"GT\n"
    # short reads (Q6) and len == k
    reads += b">s3\nACG\n>s4\nACGT\n>s5\nACGTA\n"
    rc, out = run_kmerpos(reads, db, loc)
    return dict(k=4, db=db, loc=loc, reads=reads.decode("latin-1"), out=out, rc=rc)


def mutate(rng, seq: np.ndarray, sub=0.03, dele=0.02, ins=0.01):
    out = []
    u = rng.random(len(seq) * 3).reshape(-1, 3)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i, b in enumerate(seq):
        if u[i, 0] < dele:
            continue
        if u[i, 1] < sub:
            b = alphabet[rng.integers(4)]
        out.append(b)
        if u[i, 2] < ins:
            out.append(alphabet[rng.integers(4)])
    return np.array(out, dtype=np.uint8)


def rc_bytes(a: np.ndarray):
    m = np.zeros(256, dtype=np.uint8)
    m[:] = ord("N")
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        m[x] = y
    return m[a[::-1]]


def make_asm(rng, n_contigs, L, h=0.004, dup=True):
    """two related haplotypes, SNP rate h, optional segmental duplication + N run"""
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap1, hap2 = [], []
    for c in range(n_contigs):
        s = alphabet[rng.integers(0, 4, L)]
        if dup and L >= 4000:
            a = rng.integers(0, L // 2 - 600)
            b = rng.integers(L // 2, L - 600)
            s[b:b + 500] = s[a:a + 500]
        if c == 0 and L >= 2000:
            s[L // 3:L // 3 + 37] = ord("N")
            s[L // 5:L // 5 + 60] = np.frombuffer(s[L // 5:L // 5 + 60].tobytes().lower(), dtype=np.uint8)
        t = s.copy()
        snp = rng.random(L) < h
        t[snp] = alphabet[(O.BASE_LUT[s[snp]] + rng.integers(1, 4, int(snp.sum()))) % 4]
        hap1.append((f"h1c{c}", s.tobytes()))
        hap2.append((f"h2c{c}", t.tobytes()))
    return hap1, hap2


def db_loc_text(contigs, k):
    db = O.build_sunk_db(contigs, k)
    names = [n for n, _ in contigs]
    kms = [O.decode(int(x), k) for x in db["kmer"]]
    # jellyfish.db is in hash order: shuffle deterministically
    order = np.random.default_rng(7).permutation(len(kms))
    db_txt = "".join(kms[i] + "\n" for i in order)
    loc_txt = "".join(f"{names[c]}\t{s}\t{km}\t{g}\n" for c, s, km, g in zip(db["contig"], db["start"], kms, db["group"]))
    return db_txt, loc_txt


def sim_reads(rng, hap, n_reads, mean_len, prefix, nrun_every=7):
    reads = []
    for i in range(n_reads):
        name, seq = hap[rng.integers(len(hap))]
        a = np.frombuffer(seq, dtype=np.uint8)
        ln = int(min(len(a), max(30, rng.lognormal(np.log(mean_len), 0.5))))
        st = rng.integers(0, len(a) - ln + 1)
        r = mutate(rng, a[st:st + ln])
        if rng.random() < 0.5:
            r = rc_bytes(r)
        if i % nrun_every == 3 and len(r) > 200:
            r = r.copy()
            r[100:100 + rng.integers(1, 30)] = ord("N")
        reads.append((f"{prefix}{i:05d}", r.tobytes()))
    return reads


def fasta_text(reads, fastq=False, width=0):
    out = []
    for n, s in reads:
        if fastq:
            out.append(b"@" + n.encode() + b" desc\n" + s + b"\n+\n" + b"#" * len(s) + b"\n")
        elif width:
            out.append(b">" + n.encode() + b"\n" + b"\n".join(s[i:i + width] for i in range(0, len(s), width)) + b"\n")
        else:
            out.append(b">" + n.encode() + b"\n" + s + b"\n")
    return b"".join(out)


def random_case(seed, k, n_contigs, L, n_reads, mean_len, nchunks=2):
    rng = np.random.default_rng(seed)
    hap1, hap2 = make_asm(rng, n_contigs, L)
    contigs = hap1 + hap2
    db_txt, loc_txt = db_loc_text(contigs, k)
    case = dict(k=k, db=db_txt, loc=loc_txt,
                fai1="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap1),
                fai2="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap2),
                asm1=[(n, s.decode()) for n, s in hap1], asm2=[(n, s.decode()) for n, s in hap2],
                chunks=[])
    for hapi, hap in ((1, hap1), (2, hap2)):
        reads = sim_reads(rng, hap, n_reads, mean_len, f"h{hapi}r")
        # a few chimeric reads hitting several contigs so diag_filter has real choices
        for j in range(0, len(reads), 5):
            o = reads[(j + 1) % len(reads)][1]
            reads[j] = (reads[j][0], reads[j][1] + o[:len(o) // 2])
        for ci in range(nchunks):
            part = reads[ci::nchunks]
            txt = fasta_text(part, fastq=(ci % 2 == 1), width=(70 if ci == 0 else 0))
            rc, sunkpos = run_kmerpos(txt, db_txt, loc_txt)
            assert rc == 0
            fai = case["fai1"] if hapi == 1 else case["fai2"]
            diag, diag2 = run_diag(sunkpos, fai)
            case["chunks"].append(dict(hap=hapi, reads=txt.decode("latin-1"), sunkpos=sunkpos, diag=diag,
                                       diag2=diag2, rlen=run_rlen(txt)))
    return case


def ragged_case(seed, k):
    """reads that stress the tile / boundary logic of a batched matcher: empty reads, reads of length 1 .. k+3
    (k-1 is the bogus-window quirk Q6, k is the shortest real window), hundreds of tiny reads per 512 bases,
    error-free copies (every SUNK of a group hits: long runs of suppressed hits, position drift Q3), reads that
    end in N, mixed case; one FASTA chunk and one FASTQ chunk, through the reference executables"""
    rng = np.random.default_rng(seed)
    hap1, hap2 = make_asm(rng, 2, 6000)
    contigs = hap1 + hap2
    db_txt, loc_txt = db_loc_text(contigs, k)
    case = dict(k=k, db=db_txt, loc=loc_txt,
                fai1="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap1),
                fai2="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap2), chunks=[])
    for hapi, hap in ((1, hap1), (2, hap2)):
        reads = []
        a = np.frombuffer(hap[0][1], dtype=np.uint8)
        b = np.frombuffer(hap[1][1], dtype=np.uint8)
        for i in range(420):
            src = a if i % 3 else b
            if i % 10 == 0:
                ln = [0, 1, k - 2, k - 1, k - 1, k, k, k + 1, k + 3, 2 * k][(i // 10) % 10]
            elif i % 10 < 7:
                ln = int(rng.integers(1, 3 * k))
            else:
                ln = int(rng.integers(200, 1500))
            st = int(rng.integers(0, len(src) - ln + 1))
            r = src[st:st + ln].copy()          # error-free: dense runs of hits
            if i % 4 == 1 and ln > 60:
                r = mutate(rng, r)
            if i % 2:
                r = rc_bytes(r)
            if i % 9 == 4 and len(r) > 5:
                r[-3:] = ord("N")
            if i % 7 == 2:
                r = np.frombuffer(r.tobytes().lower(), dtype=np.uint8)
            reads.append((f"g{hapi}r{i:04d}", r.tobytes()))
        for ci in range(2):
            part = reads[ci::2]
            txt = fasta_text(part, fastq=(ci == 1))
            rc, sunkpos = run_kmerpos(txt, db_txt, loc_txt)
            assert rc == 0
            diag, diag2 = run_diag(sunkpos, case["fai1"] if hapi == 1 else case["fai2"])
            case["chunks"].append(dict(hap=hapi, reads=txt.decode("latin-1"), sunkpos=sunkpos, diag=diag, diag2=diag2,
                                       rlen=run_rlen(txt)))
    return case


def diag_cases():
    cases = []
    fai = "cA\t1000000\t0\t60\t61\ncB\t1000000\t0\t60\t61\nchr1\t1\t0\t60\t61\nchr2\t1\t0\t60\t61\n"

    def rows(read, contig, ds, groups=None, base=10000):
        out = ""
        for i, d in enumerate(ds):
            p = base * (i + 1)
            g = groups[i] if groups else p + d
            out += f"{read}\t{p}\t{contig}\t{p + d}\t{g}\n"
        return out
    # B.5 numerics
    s = rows("q1", "cA", [0, 0, 4998, 4998])
    s += rows("q2", "cA", [-4999] * 3 + [0] * 3, groups=[1, 1, 2, 3, 4, 5])
    s += rows("q3", "cA", [0, 0, 0, 2500])
    s += rows("q4", "cA", [0, 0, 0, 2499])
    s += rows("q5", "notinfai", [0, 0, 0, 0, 0]) + rows("q5", "cA", [0, 0])
    s += rows("q6", "cA", [0, 1, 2], groups=[7, 7, 7])
    s += rows("q7", "cA", [0, 0]) + rows("q7", "cB", [0, 0])
    s += rows("q8", "cB", [0, 0]) + rows("q8", "cA", [0, 0])
    s += rows("q9", "chr1", [0, 0]) + rows("q9", "chr2", [0, 0])
    # reverse-orientation read: pos+start constant
    s += "".join(f"q10\t{1000 * i}\tcA\t{500000 - 1000 * i}\t{500000 - 1000 * i}\n" for i in range(1, 9))
    # negative start-pos
    s += "".join(f"q11\t{100000 + 1000 * i}\tcA\t{1000 * i + (3 if i % 2 else 0)}\t{1000 * i}\n" for i in range(1, 8))
    # equal good, smaller hitlen wins
    s += rows("q12", "cA", [0, 0, 90000]) + rows("q12", "cB", [0, 0])
    s += rows("q13", "cB", [0, 0]) + rows("q13", "cA", [0, 0, 90000])
    diag, diag2 = run_diag(s, fai)
    cases.append(dict(name="numerics", sunkpos=s, fai=fai, diag=diag, diag2=diag2))

    # B.6 table order: random contig-name pairs with identical geometry
    rng = np.random.default_rng(11)
    names = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "h1tg000001l", "h1tg000002l", "cA", "cB", "AMY_h1",
                                                "AMY_h2", "AMY_orphan_h1", "nonuniq_kmers", "synH1_chr1", "synH2_chr1",
                                                "ptg000001l"] + [f"h2tg{i:06d}l" for i in range(1, 60)]
    fai2 = "".join(f"{n}\t1000000\t0\t60\t61\n" for n in names)
    s = ""
    for t in range(150):
        a, b = rng.choice(len(names), 2, replace=False)
        s += rows(f"t{t:04d}", names[a], [0, 0]) + rows(f"t{t:04d}", names[b], [0, 0])
    # three-way and interleaved insertion orders
    for t in range(60):
        sel = rng.choice(len(names), 3, replace=False)
        rr = f"u{t:04d}"
        for rep in range(2):
            for c in sel:
                p = 10000 * (rep + 1)
                s += f"{rr}\t{p + int(c)}\t{names[c]}\t{p}\t{p}\n"
    diag, diag2 = run_diag(s, fai2)
    cases.append(dict(name="table_order", sunkpos=s, fai=fai2, diag=diag, diag2=diag2))

    # table growth: a read touching N distinct contigs, then tie reads (Q9 growth persistence)
    for N in (43, 44, 50, 90):
        s = ""
        for c in range(N):
            s += f"big\t{100 * c}\tg{c}\t5\t5\n"
        s += rows("big", "cA", [0, 0])
        for t in range(80):
            a, b = rng.choice(len(names), 2, replace=False)
            s += rows(f"v{t:04d}", names[a], [0, 0]) + rows(f"v{t:04d}", names[b], [0, 0])
        diag, diag2 = run_diag(s, fai2 + "cA\t1\t0\t60\t61\n")
        cases.append(dict(name=f"growth{N}", sunkpos=s, fai=fai2 + "cA\t1\t0\t60\t61\n", diag=diag, diag2=diag2))
    # growth in the middle of a read whose contigs are all in the fai (re-insertion order matters)
    big_names = [f"k{i:03d}" for i in range(120)]
    fai3 = "".join(f"{n}\t1000000\t0\t60\t61\n" for n in big_names)
    s = ""
    for rep in range(3):
        sel = rng.permutation(120)[:rng.integers(40, 120)]
        rr = f"w{rep}"
        for rr_rep in range(2):
            for c in sel:
                p = 10000 * (rr_rep + 1)
                s += f"{rr}\t{p + int(c)}\t{big_names[c]}\t{p}\t{p}\n"
        for t in range(30):
            a, b = rng.choice(120, 2, replace=False)
            s += rows(f"w{rep}t{t:03d}", big_names[a], [0, 0]) + rows(f"w{rep}t{t:03d}", big_names[b], [0, 0])
    diag, diag2 = run_diag(s, fai3)
    cases.append(dict(name="growth_mid", sunkpos=s, fai=fai3, diag=diag, diag2=diag2))
    return cases


def rlen_case():
    txt = (b">m1\tdesc tab\nACGT\nACGT\n\nAC\n>m2 crlf\r\nACGT\r\nAC\r\n>m3\n>m4 emptyprev\nAC GT\n@q1\nACGT\n+\n!!!!\n")
    return dict(reads=txt.decode("latin-1"), rlen=run_rlen(txt))


def covprob_case():
    """read-length distributions of the bundled ONT read indexes (.test/data/AMY_hap{1,2}.ONT.fa.gz.fai:
    the reads themselves are missing from the snapshot) + the oracle's covprob table for SUNK_len 20
    (anchors in SURVEY.md A.8 were cross-checked against sympy.solveset as covprob.py:68-81 uses it)"""
    out = {}
    T = "/root/reference/.test/data/"
    for hap in (1, 2):
        rl = [(l.split("\t")[0], int(l.split("\t")[1])) for l in open(f"{T}AMY_hap{hap}.ONT.fa.gz.fai")]
        G = sum(int(l.split("\t")[1]) for l in open(f"{T}AMY.hap{hap}.fa.gz.fai")) / 1000
        out[f"hap{hap}"] = dict(read_lens=rl, genome_kbp=G, table_r20=O.covprob_table(rl, G, 20))
    return out


def hg02723_asm():
    """the bundled HG02723 AMY-locus assemblies (inputs of the README known answer, README.md:36-37)"""
    T = "/root/reference/.test/data/HG02723/"
    return dict(h1=open(T + "h1.fa").read(), h2=open(T + "h2.fa").read())


KSWEEP = [2, 3, 5, 6, 7, 8, 9, 11, 12, 13, 15, 17, 19, 21, 23, 25, 27, 28, 29, 30, 32]


def ksweep_case(seed, k):
    """every SUNK_len the library accepts besides the ones of BASELINE.json's sweep (16/20/24/31, the rand_* and
    ragged_* cases) and the k=4 KATs: small related haplotypes sized to the k-mer space (so that SUNKs exist even for
    k=2), reads that are error-free substrings (dense hit runs), mutated substrings, random sequence, reads of length
    k-1 / k / k+1, lower case and N tails; k=32 is the reference's mask-overflow case (Q2).  No read is shorter than
    k-1: the reference then reads whatever the previous records left in readfq's buffer (Q6, undefined), which at
    small k hits SUNKs by chance (the ragged_* cases hold such reads at k = 16 / 20 / 31, where it cannot).  One FASTA
    and one FASTQ chunk per haplotype through the reference executables."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    L = {2: 7, 3: 14, 5: 120, 6: 400, 7: 1200}.get(k, 3000)
    hap1, hap2 = [], []
    for c in range(2):
        s = alphabet[rng.integers(0, 4, L)]
        t = s.copy()
        snp = rng.random(L) < max(0.004, 2.0 / L)
        snp[int(rng.integers(0, L))] = True
        t[snp] = alphabet[(O.BASE_LUT[s[snp]] + rng.integers(1, 4, int(snp.sum()))) % 4]
        hap1.append((f"h1c{c}", s.tobytes()))
        hap2.append((f"h2c{c}", t.tobytes()))
    contigs = hap1 + hap2
    db_txt, loc_txt = db_loc_text(contigs, k)
    case = dict(k=k, db=db_txt, loc=loc_txt,
                fai1="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap1),
                fai2="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap2), chunks=[])
    for hapi, hap in ((1, hap1), (2, hap2)):
        reads = []
        for i in range(64):
            src = np.frombuffer(hap[i % 2][1], dtype=np.uint8)
            if i % 8 == 0:
                ln = [k - 1, k, k + 1, k + 2][(i // 8) % 4]
            elif i % 8 < 4:
                ln = int(rng.integers(k - 1, 3 * k + 2))
            else:
                ln = int(rng.integers(4 * k, 16 * k + 40))
            ln = max(min(ln, len(src)), k - 1)
            st = int(rng.integers(0, len(src) - ln + 1))
            r = src[st:st + ln].copy()
            if i % 4 == 1 and ln > 3 * k:
                r = mutate(rng, r)
            if i % 8 == 7:
                r = alphabet[rng.integers(0, 4, int(rng.integers(k, 300)))]  # random sequence: hits by chance (small k)
            if i % 4 == 1 and len(r) < k - 1:
                r = src[:k - 1].copy()  # a mutated read may have lost bases
            if i % 2:
                r = rc_bytes(r)
            if i % 9 == 4 and len(r) > 5:
                r[-2:] = ord("N")
            if i % 7 == 2:
                r = np.frombuffer(r.tobytes().lower(), dtype=np.uint8)
            reads.append((f"s{hapi}r{i:03d}", r.tobytes()))
        for ci in range(2):
            part = reads[ci::2]
            txt = fasta_text(part, fastq=(ci == 1))
            rc, sunkpos = run_kmerpos(txt, db_txt, loc_txt)
            assert rc == 0, (k, rc)
            case["chunks"].append(dict(hap=hapi, reads=txt.decode("latin-1"), sunkpos=sunkpos, rlen=run_rlen(txt)))
    return case


def main():
    if sys.argv[1:] == ["ksweep"]:
        save("ksweep", [ksweep_case(300 + k, k) for k in KSWEEP])
        return
    save("covprob_amy", covprob_case())
    save("hg02723_asm", hg02723_asm())
    save("kat_b1", kat_b1())
    save("kat_bytes", kat_bytes())
    save("rlen_b8", rlen_case())
    save("diag_cases", diag_cases())
    save("rand_k20", random_case(101, 20, 2, 12000, 24, 3000))
    save("rand_k16", random_case(102, 16, 1, 8000, 10, 2500))
    save("rand_k24", random_case(103, 24, 3, 6000, 14, 2500))
    save("rand_k31", random_case(104, 31, 1, 9000, 10, 3000))
    save("rand_k20_many", random_case(105, 20, 6, 5000, 30, 6000, nchunks=3))
    save("ragged_k20", ragged_case(106, 20))
    save("ragged_k31", ragged_case(107, 31))
    save("ragged_k16", ragged_case(108, 16))


if __name__ == "__main__":
    main()
