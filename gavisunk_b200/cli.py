"""Command-line shims with the reference's argv on top of the GPU engine.

  python -m gavisunk_b200.cli kmerpos_annot3 <reads.fa|fq[.gz]> <db.txt> <kmer.loc> <out.sunkpos>
  python -m gavisunk_b200.cli rlen <reads> <out.rlen>
  python -m gavisunk_b200.cli diag_filter_v3 <sunkpos> <asm.fai>            (stdout)
  python -m gavisunk_b200.cli diag_filter_step2 <sunkpos[.gz]> <diag[.gz]>  (stdout)
  python -m gavisunk_b200.cli badsunks_AR <fai1> <fai2> <sunkpos1> <sunkpos2> <out>
  python -m gavisunk_b200.cli process_by_contig <locs> <sunkpos> <rlen> <badsunks> <out.tsv> <out.bed> [--minlen N] [--opt_filt]
  python -m gavisunk_b200.cli get_gaps <fai1> <fai2> <sample> <indir/> <outdir/>
  python -m gavisunk_b200.cli slop_gaps <gaps.bed> <asm.fai> <out.bed>
  python -m gavisunk_b200.cli covprob --bed B --locs L --rlen R --fai F --sunk-len K --tsv OUT
  python -m gavisunk_b200.cli split_locs --ont-pos P --kmer-loc L --flag results/S/breaks/hapN_splits_pos.done --hap hapN
  python -m gavisunk_b200.cli fused --sample S --k K --hap1-asm .. --hap2-asm .. --hap1-reads f.. --hap2-reads f.. --outdir D

The first four replace workflow/scripts/{kmerpos_annot3,rlen,diag_filter_v3,diag_filter_step2}
(workflow/rules/tagONT.smk:36,73,92,110) one to one; the next six replace badsunks_AR.py,
process-by-contig_lowmem_AR.py, get_gaps.py, `bedtools slop`, covprob.py and split_locs.py with the
argv of their rules (tagONT.smk:150,189,230,249; the two `script:` rules take their snakemake.input /
.output / .config values as options, and `covprob_script(snakemake)` / `split_locs_script(snakemake)`
accept the snakemake object itself); `fused` replaces the rules from jellyfish_count
to get_gaps and writes the same files under results/{sample}/.  Text parsing / formatting is host
plumbing; all per-base and per-row work runs in libgavisunk_b200.so.  Errors exit non-zero with a
message on stderr and never leave a partial output file behind.
"""
from __future__ import annotations

import argparse
import os
import sys
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import io as gio
from .engine import Engine, encode_kmers


def _atomic_write(path: str, text: str):
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        f.write(text)
    os.replace(tmp, path)


def _fmt_rows(rows, read_names, contig_names) -> str:
    out = []
    for r, p, c, s, g in zip(rows["read"].tolist(), rows["pos"].tolist(), rows["contig"].tolist(), rows["start"].tolist(),
                             rows["group"].tolist()):
        out.append(f"{read_names[r]}\t{p}\t{contig_names[c]}\t{s}\t{g}\n")
    return "".join(out)


def _k_from_db(db_lines: Sequence[bytes]) -> int:
    return len(db_lines[0]) if db_lines else 0  # kmerpos_annot3.nim:28-38: k = length of the first db line


def cmd_kmerpos_annot3(argv):
    reads_p, db_p, loc_p, out_p = argv
    db_lines = [l.encode() for l in gio.read_db(db_p)]
    sys.stderr.write("FINISHED READING KMERS\n")
    contig, start, kmer, group = gio.read_loc(loc_p)
    sys.stderr.write("FINISHED READING LOCS\n")
    k = _k_from_db(db_lines)
    eng = Engine(k if 1 <= k <= 32 else 20)
    eng.load_loc_text(db_lines, list(zip(contig, start, [km.encode() for km in kmer], group)))
    reads = gio.NativeReads([reads_p])  # multi-threaded gz/parse into page-locked memory (csrc/fastx.cu)
    eng.set_reads(reads.seq, reads.read_off)
    eng.match()
    _atomic_write(out_p, _fmt_rows(eng.rows(0), reads.names, eng.contig_names))


def cmd_rlen(argv):
    reads_p, out_p = argv
    reads = gio.NativeReads([reads_p], pin=False)
    _atomic_write(out_p, "".join(f"{n}\t{l}\n" for n, l in zip(reads.names, reads.lengths().tolist())))


def _rows_from_sunkpos(path):
    rows = gio.read_sunkpos(path)
    cnames, cidx, rnames, ridx, prev = [], {}, [], [], None
    for r in rows:
        if r[2] not in cidx:
            cidx[r[2]] = len(cnames)
            cnames.append(r[2])
        if r[0] != prev:  # consecutive rows with equal names form a read (diag_filter_v3.nim:64)
            rnames.append(r[0])
            prev = r[0]
        ridx.append(len(rnames) - 1)
    cols = (np.asarray(ridx, np.uint32), np.asarray([r[1] for r in rows], np.uint32),
            np.asarray([cidx[r[2]] for r in rows], np.uint32), np.asarray([r[3] for r in rows], np.uint32),
            np.asarray([r[4] for r in rows], np.uint32))
    return rows, cnames, rnames, cols


def cmd_diag_filter_v3(argv):
    sunkpos_p, fai_p = argv
    rows, cnames, rnames, cols = _rows_from_sunkpos(sunkpos_p)
    fai = {n for n, _ in gio.read_fai(fai_p)}
    eng = Engine(20)
    eng.contig_names = cnames
    eng.set_reads_meta(np.zeros(len(rnames), np.uint32))
    eng.set_rows(0, *cols)
    eng.diag_filter([0 if n in fai else 255 for n in cnames])
    b = eng.best()
    out = sys.stdout
    for r, c, g, d in zip(b["read"].tolist(), b["contig"].tolist(), b["ngood"].tolist(), b["dir"].tolist()):
        out.write(f"{rnames[r]}\t{cnames[c]}\t{g}\t{'+' if d else '-'}\t{g}\n")


def cmd_diag_filter_step2(argv):
    sunkpos_p, diag_p = argv
    rows, cnames, rnames, cols = _rows_from_sunkpos(sunkpos_p)
    best: Dict[str, str] = {}
    with gio.open_maybe_gz(diag_p) as f:
        for l in f.read().decode().splitlines():
            p = l.split("\t")
            if len(p) >= 2:
                best[p[0]] = p[1]  # later rows of a read name overwrite earlier ones (nim:17-22)
    cidx = {n: i for i, n in enumerate(cnames)}
    table = np.array([cidx.get(best.get(n, ""), 0xFFFFFFFF) for n in rnames], dtype=np.uint32)
    eng = Engine(20)
    eng.contig_names = cnames
    eng.set_reads_meta(np.zeros(len(rnames), np.uint32))
    eng.set_rows(0, *cols)
    eng.filter_best(table)
    sys.stdout.write(_fmt_rows(eng.rows(1), rnames, cnames))



# ------------------------------------------------------------------------------------------------
# per-rule shims of the Python stages
# ------------------------------------------------------------------------------------------------
def _index_names(names: Sequence[str], known: Dict[str, int] = None):
    idx = dict(known or {})
    out = np.zeros(len(names), np.uint32)
    for i, n in enumerate(names):
        j = idx.get(n)
        if j is None:
            j = idx[n] = len(idx)
        out[i] = j
    return out, idx


def _engine_with_rows(rows, contig_order: Sequence[str] = ()):
    """kept-row engine from parsed sunkpos rows: reads = distinct names in sorted order (pandas groupby
    order), each read's rows contiguous in file order."""
    rnames = sorted({r[0] for r in rows})
    ridx = {n: i for i, n in enumerate(rnames)}
    rows = sorted(rows, key=lambda r: ridx[r[0]])  # stable: file order inside a read
    cidx = {n: i for i, n in enumerate(contig_order)}
    ccol, cidx = _index_names([r[2] for r in rows], cidx)
    cnames = [None] * len(cidx)
    for n, i in cidx.items():
        cnames[i] = n
    eng = Engine(20)
    eng.contig_names = cnames
    cols = (np.asarray([ridx[r[0]] for r in rows], np.uint32), np.asarray([r[1] for r in rows], np.uint32), ccol,
            np.asarray([r[3] for r in rows], np.uint32), np.asarray([r[4] for r in rows], np.uint32))
    return eng, rows, rnames, cnames, cols


def bad_sunks_of_hap(sunkpos_p: str, fai_p: str, hap: int):
    """badsunks_AR.py:20-52 (hap1) / :56-93 (hap2) -> set of 'contig:ID'"""
    rows = gio.read_sunkpos(sunkpos_p)
    fai = {n for n, _ in gio.read_fai(fai_p)}
    eng, rows, rnames, cnames, cols = _engine_with_rows(rows)
    if not any(c in fai for c in cnames):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0 (mode of no correct-haplotype counts, badsunks_AR.py:43)")
    eng.set_reads_meta(np.zeros(len(rnames), np.uint32), [0, len(rnames)], [hap])
    eng.set_rows(1, *cols)
    # contigs of this haplotype's .fai are "correct"; the others only get the upper limit (:48-51)
    eng.set_contigs([hap if c in fai else 2 + hap for c in cnames])
    eng.group_hist()
    eng.bad_groups()
    sys.stdout.write(f"mean coverage {eng.modes[hap]}\n")
    g = eng.groups()
    return {f"{cnames[g['contig'][i]]}:{g['group'][i]}" for i in eng.bad_list().tolist()}


def cmd_badsunks(argv):
    fai1, fai2, sp1, sp2, out_p = argv
    bad = bad_sunks_of_hap(sp1, fai1, 0) | bad_sunks_of_hap(sp2, fai2, 1)
    sys.stdout.write(f"{len(bad)}\n")
    _atomic_write(out_p, "".join(b + "\n" for b in sorted(bad)))  # the reference writes set order


def process_by_contig(locs_p, sunkpos_p, rlen_p, badsunks_p, out_tsv, out_bed):
    """process-by-contig_lowmem_AR.py:30-260 for one breaks/{contig}_{hap}.sunkpos"""
    rlen = {}
    with gio.open_maybe_gz(rlen_p) as f:
        for l in f.read().decode().splitlines():
            p = l.split("\t")
            if len(p) >= 2:
                rlen.setdefault(p[0], int(p[1]))
    open(locs_p).close()  # the reference reads it (:54-57) but none of its outputs depend on it
    rows = gio.read_sunkpos(sunkpos_p)
    if not rows:
        raise ValueError("No columns to parse from file (empty sunkpos, process-by-contig_lowmem_AR.py:60)")
    contig = rows[0][2]  # sunkposcat['chrom'].unique()[0] (:62)
    with open(badsunks_p) as f:
        badset = set(f.read().splitlines())
    rows = list(dict.fromkeys(rows))  # drop_duplicates (:104) keeps the first of identical rows
    eng, rows, rnames, cnames, cols = _engine_with_rows(rows)
    eng.set_reads_meta(np.asarray([min(rlen.get(n, 0), 0xFFFFFFFF) for n in rnames], np.uint32))
    eng.set_rows(1, *cols)
    eng.set_contigs([0] * len(cnames))
    bc, bg = [], []
    cidx = {n: i for i, n in enumerate(cnames)}
    for b in badset:
        c, _, g = b.partition(":")  # ID2 = chrom + ":" + ID (:68)
        if c in cidx and g.isdigit() and int(g) <= 0xFFFFFFFF:
            bc.append(cidx[c])
            bg.append(int(g))
    eng.bad_set(bc, bg)
    eng.validate(10000)  # the reference overwrites --minlen with 10000 (:106)
    if eng.n_pairs == 0:  # :92-94, :202-204: contig name only, no bed (the rule touches it)
        _atomic_write(out_tsv, contig + "\n")
        return None
    pairs = eng.pairs()
    _atomic_write(out_tsv, "".join(f"{g}\t{rnames[r]}\n" for g, r in zip(pairs["group"].tolist(), pairs["read"].tolist())))
    eng.components_local()
    iv = eng.intervals()
    _atomic_write(out_bed, "".join(f"{contig}\t{s}\t{e}\n" for s, e in sorted(zip(iv["start"].tolist(), iv["end"].tolist()))))
    return iv


def cmd_process_by_contig(argv):
    ap = argparse.ArgumentParser(prog="gavisunk_b200.cli process_by_contig")
    for a in ("SUNKs", "sunkpos", "rlen", "badsunks", "outputfile", "outputbed"):
        ap.add_argument(a)
    ap.add_argument("--minlen", default=10000, type=int)
    ap.add_argument("--opt_filt", action="store_true")
    a = ap.parse_args(argv)
    process_by_contig(a.SUNKs, a.sunkpos, a.rlen, a.badsunks, a.outputfile, a.outputbed)


def cmd_get_gaps(argv):
    """get_gaps.py: paths are string-concatenated, so indir / outdir need their trailing slash (:30,70)"""
    fai1, fai2, sample, indir, outdir = argv
    fais = [gio.read_fai(fai1), gio.read_fai(fai2)]
    names, lens, hap_of = [], [], []
    for hap in range(2):
        for n, l in fais[hap]:
            names.append(n)
            lens.append(l)
            hap_of.append(hap)
    ic, is_, ie = [], [], []
    for ci, n in enumerate(names):
        bp = indir + n.replace("#", "_") + f"_hap{hap_of[ci] + 1}.bed"
        if not os.path.exists(bp):
            sys.stdout.write(f"No bed for {n}\n")
            continue
        with open(bp) as f:
            for l in f:
                p = l.rstrip("\n").split("\t")
                if len(p) >= 3:
                    ic.append(ci)
                    is_.append(int(p[1]))
                    ie.append(int(p[2]))
    eng = Engine(20)
    eng.contig_names = names
    eng.set_intervals(ic, is_, ie, len(names))
    gaps, nodata = eng.gaps(np.asarray(lens, np.uint32))
    for hap in range(2):
        g = [(names[c], s, e) for c, s, e in zip(gaps["contig"].tolist(), gaps["start"].tolist(), gaps["end"].tolist()) if hap_of[c] == hap]
        nd = [(names[c], 0, lens[c]) for c in nodata.tolist() if hap_of[c] == hap]
        sys.stdout.write(f"hap{hap + 1}\nbreaks:  {len(g)} {sum(e + 1 - s for _, s, e in g)}\nno data:  {len(nd)} {sum(l for _, _, l in nd)}\n")
        _atomic_write(outdir + f"hap{hap + 1}.nodata.bed", "".join(f"{a}\t{s}\t{e}\n" for a, s, e in nd))
        _atomic_write(outdir + f"hap{hap + 1}.gaps.bed", "".join(f"{a}\t{s}\t{e}\n" for a, s, e in g))


def _read_bed3(path):
    out = []
    with open(path) as f:
        for l in f:
            p = l.rstrip("\n").split("\t")
            if len(p) >= 3:
                out.append((p[0], int(p[1]), int(p[2])))
    return out


def cmd_slop_gaps(argv):
    """bedtools slop -i gaps -g fai -b 200000 (tagONT.smk:249)"""
    bed_p, fai_p, out_p = argv
    fai = gio.read_fai(fai_p)
    cidx = {n: i for i, (n, _) in enumerate(fai)}
    rows = _read_bed3(bed_p)
    for c, _, _ in rows:
        if c not in cidx:
            raise KeyError(f"chromosome {c} is not in the genome file")
    eng = Engine(20)
    s, e = eng.slop([cidx[c] for c, _, _ in rows], [r[1] for r in rows], [r[2] for r in rows], [l for _, l in fai], 200000)
    _atomic_write(out_p, "".join(f"{c}\t{a}\t{b}\n" for (c, _, _), a, b in zip(rows, s.tolist(), e.tolist())))


def _natural_key(s: str):
    import re
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def covprob(bed_p, locs_p, rlen_p, fai_p, sunk_len: int, tsv_p):
    """covprob.py:14-134"""
    from .engine import covprob_bins, covprob_pn
    genome_kbp = sum(l for _, l in gio.read_fai(fai_p)) / 1000
    gaps = _read_bed3(bed_p)
    contig, start, kmer, group = gio.read_loc(locs_p)
    seen_k, seen_g, gc, gi = set(), set(), [], []
    cidx: Dict[str, int] = {}
    for c, km, g in zip(contig, kmer, group):  # drop_duplicates(kmer) then drop_duplicates(ID2) (:36-38)
        if km in seen_k:
            continue
        seen_k.add(km)
        if (c, g) in seen_g:
            continue
        seen_g.add((c, g))
        gc.append(cidx.setdefault(c, len(cidx)))
        gi.append(g)
    lens = []
    with gio.open_maybe_gz(rlen_p) as f:
        for l in f.read().decode().splitlines():
            p = l.split("\t")
            if len(p) >= 2:
                lens.append((p[0], int(p[1])))
    kbp, cnt = covprob_bins(lens)
    eng = Engine(int(sunk_len) if 1 <= int(sunk_len) <= 32 else 20)
    table = eng.covprob_table(kbp, cnt, genome_kbp, covprob_pn(int(sunk_len)))
    # breaks.df: pyranges groups the rows by chromosome (natural order of the names), file order inside
    order = sorted(range(len(gaps)), key=lambda i: (_natural_key(gaps[i][0]), i))
    g_sorted = [gaps[i] for i in order]
    unknown = len(cidx)
    mg, pr = eng.covprob_gaps(gc, gi, [cidx.get(c, unknown) for c, _, _ in g_sorted], [s for _, s, _ in g_sorted],
                              [e for _, _, e in g_sorted], table)
    out = ["index\tChromosome\tStart\tEnd\ttype\tmax_gap\tcovprob\n"]
    for i, (c, s, e), m, p in zip(order, g_sorted, mg.tolist(), pr.tolist()):
        out.append(f"{i}\t{c}\t{s}\t{e}\tgap\t{m}\t{p!r}\n")
    _atomic_write(tsv_p, "".join(out))
    return table


def covprob_script(snakemake):
    """drop-in body for `script: ../scripts/covprob.py` (tagONT.smk:211-231)"""
    covprob(snakemake.input.bed, snakemake.input.locs, snakemake.input.rlen, snakemake.input.fai,
            int(snakemake.config["SUNK_len"]), snakemake.output.tsv)


def cmd_covprob(argv):
    ap = argparse.ArgumentParser(prog="gavisunk_b200.cli covprob")
    for a in ("bed", "locs", "rlen", "fai", "tsv"):
        ap.add_argument("--" + a, required=True)
    ap.add_argument("--sunk-len", type=int, required=True)
    a = ap.parse_args(argv)
    covprob(a.bed, a.locs, a.rlen, a.fai, a.sunk_len, a.tsv)


def split_locs(ont_pos_p, kmer_loc_p, flag_p, hap: str):
    """split_locs.py:5-22: byte-level partition by contig, no arithmetic (host only)"""
    dirname = os.path.dirname(flag_p)
    os.makedirs(dirname or ".", exist_ok=True)
    by_c: Dict[str, List[str]] = defaultdict(list)
    with gio.open_maybe_gz(ont_pos_p) as f:
        for l in f.read().decode().splitlines():
            p = l.split("\t")
            if len(p) >= 5:
                by_c[p[2]].append(l + "\n")
    loc_c: Dict[str, List[str]] = defaultdict(list)
    with open(kmer_loc_p) as f:
        for l in f:
            c = l.split("\t", 1)[0]
            if c in by_c:
                loc_c[c].append(l if l.endswith("\n") else l + "\n")
    for c, lines in by_c.items():
        _atomic_write(f"{dirname}/{c.replace('#', '_')}_{hap}.sunkpos", "".join(lines))
    for c, lines in loc_c.items():
        _atomic_write(f"{dirname}/{c.replace('#', '_')}_{hap}.loc", "".join(lines))


def split_locs_script(snakemake):
    split_locs(snakemake.input.ONT_pos, snakemake.input.kmer_loc, snakemake.output.flag, snakemake.wildcards.hap)


def cmd_split_locs(argv):
    ap = argparse.ArgumentParser(prog="gavisunk_b200.cli split_locs")
    ap.add_argument("--ont-pos", required=True)
    ap.add_argument("--kmer-loc", required=True)
    ap.add_argument("--flag", required=True)
    ap.add_argument("--hap", required=True)
    a = ap.parse_args(argv)
    split_locs(a.ont_pos, a.kmer_loc, a.flag, a.hap)
    open(a.flag, "a").close()  # touch() of the checkpoint rule (tagONT.smk:158)


# ------------------------------------------------------------------------------------------------
# fused: defineSUNKs.smk + tagONT.smk (SUNK_annot .. get_gaps) for one sample
# ------------------------------------------------------------------------------------------------
def decode_kmers(km: np.ndarray, k: int) -> List[str]:
    if len(km) == 0:
        return []
    shifts = (2 * np.arange(k - 1, -1, -1)).astype(np.uint64)
    codes = ((km[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    return letters.view(f"S{k}").ravel().astype(str).tolist()


def run_fused(k: int, hap_asm: Sequence[str], hap_reads: Sequence[Sequence[str]], outdir: str, device: int = 0):
    """hap_asm: two FASTA paths; hap_reads: two lists of chunk files in scatter order."""
    contigs, contig_hap, fai = [], [], [[], []]
    for hap in range(2):
        for n, s in gio.read_fastx(hap_asm[hap]):
            contigs.append((n, s))
            contig_hap.append(hap)
            fai[hap].append((n, len(s)))
    eng = Engine(k, device=device)
    eng.build_db(contigs)
    names = eng.contig_names
    files = list(hap_reads[0]) + list(hap_reads[1])
    chunk_hap = [0] * len(hap_reads[0]) + [1] * len(hap_reads[1])
    reads = gio.NativeReads(files)  # one chunk per file, in scatter order
    rnames, lens, chunk_first = reads.names, reads.lengths().tolist(), reads.chunk_first.tolist()
    eng.set_reads(reads.seq, reads.read_off, chunk_first, chunk_hap)
    iv = eng.run_all(contig_hap)
    gaps, nodata = eng.gaps(np.asarray([len(s) for _, s in contigs], dtype=np.uint32))
    # ---- files ----
    d = lambda *p: os.path.join(outdir, *p)
    for sub in ("db", "mrsfast", "sunkpos", "breaks", "inter_outs", "bed_files", "final_out"):
        os.makedirs(d(sub), exist_ok=True)
    db = eng.db_export()
    kmers = decode_kmers(db["kmer"], k)
    _atomic_write(d("db", "jellyfish.db"), "".join(km + "\n" for km in kmers))
    _atomic_write(d("db", "jellyfish.fa"), "".join(f">{km}\n{km}\n" for km in kmers))
    _atomic_write(d("mrsfast", "kmer.loc"), "".join(
        f"{names[c]}\t{s}\t{km}\t{g}\n" for c, s, km, g in zip(db["contig"].tolist(), db["start"].tolist(), kmers, db["group"].tolist())))
    kept = eng.rows(1)
    first_read_of_hap = [chunk_first[chunk_hap.index(h)] if h in chunk_hap else len(rnames) for h in (0, 1)] + [len(rnames)]
    hap_of_read = lambda r: 0 if r < first_read_of_hap[1] else 1
    for hap in range(2):
        lo, hi = (0, first_read_of_hap[1]) if hap == 0 else (first_read_of_hap[1], len(rnames))
        sel = (kept["read"] >= lo) & (kept["read"] < hi)
        sub = {c: v[sel] for c, v in kept.items()}
        _atomic_write(d("sunkpos", f"hap{hap + 1}.sunkpos"), _fmt_rows(sub, rnames, names))
        _atomic_write(d("sunkpos", f"hap{hap + 1}.rlen"), "".join(f"{rnames[r]}\t{lens[r]}\n" for r in range(lo, hi)))
    gidx_name = {}
    for c, g, gi in zip(db["contig"].tolist(), db["group"].tolist(), db["gidx"].tolist()):
        gidx_name[gi] = f"{names[c]}:{g}"
    _atomic_write(d("sunkpos", "bad_sunks.txt"), "".join(gidx_name[int(g)] + "\n" for g in eng.bad_list()))
    # per-contig files: breaks/*.sunkpos|.loc (split_locs.py:5-22), inter_outs/*.tsv, bed_files/*.bed
    safe = lambda n: n.replace("#", "_")
    by_c = defaultdict(list)
    for i, c in enumerate(kept["contig"].tolist()):
        by_c[c].append(i)
    for c, idx in by_c.items():
        idx = np.asarray(idx)
        sub = {col: v[idx] for col, v in kept.items()}
        _atomic_write(d("breaks", f"{safe(names[c])}_hap{contig_hap[c] + 1}.sunkpos"), _fmt_rows(sub, rnames, names))
    loc_by_c = defaultdict(list)
    for c, s, km, g in zip(db["contig"].tolist(), db["start"].tolist(), kmers, db["group"].tolist()):
        loc_by_c[c].append(f"{names[c]}\t{s}\t{km}\t{g}\n")
    for c, lines in loc_by_c.items():
        if c in by_c:
            _atomic_write(d("breaks", f"{safe(names[c])}_hap{contig_hap[c] + 1}.loc"), "".join(lines))
    pairs = eng.pairs()
    inter = defaultdict(list)
    for r, c, g in zip(pairs["read"].tolist(), pairs["contig"].tolist(), pairs["group"].tolist()):
        inter[c].append((rnames[r], r, g))
    bed = defaultdict(list)
    for c, s, e in zip(iv["contig"].tolist(), iv["start"].tolist(), iv["end"].tolist()):
        bed[c].append((names[c], s, e))
    for c in by_c:
        hapn = contig_hap[c] + 1
        rows = inter.get(c, [])
        if rows:
            # `for rname, g in grouped`: groupby sorts read names (process-by-contig_lowmem_AR.py:135-136);
            # the sort is stable, so the vertex order inside a read is preserved
            rows = sorted(rows, key=lambda t: t[0])
            _atomic_write(d("inter_outs", f"{safe(names[c])}_hap{hapn}.tsv"), "".join(f"{g}\t{n}\n" for n, _, g in rows))
            _atomic_write(d("bed_files", f"{safe(names[c])}_hap{hapn}.bed"), "".join(f"{a}\t{s}\t{e}\n" for a, s, e in bed.get(c, [])))
        else:  # "no usable reads": contig name only, bed touched empty by the rule (tagONT.smk:190)
            _atomic_write(d("inter_outs", f"{safe(names[c])}_hap{hapn}.tsv"), names[c] + "\n")
            _atomic_write(d("bed_files", f"{safe(names[c])}_hap{hapn}.bed"), "")
    for hap in range(2):
        hn = hap + 1
        valid = [f"{names[c]}\t{s}\t{e}\n" for c, s, e in zip(iv["contig"].tolist(), iv["start"].tolist(), iv["end"].tolist())
                 if contig_hap[c] == hap]
        _atomic_write(d("final_out", f"hap{hn}.valid.bed"), "".join(valid))
        g_rows = [(names[c], s, e) for c, s, e in zip(gaps["contig"].tolist(), gaps["start"].tolist(), gaps["end"].tolist())
                  if contig_hap[c] == hap]
        _atomic_write(d("final_out", f"hap{hn}.gaps.bed"), "".join(f"{a}\t{s}\t{e}\n" for a, s, e in g_rows))
        _atomic_write(d("final_out", f"hap{hn}.nodata.bed"), "".join(
            f"{names[c]}\t0\t{len(contigs[c][1])}\n" for c in nodata.tolist() if contig_hap[c] == hap))
        clen = {n: l for n, l in fai[hap]}
        _atomic_write(d("final_out", f"hap{hn}.gaps.slop.bed"), "".join(  # bedtools slop -b 200000 (tagONT.smk:249)
            f"{a}\t{max(0, s - 200000)}\t{min(clen[a], e + 200000)}\n" for a, s, e in g_rows))
    return eng


def cmd_fused(argv):
    ap = argparse.ArgumentParser(prog="gavisunk_b200.cli fused")
    ap.add_argument("--sample", default="sample")
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--hap1-asm", required=True)
    ap.add_argument("--hap2-asm", required=True)
    ap.add_argument("--hap1-reads", nargs="+", required=True)
    ap.add_argument("--hap2-reads", nargs="+", required=True)
    ap.add_argument("--outdir", required=True)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    run_fused(a.k, [a.hap1_asm, a.hap2_asm], [a.hap1_reads, a.hap2_reads], a.outdir, a.device)


COMMANDS = {"kmerpos_annot3": (cmd_kmerpos_annot3, 4), "rlen": (cmd_rlen, 2), "diag_filter_v3": (cmd_diag_filter_v3, 2),
            "diag_filter_step2": (cmd_diag_filter_step2, 2), "badsunks_AR": (cmd_badsunks, 5),
            "process_by_contig": (cmd_process_by_contig, None), "get_gaps": (cmd_get_gaps, 5), "slop_gaps": (cmd_slop_gaps, 3),
            "covprob": (cmd_covprob, None), "split_locs": (cmd_split_locs, None), "fused": (cmd_fused, None)}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in COMMANDS:
        sys.stderr.write(__doc__)
        return 2
    fn, nargs = COMMANDS[argv[0]]
    if nargs is not None and len(argv) - 1 != nargs:
        sys.stderr.write(f"{argv[0]}: expected {nargs} arguments\n")
        return 2
    try:
        fn(argv[1:])
    except Exception as ex:  # non-zero exit + message, nothing partial written
        sys.stderr.write(f"gavisunk_b200 {argv[0]}: {type(ex).__name__}: {ex}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
