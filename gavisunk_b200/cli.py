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
  python -m gavisunk_b200.cli split_ont --reads R.. --out temp/S/reads/hapN_1-of-10.fq.gz ..   (seqtk seq -F '#' | rustybam fastq-split)
  python -m gavisunk_b200.cli combine_ont_nofilt <chunk.sunkpos>.. <hapN_detailed.sunkpos>
  python -m gavisunk_b200.cli fused --sample S --k K --hap1-asm .. --hap2-asm .. --hap1-reads f.. --hap2-reads f.. --outdir D
                                    [--devices 0,1,..] [--batch-files N] [--detailed]

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
    reads = gio.NativeReads([reads_p], packed=True)  # parse + 2-bit pack in one pass into page-locked memory (csrc/fastx.cu)
    eng.set_reads_packed(reads.words, reads.read_off)
    eng.match()
    rows = eng.rows(0)
    rtab, ctab = gio.NameTable(blob=reads.name_blob, off=reads.name_off), gio.NameTable(eng.contig_names)
    _write_bytes(out_p, gio.format_rows([("name", rows["read"], rtab), ("u32", rows["pos"]), ("name", rows["contig"], ctab),
                                         ("u32", rows["start"]), ("u32", rows["group"])]))


def cmd_rlen(argv):
    reads_p, out_p = argv
    reads = gio.NativeReads([reads_p], pin=False, packed=True)
    rtab = gio.NameTable(blob=reads.name_blob, off=reads.name_off)
    _write_bytes(out_p, gio.format_rows([("name", np.arange(reads.n_reads, dtype=np.uint32), rtab), ("u64", reads.lengths())]))


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
# split_ONT (workflow/rules/tagONT.smk:2-18): cat reads | seqtk seq -F '#' | rustybam fastq-split outs
# ------------------------------------------------------------------------------------------------
def split_ont(reads_paths: Sequence[str], out_paths: Sequence[str], threads: int = 0):
    """The scatter step of the reference: every record becomes a 4-line FASTQ record with a fake quality of '#'
    (`seqtk seq -F '#'`) and record i goes to output i mod N (`rustybam fastq-split`), outputs gzip-compressed when
    they end in .gz.  seqtk and rustybam are not installed here, so this step is PARITY UNPINNED: the round-robin
    distribution is rustybam's documented behaviour, and the header comment seqtk would keep is dropped (no
    consumer reads past the name: kmerpos_annot3.nim:85, rlen.nim:13).  Outputs that depend on the chunking (the
    prevLoc carry and the Table capacity of a chunk) are identical to the reference's whenever
    the chunk files are."""
    import gzip
    from concurrent.futures import ThreadPoolExecutor
    reads = gio.NativeReads(list(reads_paths), pin=False)
    n_out = len(out_paths)
    off = reads.read_off
    seq = reads.seq
    nt = gio.NameTable(blob=reads.name_blob, off=reads.name_off)
    no = nt.off.tolist()

    def write(c):
        opener = (lambda p: gzip.open(p, "wb", compresslevel=1)) if out_paths[c].endswith(".gz") else (lambda p: open(p, "wb"))
        tmp = out_paths[c] + ".tmp%d" % os.getpid()
        with opener(tmp) as f:
            for r in range(c, reads.n_reads, n_out):
                a, b = int(off[r]), int(off[r + 1])
                f.write(b"@" + nt.blob[no[r]:no[r + 1]] + b"\n")
                f.write(seq[a:b].tobytes())
                f.write(b"\n+\n" + b"#" * (b - a) + b"\n")
        os.replace(tmp, out_paths[c])
    with ThreadPoolExecutor(max_workers=threads or min(n_out, os.cpu_count() or 1)) as ex:
        list(ex.map(write, range(n_out)))
    reads.close()


def cmd_split_ont(argv):
    ap = argparse.ArgumentParser(prog="gavisunk_b200.cli split_ont")
    ap.add_argument("--reads", nargs="+", required=True)
    ap.add_argument("--out", nargs="+", required=True, help="temp/{sample}/reads/{hap}_{i}-of-{N}.fq.gz in scatter order")
    a = ap.parse_args(argv)
    split_ont(a.reads, a.out)


def cmd_combine_ont_nofilt(argv):
    """combine_ont_nofilt (workflow/rules/tagONT.smk:39-55): cat of the unfiltered chunk outputs in gather order"""
    *ins, out_p = argv
    tmp = out_p + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as o:
        for p in ins:
            with open(p, "rb") as f:
                while True:
                    blk = f.read(1 << 24)
                    if not blk:
                        break
                    o.write(blk)
    os.replace(tmp, out_p)


# ------------------------------------------------------------------------------------------------
# fused: defineSUNKs.smk + tagONT.smk (SUNK_annot .. slop_gaps) for one sample, on one or several GPUs
# ------------------------------------------------------------------------------------------------
def decode_kmers(km: np.ndarray, k: int) -> List[str]:
    if len(km) == 0:
        return []
    shifts = (2 * np.arange(k - 1, -1, -1)).astype(np.uint64)
    codes = ((km[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    return letters.view(f"S{k}").ravel().astype(str).tolist()


def _write_bytes(path: str, data):
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, path)


def plan_batches(files: Sequence[str], n_devices: int, batch_files: int = 0, batch_gbp: float = 12.0) -> List[List[int]]:
    """groups of consecutive chunk files (scatter order: hap1 chunks, then hap2 chunks).  batch_files > 0 fixes the
    group size; otherwise groups hold about `batch_gbp` Gbp (estimated from the file sizes: FASTQ = 2 bytes per
    base, gzip ~ 4x) and there are at least as many groups as devices when the files allow it."""
    n = len(files)
    if n == 0:
        return []
    if batch_files > 0:
        return [list(range(i, min(n, i + batch_files))) for i in range(0, n, batch_files)]

    def est_bases(p):
        try:
            sz = os.path.getsize(p)
        except OSError:
            return 0
        with open(p, "rb") as f:
            head = f.read(2)
        gz = head == b"\x1f\x8b"
        fq = p.endswith((".fq", ".fastq", ".fq.gz", ".fastq.gz"))
        return sz * (4 if gz else 1) // (2 if fq else 1)
    est = [est_bases(p) for p in files]
    want = max(1.0, min(batch_gbp * 1e9, sum(est) / max(1, n_devices)))
    out, cur, acc = [], [], 0
    for i, e in enumerate(est):
        if cur and acc + e > want * 1.05 and len(out) + 1 + (n - i) >= n_devices:
            out.append(cur)
            cur, acc = [], 0
        cur.append(i)
        acc += e
    if cur:
        out.append(cur)
    return out


class ThreadExchange:
    """The two exchanges of the path between the engines of one process (one host thread per GPU): the histogram
    sum and the forest gather of gavisunk_b200.parallel, done with torch copies between the devices (plumbing)."""

    def __init__(self, engines):
        import threading
        self.engines = engines
        self.n = len(engines)
        self.bar = threading.Barrier(self.n)
        self.slots = [None] * self.n
        self.keep = [None] * self.n
        self.failed = False

    def abort(self):
        self.failed = True
        self.bar.abort()

    def for_rank(self, r):
        x = self

        class _R:
            @staticmethod
            def allreduce_hist(ptr, n_groups):
                if x.n == 1 or not n_groups:
                    return
                import torch
                from .parallel import alias_device_array
                dev = torch.device("cuda", x.engines[r].device)
                x.engines[r].sync()
                x.slots[r] = alias_device_array(ptr, n_groups, "<i4", dev)
                x.bar.wait()
                if r == 0:
                    tot = x.slots[0].clone()
                    for p in range(1, x.n):
                        tot += x.slots[p].to(tot.device)
                    for p in range(x.n):
                        x.slots[p].copy_(tot.to(x.slots[p].device))
                    for p in range(x.n):
                        torch.cuda.synchronize(x.slots[p].device)
                x.bar.wait()

            @staticmethod
            def gather_forests(ptr, n_groups):
                if x.n == 1 or not n_groups:
                    return []
                import torch
                from .parallel import alias_device_array
                dev = torch.device("cuda", x.engines[r].device)
                x.engines[r].sync()
                x.slots[r] = alias_device_array(ptr, n_groups, "<i4", dev)
                x.bar.wait()
                peers = [x.slots[p].to(dev) for p in range(x.n) if p != r]
                torch.cuda.synchronize(dev)
                x.keep[r] = peers
                x.bar.wait()  # nobody unions into its own forest while a peer still copies it
                return [int(t.data_ptr()) for t in peers]
        return _R


def run_reads(engines, contig_hap, files: Sequence[str], file_hap: Sequence[int], batch_files: int = 0, batch_gbp: float = 12.0,
              detailed: bool = False, threads: int = 0, min_read_len: int = 10000, coll=None):
    """The read side of the fused rule on engines whose SUNK database is loaded: the chunk files (scatter order, with
    the haplotype each belongs to) are grouped into batches (plan_batches); engine i owns a contiguous block of the
    batches and runs Engine.run_batches on it (one host thread per GPU, two exchanges per run).  Reads are parsed
    and 2-bit packed in one pass (gio.NativeReads(packed=True)) while the previous batch is on the GPU.
    coll: exchange object of a multi-PROCESS run (gavisunk_b200.parallel.EngineExchange) when `engines` is this rank's
    single engine.  -> one result dict per engine (read names / lengths / chunk tables, kept rows, validated pairs,
    intervals)."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    from .parallel import shard_chunks
    for p in files:
        if not os.path.exists(p):
            raise FileNotFoundError(p)
    n_dev = len(engines)
    batches = plan_batches(files, n_dev, batch_files, batch_gbp)
    shards = shard_chunks(len(batches), n_dev)
    n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ingest_threads = threads or max(1, n_cpu // n_dev)
    results = [None] * n_dev
    errors = []
    xch = ThreadExchange(engines)

    def worker(di):
        try:
            eng = engines[di]
            cx = coll if (coll is not None and n_dev == 1) else xch.for_rank(di)
            mine = batches[shards[di][0]:shards[di][1]]
            load = lambda fl: gio.NativeReads([files[i] for i in fl], threads=ingest_threads, packed=True)
            info = dict(names=[], lens=[], chunk_first=[], chunk_hap=[], rows0=[], n_reads=0, n_bases=0)
            with ThreadPoolExecutor(max_workers=1) as pre:
                fut = pre.submit(load, mine[0]) if mine else None
                held = []

                def bind_for(bi):
                    def bind(e):
                        nonlocal fut
                        reads = fut.result()
                        fut = pre.submit(load, mine[bi + 1]) if bi + 1 < len(mine) else None
                        for old in held:  # the previous batch's words: its copies finished with its match
                            old.close()
                        held.clear()
                        held.append(reads)
                        ch = [file_hap[i] for i in mine[bi]]
                        e.set_reads_packed(reads.words, reads.read_off, reads.chunk_first, ch)
                        info["names"].append(gio.NameTable(blob=reads.name_blob, off=reads.name_off))
                        info["lens"].append(reads.lengths().astype(np.uint32))
                        info["chunk_first"].append(reads.chunk_first.astype(np.int64) + info["n_reads"])
                        info["chunk_hap"].append(np.asarray(ch, np.uint8))
                        info["n_reads"] += reads.n_reads
                        info["n_bases"] += reads.total_bases
                    return bind

                def on_batch(e, b, base):
                    if detailed:
                        r0 = e.rows(0)
                        r0["read"] = r0["read"] + np.uint32(base)
                        info["rows0"].append(r0)
                iv, bases = eng.run_batches([bind_for(bi) for bi in range(len(mine))], contig_hap, min_read_len=min_read_len,
                                            allreduce_hist=cx.allreduce_hist, gather_forests=cx.gather_forests, on_batch=on_batch)
                for old in held:
                    old.close()
            info.update(kept=eng.rows(1, pinned=True), pairs=eng.pairs(pinned=True), iv=iv)
            results[di] = info
        except BaseException as ex:  # noqa: BLE001 -- reported by the caller; peers must not wait for this rank
            errors.append(ex)
            xch.abort()

    if n_dev == 1:
        worker(0)
    else:
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(n_dev)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    if errors:
        first = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)] or errors
        raise first[0]
    return results


def run_fused(k: int, hap_asm: Sequence[str], hap_reads: Sequence[Sequence[str]], outdir: str, device: int = 0,
              devices: Sequence[int] = None, batch_files: int = 0, batch_gbp: float = 12.0, detailed: bool = False,
              threads: int = 0, mrsfast_palindromes: bool = False):
    """hap_asm: two FASTA paths; hap_reads: two lists of chunk files in scatter order.  Builds the SUNK database on
    every GPU of `devices`, runs the reads (run_reads) and writes every file of the rules between jellyfish_count and
    slop_gaps under `outdir`.  Returns the first device's engine."""
    from concurrent.futures import ThreadPoolExecutor
    devices = list(devices) if devices else [device]
    contigs, contig_hap, fai = [], [], [[], []]
    for hap in range(2):
        for n, s in gio.read_fastx(hap_asm[hap]):
            contigs.append((n, s))
            contig_hap.append(hap)
            fai[hap].append((n, len(s)))
    files = list(hap_reads[0]) + list(hap_reads[1])
    file_hap = [0] * len(hap_reads[0]) + [1] * len(hap_reads[1])
    for p in files:
        if not os.path.exists(p):
            raise FileNotFoundError(p)

    def make(d):
        e = Engine(k, device=d)
        e.build_db(contigs)
        return e
    if len(devices) == 1:
        engines = [make(devices[0])]
    else:
        with ThreadPoolExecutor(max_workers=len(devices)) as ex:
            engines = list(ex.map(make, devices))
    results = run_reads(engines, contig_hap, files, file_hap, batch_files, batch_gbp, detailed, threads)
    eng = engines[0]
    gaps, nodata = eng.gaps(np.asarray([len(s) for _, s in contigs], dtype=np.uint32))
    write_outputs(k, eng, results, [len(s) for _, s in contigs], contig_hap, eng.contig_names, results[0]["iv"], gaps, nodata, outdir,
                  detailed=detailed, palindromes=mrsfast_palindromes)
    for e in engines[1:]:
        e.close()
    return eng


def write_outputs(k, eng, results, contig_len, contig_hap, names, iv, gaps, nodata, outdir, detailed=False, palindromes=False,
                  db_files=True):
    """every file of SURVEY Appendix C that the rules between jellyfish_count and slop_gaps produce (db_files=False:
    without the exports of the database itself -- db/jellyfish.db|.fa, mrsfast/kmer.loc and its per-contig copies
    breaks/*.loc -- which belong to the database build, not to a run over reads)"""
    d = lambda *p: os.path.join(outdir, *p)
    for sub in ("db", "mrsfast", "sunkpos", "breaks", "inter_outs", "bed_files", "final_out"):
        os.makedirs(d(sub), exist_ok=True)
    ctab = gio.NameTable(names)
    safe = lambda n: n.replace("#", "_")
    jobs = []  # (path, columns, options): formatted and written by a small pool at the end (the files are independent)

    def emit(path, cols, **kw):
        jobs.append((path, cols, kw))
    # ---- run-wide read table: devices in order, batches in order ----
    read_base, nreads = [], 0
    for r in results:
        read_base.append(nreads)
        nreads += r["n_reads"]
    rtab = gio.NameTable.concat([t for r in results for t in r["names"]])
    lens = np.concatenate([l for r in results for l in r["lens"]] + [np.zeros(0, np.uint32)])
    chunk_first = np.concatenate([cf[:-1] + read_base[i] for i, r in enumerate(results) for cf in r["chunk_first"]] + [np.array([nreads])])
    chunk_hap = np.concatenate([ch for r in results for ch in r["chunk_hap"]] + [np.zeros(0, np.uint8)])
    read_hap = np.zeros(nreads, np.uint8)
    for c in range(len(chunk_hap)):
        read_hap[int(chunk_first[c]):int(chunk_first[c + 1])] = chunk_hap[c]

    def cat_rows(key, cols):
        out = {}
        for c in cols:
            parts = []
            for i, r in enumerate(results):
                tabs = r[key] if isinstance(r[key], list) else [r[key]]
                for t in tabs:
                    parts.append(t[c] + np.uint32(read_base[i]) if (c == "read" and read_base[i]) else t[c])
            out[c] = parts[0] if len(parts) == 1 else (np.concatenate(parts) if parts else np.zeros(0, np.uint32))
        return out
    kept = cat_rows("kept", ("read", "pos", "contig", "start", "group"))
    pairs = cat_rows("pairs", ("read", "contig", "group"))
    sunk_cols = lambda t: [("name", t["read"], rtab), ("u32", t["pos"]), ("name", t["contig"], ctab), ("u32", t["start"]), ("u32", t["group"])]
    # ---- SUNK database files (defineSUNKs.smk:59-60, 126) ----
    from .engine import revcomp_kmers
    loc_sel, loc_c = None, {}
    if db_files:
        db = eng.db_export()
        loc_cols = [("name", db["contig"], ctab), ("u32", db["start"]), ("kmer", db["kmer"], k), ("u32", db["group"])]
        if palindromes and k % 2 == 0 and len(db["kmer"]):
            # mrsfast reports a k-mer that is its own reverse complement on both strands: two identical rows
            pal = db["kmer"] == revcomp_kmers(db["kmer"], k)
            loc_sel = np.repeat(np.arange(len(pal), dtype=np.uint64), 1 + pal.astype(np.int64))
        emit(d("db", "jellyfish.db"), [("kmer", db["kmer"], k)])
        emit(d("db", "jellyfish.fa"), [("kmer", db["kmer"], k, dict(prefix=b">", sep=b"\n")), ("kmer", db["kmer"], k)])
        emit(d("mrsfast", "kmer.loc"), loc_cols, sel=loc_sel)
    # ---- per-haplotype row files (combine_ont, combine_ont_nofilt, read_lengths) ----
    khap = read_hap[kept["read"]] if len(kept["read"]) else np.zeros(0, np.uint8)
    all_reads = np.arange(nreads, dtype=np.uint32)
    if detailed:
        rows0 = cat_rows("rows0", ("read", "pos", "contig", "start", "group"))
        r0hap = read_hap[rows0["read"]] if len(rows0["read"]) else np.zeros(0, np.uint8)
    for hap in range(2):
        emit(d("sunkpos", f"hap{hap + 1}.sunkpos"), sunk_cols(kept), sel=np.flatnonzero(khap == hap))
        emit(d("sunkpos", f"hap{hap + 1}.rlen"), [("name", all_reads, rtab), ("u32", lens)], sel=np.flatnonzero(read_hap == hap))
        if detailed:  # unfiltered rows for plot_detailed (viz_detailed.py:48)
            emit(d("sunkpos", f"hap{hap + 1}_detailed.sunkpos"), sunk_cols(rows0), sel=np.flatnonzero(r0hap == hap))
    gtab = eng.groups()
    bad = eng.bad_list()
    emit(d("sunkpos", "bad_sunks.txt"), [("name", gtab["contig"][bad], ctab, dict(sep=b":")), ("u32", gtab["group"][bad])])
    # ---- per-contig files: breaks/*.sunkpos|.loc (split_locs.py:5-22), inter_outs/*.tsv, bed_files/*.bed ----
    def by_contig(col, read=None, rank=None):
        """row indices per contig, in row order (rank given: by rank[read] first, stable).  Rows come in runs of one
        read on one contig, so the runs are sorted (a few per read), not the rows."""
        n = len(col)
        if n == 0:
            return {}
        brk = np.ones(n, bool)
        if read is not None:
            brk[1:] = (col[1:] != col[:-1]) | (read[1:] != read[:-1])
        else:
            brk[1:] = col[1:] != col[:-1]
        starts = np.flatnonzero(brk)
        lens = np.diff(np.append(starts, n))
        rc = col[starts].astype(np.int64)
        key = rc if rank is None else rc * np.int64(len(rank) + 1) + rank[read[starts]]
        order = np.argsort(key, kind="stable")
        starts, lens, rc = starts[order], lens[order], rc[order]
        ends = np.cumsum(lens)
        rows = np.repeat(starts - (ends - lens), lens) + np.arange(int(ends[-1]), dtype=np.int64)  # expand the sorted runs
        cs, first = np.unique(rc, return_index=True)
        lo = (ends - lens)[first]
        hi = np.append(lo[1:], ends[-1])
        return {int(c): rows[a:b] for c, a, b in zip(cs, lo, hi)}
    # `for rname, g in grouped`: groupby sorts read names (process-by-contig_lowmem_AR.py:135-136); rank of every read
    # name in that order (stable for equal names), pairs of a read stay in vertex order
    name_rank = np.empty(nreads, np.int64)
    if nreads:
        name_rank[np.argsort(rtab.as_bytes_array(), kind="stable")] = np.arange(nreads)
    kept_c = by_contig(kept["contig"], kept["read"])
    pair_c = by_contig(pairs["contig"], pairs["read"], name_rank)
    if db_files:
        loc_c = by_contig(db["contig"])
    iv_c = by_contig(iv["contig"])
    for c, idx in kept_c.items():
        stem = f"{safe(names[c])}_hap{contig_hap[c] + 1}"
        emit(d("breaks", stem + ".sunkpos"), sunk_cols(kept), sel=idx)
        if c in loc_c:
            sel = loc_c[c]
            if loc_sel is not None:
                sel = np.repeat(sel, 1 + (db["kmer"][sel] == revcomp_kmers(db["kmer"][sel], k)).astype(np.int64))
            emit(d("breaks", stem + ".loc"), loc_cols, sel=sel)
        if c in pair_c:
            pidx = pair_c[c]
            emit(d("inter_outs", stem + ".tsv"), [("u32", pairs["group"]), ("name", pairs["read"], rtab)], sel=pidx)
            emit(d("bed_files", stem + ".bed"), [("name", iv["contig"], ctab), ("u32", iv["start"]), ("u32", iv["end"])],
                                                                    sel=iv_c.get(c, np.zeros(0, np.uint64)))
        else:  # "no usable reads": contig name only, bed touched empty by the rule (tagONT.smk:190)
            _write_bytes(d("inter_outs", stem + ".tsv"), (names[c] + "\n").encode("latin-1"))
            _write_bytes(d("bed_files", stem + ".bed"), b"")
    clen = np.asarray(contig_len, dtype=np.uint32)
    chap = np.asarray(contig_hap, np.uint8)
    for hap in range(2):
        hn = hap + 1
        open(d("breaks", f"hap{hn}_splits_pos.done"), "a").close()  # touch() of the checkpoint (tagONT.smk:158)
        bed3 = lambda t: [("name", t["contig"], ctab), ("u32", t["start"]), ("u32", t["end"])]
        emit(d("final_out", f"hap{hn}.valid.bed"), bed3(iv), sel=np.flatnonzero(chap[iv["contig"]] == hap))
        gsel = np.flatnonzero(chap[gaps["contig"]] == hap)
        emit(d("final_out", f"hap{hn}.gaps.bed"), bed3(gaps), sel=gsel)
        nd = nodata[chap[nodata] == hap]
        emit(d("final_out", f"hap{hn}.nodata.bed"), [("name", nd, ctab), ("u32", np.zeros(len(nd), np.uint32)), ("u32", clen[nd])])
        s2, e2 = eng.slop(gaps["contig"][gsel], gaps["start"][gsel], gaps["end"][gsel], clen, 200000)  # bedtools slop -b 200000 (tagONT.smk:249)
        emit(d("final_out", f"hap{hn}.gaps.slop.bed"), [("name", gaps["contig"][gsel], ctab), ("i64", s2), ("i64", e2)])
    from concurrent.futures import ThreadPoolExecutor
    n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(8, n_cpu // 2))
    per_job = max(1, n_cpu // workers)
    jobs.sort(key=lambda j: -(len(j[2]["sel"]) if j[2].get("sel") is not None else len(j[1][0][1])))  # big files first
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(lambda j: _write_bytes(j[0], gio.format_rows(j[1], threads=per_job, **j[2])), jobs))


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
    ap.add_argument("--devices", default=None, help="comma-separated GPU indices: the chunk files are sharded over them")
    ap.add_argument("--batch-files", type=int, default=0, help="chunk files per batch (0 = by size, about --batch-gbp each)")
    ap.add_argument("--batch-gbp", type=float, default=12.0)
    ap.add_argument("--detailed", action="store_true", help="also write sunkpos/{hap}_detailed.sunkpos (plot_detailed: true)")
    ap.add_argument("--threads", type=int, default=0, help="ingest threads per GPU")
    ap.add_argument("--mrsfast-palindromes", action="store_true",
                    help="write a palindromic SUNK twice into kmer.loc, as mrsfast reports it on both strands (SURVEY A.2; unpinned)")
    a = ap.parse_args(argv)
    devs = [int(x) for x in a.devices.split(",")] if a.devices else [a.device]
    run_fused(a.k, [a.hap1_asm, a.hap2_asm], [a.hap1_reads, a.hap2_reads], a.outdir, devices=devs, batch_files=a.batch_files,
              batch_gbp=a.batch_gbp, detailed=a.detailed, threads=a.threads, mrsfast_palindromes=a.mrsfast_palindromes)


COMMANDS = {"kmerpos_annot3": (cmd_kmerpos_annot3, 4), "rlen": (cmd_rlen, 2), "diag_filter_v3": (cmd_diag_filter_v3, 2),
            "diag_filter_step2": (cmd_diag_filter_step2, 2), "badsunks_AR": (cmd_badsunks, 5),
            "process_by_contig": (cmd_process_by_contig, None), "get_gaps": (cmd_get_gaps, 5), "slop_gaps": (cmd_slop_gaps, 3),
            "covprob": (cmd_covprob, None), "split_locs": (cmd_split_locs, None), "fused": (cmd_fused, None),
            "split_ont": (cmd_split_ont, None), "combine_ont_nofilt": (cmd_combine_ont_nofilt, None)}


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
