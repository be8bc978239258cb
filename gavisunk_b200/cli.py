"""Command-line shims with the reference's argv on top of the GPU engine.

  python -m gavisunk_b200.cli kmerpos_annot3 <reads.fa|fq[.gz]> <db.txt> <kmer.loc> <out.sunkpos>
  python -m gavisunk_b200.cli rlen <reads> <out.rlen>
  python -m gavisunk_b200.cli diag_filter_v3 <sunkpos> <asm.fai>            (stdout)
  python -m gavisunk_b200.cli diag_filter_step2 <sunkpos[.gz]> <diag[.gz]>  (stdout)
  python -m gavisunk_b200.cli fused --sample S --k K --hap1-asm .. --hap2-asm .. --hap1-reads f.. --hap2-reads f.. --outdir D

The first four replace workflow/scripts/{kmerpos_annot3,rlen,diag_filter_v3,diag_filter_step2}
(workflow/rules/tagONT.smk:36,73,92,110) one to one; `fused` replaces the rules from jellyfish_count
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
    names, seq, off = gio.pack_reads(gio.read_fastx(reads_p))
    eng.set_reads(seq, off)
    eng.match()
    _atomic_write(out_p, _fmt_rows(eng.rows(0), names, eng.contig_names))


def cmd_rlen(argv):
    reads_p, out_p = argv
    _atomic_write(out_p, "".join(f"{n}\t{len(s)}\n" for n, s in gio.read_fastx(reads_p)))


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
    rnames, parts, lens, chunk_first, chunk_hap = [], [], [], [0], []
    for hap in range(2):
        for fp in hap_reads[hap]:
            for n, s in gio.read_fastx(fp):
                rnames.append(n)
                parts.append(s)
                lens.append(len(s))
            chunk_first.append(len(rnames))
            chunk_hap.append(hap)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8)
    eng.set_reads(seq, off, chunk_first, chunk_hap)
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
            "diag_filter_step2": (cmd_diag_filter_step2, 2), "fused": (cmd_fused, None)}


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
