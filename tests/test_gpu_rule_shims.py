"""Edge cases of the per-rule shims (gavisunk_b200.cli) on hand-made files, against the oracle port of
the reference scripts: badsunks_AR.py (off-haplotype contigs, empty mode), process-by-contig
("contig name only" outputs, duplicate rows, reads missing from .rlen, bad SUNK removal), get_gaps.py
(missing / empty / overlapping BEDs, '#' in contig names), bedtools slop, covprob.py (errors)."""
import os

import pytest

import gavisunk_oracle as O
from gavisunk_b200 import cli, io as gio

pytestmark = pytest.mark.gpu


def _w(p, text):
    p.write_text(text)
    return str(p)


def _sunkpos(rows):
    return "".join(f"{r}\t{p}\t{c}\t{s}\t{g}\n" for r, p, c, s, g in rows)


def test_badsunks_off_haplotype_and_limits(tmp_path):
    # hap1 reads: groups on h1 contigs with counts 1, 4 (x3 groups = mode), 9, 30; an h2 contig with 30 and 3
    rows1 = []

    def add(rows, contig, group, n, tag):
        for i in range(n):
            rows.append((f"{tag}{i:03d}", 10 * i, contig, group + 3, group))

    for g, n in ((100, 1), (200, 4), (300, 4), (400, 4), (500, 9), (600, 30)):
        add(rows1, "h1#a", g, n, "r")
    add(rows1, "h2:b", 100, 30, "x")
    add(rows1, "h2:b", 900, 3, "x")
    rows2 = []
    for g, n in ((100, 2), (200, 2), (300, 1), (900, 12)):
        add(rows2, "h2:b", g, n, "q")
    add(rows2, "h1#a", 100, 50, "q")
    fai1 = _w(tmp_path / "h1.fai", "h1#a\t100000\t0\t60\t61\n")
    fai2 = _w(tmp_path / "h2.fai", "h2:b\t100000\t0\t60\t61\n")
    sp1, sp2 = _w(tmp_path / "1.sunkpos", _sunkpos(rows1)), _w(tmp_path / "2.sunkpos", _sunkpos(rows2))
    out = tmp_path / "bad.txt"
    assert cli.main(["badsunks_AR", fai1, fai2, sp1, sp2, str(out)]) == 0
    want = O.bad_sunks(rows1, {"h1#a"}, rows2, {"h2:b"})
    assert set(out.read_text().split()) == {f"{c}:{g}" for c, g in want}
    assert "h2:b:100" in out.read_text() and "h2:b:900" in out.read_text()  # off-hap > limit, on-hap > limit
    # no row on a contig of the haplotype's own .fai: the reference dies in .mode().values[0]
    assert cli.main(["badsunks_AR", fai2, fai2, sp2, _w(tmp_path / "3.sunkpos", _sunkpos([r for r in rows1 if r[2] == "h1#a"])),
                     str(tmp_path / "bad2.txt")]) == 1
    assert not (tmp_path / "bad2.txt").exists()


def _collinear_read(name, contig, ids, pos0=100, scale=1.0, rev=False):
    rows = []
    for i, g in enumerate(ids):
        p = pos0 + int((g - ids[0]) * scale)
        if rev:
            p = 200000 - p
        rows.append((name, p, contig, g + 5, g))
    return rows


def test_process_by_contig_edge_cases(tmp_path):
    ids = [1000, 3000, 7000, 12000, 20000, 26000]
    rows = []
    rows += _collinear_read("rB", "ctg#1", ids)                       # forward, all consistent
    rows += _collinear_read("rA", "ctg#1", ids[1:], rev=True)          # reverse strand
    rows += _collinear_read("rC", "ctg#1", ids[:4], scale=1.5)         # wrong spacing: no edge
    rows += _collinear_read("rD", "ctg#1", ids)                        # not in .rlen
    rows += _collinear_read("rE", "ctg#1", ids[2:])                    # too short
    rows += _collinear_read("rF", "ctg#1", [40000, 41000, 43000])      # second component
    rows += _collinear_read("rG", "ctg#1", [41000, 43000, 47000])
    rows += [rows[0], rows[3]]                                         # exact duplicates (drop_duplicates)
    rows += [("rB", 999999, "ctg#1", 5555, 5550)]                      # a bad SUNK group
    rlen = {"rA": 50000, "rB": 60000, "rC": 30000, "rE": 9999, "rF": 20000, "rG": 10000}
    sp = _w(tmp_path / "c.sunkpos", _sunkpos(rows))
    loc = _w(tmp_path / "c.loc", "ctg#1\t1005\tACGT\t1000\n")
    rl = _w(tmp_path / "a.rlen", "".join(f"{n}\t{l}\n" for n, l in rlen.items()))
    bad = _w(tmp_path / "bad.txt", "ctg#1:5550\nother:1000\nctg#1:notanumber\n")
    tsv, bed = tmp_path / "o.tsv", tmp_path / "o.bed"
    assert cli.main(["process_by_contig", loc, sp, rl, bad, str(tsv), str(bed), "--minlen", "5"]) == 0
    inter, regions = O.process_by_contig(rows, rlen, {("ctg#1", 5550)}, "ctg#1")
    assert tsv.read_text() == "".join(f"{g}\t{n}\n" for g, n in inter)
    assert bed.read_text() == "".join(f"{c}\t{s}\t{e}\n" for c, s, e in regions)
    assert len(regions) == 2 and "rD" not in tsv.read_text() and "rE" not in tsv.read_text()
    # nothing usable: contig name only, no bed (process-by-contig_lowmem_AR.py:92-94, 202-204)
    sp2 = _w(tmp_path / "d.sunkpos", _sunkpos(_collinear_read("rC", "ctg#1", ids[:4], scale=1.5) + [("rZ", 5, "ctg#1", 1005, 1000)]))
    tsv2, bed2 = tmp_path / "o2.tsv", tmp_path / "o2.bed"
    assert cli.main(["process_by_contig", loc, sp2, rl, bad, str(tsv2), str(bed2)]) == 0
    assert tsv2.read_text() == "ctg#1\n" and not bed2.exists()
    # unreadable input: non-zero exit, nothing written
    assert cli.main(["process_by_contig", loc, str(tmp_path / "missing"), rl, bad, str(tmp_path / "o3.tsv"), str(tmp_path / "o3.bed")]) == 1
    assert not (tmp_path / "o3.tsv").exists()


def test_get_gaps_and_slop(tmp_path):
    fai1 = _w(tmp_path / "h1.fai", "c#1\t900000\t0\t60\t61\nc2\t50000\t0\t60\t61\nc3\t70000\t0\t60\t61\n")
    fai2 = _w(tmp_path / "h2.fai", "d1\t800000\t0\t60\t61\nd2\t1000\t0\t60\t61\n")
    ind, outd = tmp_path / "beds", tmp_path / "out"
    os.makedirs(ind)
    os.makedirs(outd)
    beds = {"c#1": [(1000, 5000), (4000, 9000), (9000, 9500), (250000, 300000), (20000, 21000), (700000, 880000)],
            "c3": [], "d1": [(5, 10), (300000, 500000)]}
    _w(ind / "c_1_hap1.bed", "".join(f"c#1\t{s}\t{e}\n" for s, e in beds["c#1"]))  # '#' -> '_' in the file name only
    _w(ind / "c3_hap1.bed", "")
    _w(ind / "d1_hap2.bed", "".join(f"d1\t{s}\t{e}\n" for s, e in beds["d1"]))
    assert cli.main(["get_gaps", fai1, fai2, "s", str(ind) + "/", str(outd) + "/"]) == 0
    for hap, fai in ((1, fai1), (2, fai2)):
        gaps, nodata = O.get_gaps(gio.read_fai(fai), beds)
        assert (outd / f"hap{hap}.gaps.bed").read_text() == "".join(f"{c}\t{s}\t{e}\n" for c, s, e in gaps)
        assert (outd / f"hap{hap}.nodata.bed").read_text() == "".join(f"{c}\t{s}\t{e}\n" for c, s, e in nodata)
        assert cli.main(["slop_gaps", str(outd / f"hap{hap}.gaps.bed"), fai, str(outd / f"hap{hap}.slop.bed")]) == 0
        want = O.slop_gaps(gaps, dict(gio.read_fai(fai)))
        assert (outd / f"hap{hap}.slop.bed").read_text() == "".join(f"{c}\t{s}\t{e}\n" for c, s, e in want)
    assert (outd / "hap1.gaps.bed").read_text().startswith("c#1\t9500\t19999\n")
    assert (outd / "hap1.nodata.bed").read_text() == "c2\t0\t50000\nc3\t0\t70000\n"


def test_covprob_lookup_and_errors(tmp_path):
    k = 20
    loc_rows = []
    for c, ids in (("chr2", [500, 40500, 41000, 2600000]), ("chr10", [100, 300000, 300100, 900000])):
        for g in ids:
            for j in range(2):
                loc_rows.append((c, g + j, f"K{c}{g}{j}", g))
    loc_rows.append(("chr10", 77, "Kchr25000", 77))  # duplicate k-mer text: dropped before grouping
    loc = _w(tmp_path / "k.loc", "".join(f"{c}\t{s}\t{km}\t{g}\n" for c, s, km, g in loc_rows))
    rl_rows = [(f"r{i}", 1000 * (5 + 7 * i % 400) + i) for i in range(300)] + [("r1", 1000 * (5 + 7 % 400) + 1)]
    rl = _w(tmp_path / "a.rlen", "".join(f"{n}\t{l}\n" for n, l in rl_rows))
    fai = _w(tmp_path / "h.fai", "chr2\t3000000\t0\t60\t61\nchr10\t1000000\t0\t60\t61\n")
    gaps = [("chr10", 150, 299999), ("chr2", 600, 40499), ("chr10", 300200, 899999), ("chr2", 41001, 2599999)]
    bed = _w(tmp_path / "g.bed", "".join(f"{c}\t{s}\t{e}\n" for c, s, e in gaps))
    tsv = tmp_path / "o.tsv"
    assert cli.main(["covprob", "--bed", bed, "--locs", loc, "--rlen", rl, "--fai", fai, "--sunk-len", str(k), "--tsv", str(tsv)]) == 0
    table = O.covprob_table(rl_rows, 4000.0, k)
    want = O.covprob_gaps(gaps, loc_rows, table)
    lines = tsv.read_text().splitlines()
    assert lines[0].split("\t") == ["index", "Chromosome", "Start", "End", "type", "max_gap", "covprob"]
    got = {int(l.split("\t")[0]): l.split("\t")[1:] for l in lines[1:]}
    assert [int(l.split("\t")[0]) for l in lines[1:]] == [1, 3, 0, 2]  # chr2 before chr10 (natural order), file order inside
    for i, (c, s, e, mg, p) in enumerate(want):
        assert got[i][:5] == [c, str(s), str(e), "gap", str(mg)]
        assert float(got[i][5]) == pytest.approx(p, rel=1e-6, abs=1e-15)
    # a gap with no SUNK group within [start-2, end+2]: the reference dies on int(nan)
    bad_bed = _w(tmp_path / "g2.bed", "chr2\t100000\t200000\n")
    assert cli.main(["covprob", "--bed", bad_bed, "--locs", loc, "--rlen", rl, "--fai", fai, "--sunk-len", str(k), "--tsv", str(tmp_path / "o2.tsv")]) == 1
    assert not (tmp_path / "o2.tsv").exists()
    # max_gap >= 3500 kbp: KeyError in covprobsdict
    loc3 = _w(tmp_path / "k3.loc", "chr2\t10\tAAA\t10\nchr2\t3600000\tCCC\t3600000\n")
    bed3 = _w(tmp_path / "g3.bed", "chr2\t11\t3599999\n")
    assert cli.main(["covprob", "--bed", bed3, "--locs", loc3, "--rlen", rl, "--fai", fai, "--sunk-len", str(k), "--tsv", str(tmp_path / "o3.tsv")]) == 1


@pytest.mark.parametrize("name", ["pystages_a", "pystages_b", "pystages_c"])
def test_shims_against_reference_python_scripts(tmp_path, name):
    """every per-rule shim against the files the reference's OWN Python scripts wrote for the same inputs
    (tests/golden/pystages_*.json.gz, see tests/golden/make_golden_py.py): bad_sunks.txt as a set, breaks/,
    inter_outs/, bed_files/, gaps / nodata BEDs byte for byte, covprob within 1e-6"""
    from conftest import load_golden
    c = load_golden(name)
    for sub in ("sunkpos", "breaks", "inter_outs", "bed_files", "final_out"):
        os.makedirs(tmp_path / sub)
    P = lambda *a: str(tmp_path.joinpath(*a))
    _w(tmp_path / "kmer.loc", c["loc"])
    for h in ("1", "2"):
        _w(tmp_path / "sunkpos" / f"hap{h}.sunkpos", c["hap"][h]["sunkpos"])
        _w(tmp_path / "sunkpos" / f"hap{h}.rlen", c["hap"][h]["rlen"])
        _w(tmp_path / f"hap{h}.fai", c[f"fai{h}"])
    assert cli.main(["badsunks_AR", P("hap1.fai"), P("hap2.fai"), P("sunkpos", "hap1.sunkpos"), P("sunkpos", "hap2.sunkpos"),
                     P("sunkpos", "bad_sunks.txt")]) == 0
    assert sorted(open(P("sunkpos", "bad_sunks.txt")).read().split()) == c["bad_sunks"]
    for h in ("1", "2"):
        assert cli.main(["split_locs", "--ont-pos", P("sunkpos", f"hap{h}.sunkpos"), "--kmer-loc", P("kmer.loc"),
                         "--flag", P("breaks", f"hap{h}_splits_pos.done"), "--hap", f"hap{h}"]) == 0
    for fn, txt in c["breaks"].items():
        assert open(P("breaks", fn)).read() == txt, fn
    for stem, tsv in c["inter_outs"].items():
        h = stem[-1]
        assert cli.main(["process_by_contig", P("breaks", stem + ".loc"), P("breaks", stem + ".sunkpos"), P("sunkpos", f"hap{h}.rlen"),
                         P("sunkpos", "bad_sunks.txt"), P("inter_outs", stem + ".tsv"), P("bed_files", stem + ".bed")]) == 0
        assert open(P("inter_outs", stem + ".tsv")).read() == tsv, stem
        if c["bed_files"][stem] is None:
            assert not os.path.exists(P("bed_files", stem + ".bed"))
            open(P("bed_files", stem + ".bed"), "w").close()  # `touch {output.bed}`
        else:
            assert open(P("bed_files", stem + ".bed")).read() == c["bed_files"][stem], stem
    assert cli.main(["get_gaps", P("hap1.fai"), P("hap2.fai"), "sample", P("bed_files") + "/", P("final_out") + "/"]) == 0
    n_cov = 0
    for h in ("1", "2"):
        for f in ("gaps.bed", "nodata.bed"):
            assert open(P("final_out", f"hap{h}.{f}")).read() == c["final_out"][f"hap{h}.{f}"], f
        rc = cli.main(["covprob", "--bed", P("final_out", f"hap{h}.gaps.bed"), "--locs", P("kmer.loc"), "--rlen", P("sunkpos", f"hap{h}.rlen"),
                       "--fai", P(f"hap{h}.fai"), "--sunk-len", str(c["k"]), "--tsv", P("final_out", f"hap{h}.covprob.tsv")])
        want = c["final_out"][f"hap{h}.gaps.covprob.tsv"]
        if c["final_out"][f"hap{h}.covprob_rc"] != 0:
            assert rc == 1
            continue
        assert rc == 0
        got = open(P("final_out", f"hap{h}.covprob.tsv")).read().splitlines()
        wl = want.splitlines()
        assert got[0] == wl[0] and len(got) == len(wl)
        for a, b in zip(got[1:], wl[1:]):
            a, b = a.split("\t"), b.split("\t")
            assert a[:6] == b[:6]
            assert float(a[6]) == pytest.approx(float(b[6]), rel=1e-6, abs=1e-15)
            n_cov += 1
    assert n_cov >= 1
