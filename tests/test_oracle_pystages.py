"""The oracle's restatement of the reference's PYTHON stages against outputs of the reference's own scripts
(tests/golden/pystages_*.json.gz, made by tests/golden/make_golden_py.py: badsunks_AR.py, split_locs.py,
process-by-contig_lowmem_AR.py, get_gaps.py, covprob.py run unmodified on sunkpos rows from the reference
executables, with graph-tool / pyranges replaced by the minimal stand-ins in tests/golden/refpy_stubs)."""
import pytest

import gavisunk_oracle as O
from conftest import load_golden
from gavisunk_b200 import io as gio

CASES = ["pystages_a", "pystages_b", "pystages_c"]


def _fai(txt):
    return [(l.split("\t")[0], int(l.split("\t")[1])) for l in txt.splitlines()]


def _rlen(txt):
    return {l.split("\t")[0]: int(l.split("\t")[1]) for l in txt.splitlines()}


@pytest.mark.parametrize("name", CASES)
def test_bad_sunks(name):
    c = load_golden(name)
    rows = [gio.parse_sunkpos(c["hap"][h]["sunkpos"]) for h in ("1", "2")]
    got = O.bad_sunks(rows[0], {n for n, _ in _fai(c["fai1"])}, rows[1], {n for n, _ in _fai(c["fai2"])})
    assert sorted(f"{a}:{b}" for a, b in got) == c["bad_sunks"]
    assert len(got) > 50


@pytest.mark.parametrize("name", CASES)
def test_split_and_process_by_contig(name):
    c = load_golden(name)
    bad = {(b.rsplit(":", 1)[0], int(b.rsplit(":", 1)[1])) for b in c["bad_sunks"]}
    n_iv = 0
    for h in ("1", "2"):
        rows = gio.parse_sunkpos(c["hap"][h]["sunkpos"])
        rlen = _rlen(c["hap"][h]["rlen"])
        by_c = {}
        for r in rows:
            by_c.setdefault(r[2], []).append(r)
        for ctg, rr in by_c.items():
            stem = f"{ctg.replace('#', '_')}_hap{h}"
            assert c["breaks"][stem + ".sunkpos"] == gio.format_sunkpos(rr)           # split_locs.py
            assert c["breaks"][stem + ".loc"] == "".join(l + "\n" for l in c["loc"].splitlines() if l.split("\t")[0] == ctg)
            inter, bed = O.process_by_contig(rr, rlen, bad, ctg)
            want_tsv = "".join(f"{g}\t{n}\n" for g, n in inter) if inter else ctg + "\n"
            assert c["inter_outs"][stem] == want_tsv, stem
            want_bed = "".join(f"{a}\t{s}\t{e}\n" for a, s, e in bed) if bed is not None else None
            assert c["bed_files"][stem] == want_bed, stem
            n_iv += len(bed or [])
    assert n_iv >= 5


@pytest.mark.parametrize("name", CASES)
def test_get_gaps_and_covprob(name):
    c = load_golden(name)
    k = c["k"]
    beds = {}
    for stem, txt in c["bed_files"].items():
        ctg = stem.rsplit("_hap", 1)[0]
        beds[ctg] = [(int(l.split("\t")[1]), int(l.split("\t")[2])) for l in (txt or "").splitlines()]
    loc = [(p[0], int(p[1]), p[2], int(p[3])) for p in (l.split("\t") for l in c["loc"].splitlines())]
    n_rows = 0
    for h in ("1", "2"):
        fai = _fai(c[f"fai{h}"])
        gaps, nodata = O.get_gaps(fai, beds)
        fmt = lambda rows: "".join(f"{a}\t{s}\t{e}\n" for a, s, e in rows)
        assert c["final_out"][f"hap{h}.gaps.bed"] == fmt(gaps)
        assert c["final_out"][f"hap{h}.nodata.bed"] == fmt(nodata)
        tsv = c["final_out"][f"hap{h}.gaps.covprob.tsv"]
        rl = [(l.split("\t")[0], int(l.split("\t")[1])) for l in c["hap"][h]["rlen"].splitlines()]
        table = O.covprob_table(rl, sum(l for _, l in fai) / 1000, k)
        if c["final_out"][f"hap{h}.covprob_rc"] != 0:
            with pytest.raises(Exception):
                O.covprob_gaps(gaps, loc, table)
            continue
        want = O.covprob_gaps(gaps, loc, table)
        lines = tsv.splitlines()
        assert lines[0].split("\t") == ["index", "Chromosome", "Start", "End", "type", "max_gap", "covprob"]
        got = {int(l.split("\t")[0]): l.split("\t")[1:] for l in lines[1:]}
        assert len(got) == len(want)
        for i, (ctg, s, e, mg, p) in enumerate(want):
            assert got[i][:5] == [ctg, str(s), str(e), "gap", str(mg)]
            assert float(got[i][5]) == pytest.approx(p, rel=1e-6, abs=1e-15)
            n_rows += 1
    assert n_rows >= 1
