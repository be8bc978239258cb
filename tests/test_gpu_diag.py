"""GPU parity of gvs_diag_filter against the reference ELFs diag_filter_v3 / diag_filter_step2
(outputs recorded in tests/golden): numerics of the interpolated median (Q8), distinct-group
band vote, Nim Table tie order incl. capacity growth (Q9), .fai membership (Q10), keep-all (Q11)."""
import numpy as np
import pytest

from conftest import load_golden, engine_from_case, run_match_chunks, format_rows
from gavisunk_b200 import io as gio

pytestmark = pytest.mark.gpu


def _fmt_best(best, read_names, contig_names, lo, hi):
    out = []
    for r, c, g, d in zip(best["read"], best["contig"], best["ngood"], best["dir"]):
        if lo <= r < hi:
            out.append(f"{read_names[r]}\t{contig_names[c]}\t{g}\t{'+' if d else '-'}\t{g}\n")
    return "".join(out)


@pytest.mark.parametrize("name", ["rand_k20", "rand_k16", "rand_k24", "rand_k31", "rand_k20_many", "ragged_k20", "ragged_k31", "ragged_k16"])
def test_diag_after_match(name):
    case = load_golden(name)
    eng, names = engine_from_case(case)
    fai1 = {l.split("\t")[0] for l in case["fai1"].splitlines()}
    fai2 = {l.split("\t")[0] for l in case["fai2"].splitlines()}
    contig_hap = [0 if n in fai1 else 1 if n in fai2 else 255 for n in names]
    chunks = case["chunks"]
    got = run_match_chunks(eng, names, [ch["reads"] for ch in chunks], [ch["hap"] - 1 for ch in chunks])
    eng.diag_filter(contig_hap)
    best, kept = eng.best(), eng.rows(1)
    cf = eng._chunk_first
    for i, ch in enumerate(chunks):
        assert got[i] == ch["sunkpos"]
        assert _fmt_best(best, eng._read_names, names, cf[i], cf[i + 1]) == ch["diag"]
        assert format_rows(kept, eng._read_names, names, cf[i], cf[i + 1]) == ch["diag2"]


def test_diag_cases_from_rows():
    from gavisunk_b200.engine import Engine
    for case in load_golden("diag_cases"):
        rows = gio.parse_sunkpos(case["sunkpos"])
        fai = [l.split("\t")[0] for l in case["fai"].splitlines()]
        names, cidx = [], {}
        for r in rows:
            if r[2] not in cidx:
                cidx[r[2]] = len(names)
                names.append(r[2])
        rnames, ridx, prev = [], [], None
        for r in rows:
            if r[0] != prev:
                rnames.append(r[0])
                prev = r[0]
            ridx.append(len(rnames) - 1)
        eng = Engine(20)
        eng.contig_names = names
        eng.set_reads_meta(np.full(len(rnames), 20000, np.uint32))
        eng.set_rows(0, ridx, [r[1] for r in rows], [cidx[r[2]] for r in rows], [r[3] for r in rows], [r[4] for r in rows])
        fs = set(fai)
        eng.diag_filter([0 if n in fs else 255 for n in names])
        best, kept = eng.best(), eng.rows(1)
        assert _fmt_best(best, rnames, names, 0, len(rnames)) == case["diag"], case["name"]
        assert format_rows(kept, rnames, names, 0, len(rnames)) == case["diag2"], case["name"]
