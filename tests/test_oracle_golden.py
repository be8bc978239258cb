"""Oracle (oracle/gavisunk_oracle.py) pinned against the outputs of the reference's own ELF
binaries recorded in tests/golden/*.json.gz (made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import gavisunk_oracle as O
from gavisunk_b200 import io as gio
from conftest import load_golden


def _parse_db_loc(case):
    db = [O.encode(l) for l in case["db"].splitlines() if l]
    loc = []
    for l in case["loc"].splitlines():
        c, s, km, g = l.split("\t")
        loc.append((c, int(s), O.encode(km), int(g)))
    return db, loc


@pytest.mark.parametrize("name", ["kat_b1", "kat_bytes", "kat_k32", "kat_k32b"])
def test_match_kats(name):
    case = load_golden(name)
    db, loc = _parse_db_loc(case)
    reads = gio.read_fastx(case["reads"].encode("latin-1"))
    rows = O.match_chunk(reads, db, loc, case["k"])
    assert gio.format_sunkpos(rows) == case["out"]


def test_match_q7_duplicate_loc_rows_and_orphan_db_kmer():
    """SURVEY Q7 through the executable: a k-mer listed twice in kmer.loc reports its LAST row
    (kmerpos_annot3.nim:68), a .loc k-mer that is not in the db set never matches (nim:89), a read that hits a db
    k-mer without a .loc row ends the run with a KeyError (nim:90: exit code 1 after the rows printed so far)"""
    case = load_golden("kat_q7")
    db, loc = _parse_db_loc(case)
    rows = O.match_chunk(gio.read_fastx(case["reads"].encode("latin-1")), db, loc, case["k"])
    assert gio.format_sunkpos(rows) == case["out"] and case["rc"] == 0
    assert "c2\t900\t880" in case["out"] and "c1\t100\t100" not in case["out"] and "\t500\t500" not in case["out"]
    assert case["rc_keyerror"] != 0
    with pytest.raises(KeyError):
        O.match_chunk(gio.read_fastx(case["reads_keyerror"].encode("latin-1")), db, loc, case["k"])


def test_b1_expected_rows():
    # SURVEY Appendix B.1 literal expectation
    case = load_golden("kat_b1")
    got = [tuple(l.split("\t")) for l in case["out"].splitlines()]
    assert got == [("r1", "2", "c1", "100", "100"), ("r1", "7", "c1", "200", "200"), ("r2", "5", "c1", "100", "100"),
                   ("r3", "2", "c1", "200", "200"), ("r3", "7", "c1", "101", "100"), ("r4", "7", "c1", "200", "200"),
                   ("r4", "14", "c1", "100", "100")]


def test_rlen_b8():
    case = load_golden("rlen_b8")
    reads = gio.read_fastx(case["reads"].encode("latin-1"))
    assert "".join(f"{n}\t{len(s)}\n" for n, s in reads) == case["rlen"]


@pytest.mark.parametrize("name", ["rand_k20", "rand_k16", "rand_k24", "rand_k31", "rand_k20_many", "ragged_k20", "ragged_k31", "ragged_k16"])
def test_match_and_diag_random(name):
    case = load_golden(name)
    db, loc = _parse_db_loc(case)
    for ch in case["chunks"]:
        reads = gio.read_fastx(ch["reads"].encode("latin-1"))
        assert "".join(f"{n}\t{len(s)}\n" for n, s in reads) == ch["rlen"]
        rows = O.match_chunk(reads, db, loc, case["k"])
        assert gio.format_sunkpos(rows) == ch["sunkpos"]
        fai = case["fai1"] if ch["hap"] == 1 else case["fai2"]
        contigs = {l.split("\t")[0] for l in fai.splitlines()}
        diag = O.diag_filter_v3(rows, contigs)
        assert "".join("\t".join(str(x) for x in d) + "\n" for d in diag) == ch["diag"]
        kept = O.diag_filter_step2(rows, diag)
        assert gio.format_sunkpos(kept) == ch["diag2"]


def test_match_every_sunk_len():
    """ksweep: every SUNK_len besides 4 / 16 / 20 / 24 / 31 (2 ... 30 and the reference's broken k = 32, Q2)"""
    cases = load_golden("ksweep")
    assert sorted(c["k"] for c in cases) == [2, 3, 5, 6, 7, 8, 9, 11, 12, 13, 15, 17, 19, 21, 23, 25, 27, 28, 29, 30, 32]
    for case in cases:
        db, loc = _parse_db_loc(case)
        n_rows = 0
        for ch in case["chunks"]:
            reads = gio.read_fastx(ch["reads"].encode("latin-1"))
            assert "".join(f"{n}\t{len(s)}\n" for n, s in reads) == ch["rlen"]
            rows = O.match_chunk(reads, db, loc, case["k"])
            assert gio.format_sunkpos(rows) == ch["sunkpos"], case["k"]
            n_rows += len(rows)
        assert n_rows > 0, case["k"]


def test_diag_cases():
    for case in load_golden("diag_cases"):
        rows = gio.parse_sunkpos(case["sunkpos"])
        contigs = {l.split("\t")[0] for l in case["fai"].splitlines()}
        diag = O.diag_filter_v3(rows, contigs)
        got = "".join("\t".join(str(x) for x in d) + "\n" for d in diag)
        assert got == case["diag"], case["name"]
        kept = O.diag_filter_step2(rows, diag)
        assert gio.format_sunkpos(kept) == case["diag2"], case["name"]
