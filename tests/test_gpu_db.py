"""GPU SUNK database build (gvs_db_build) and device-generated synthetic workloads against the
oracle's restatement of defineSUNKs.smk, plus the README known answer when the bundled
assemblies are available."""
import os

import numpy as np
import pytest

import gavisunk_oracle as O
from gavisunk_b200 import io as gio

pytestmark = pytest.mark.gpu


def _contigs_from_workload(wl):
    import torch
    asm = wl.asm.cpu().numpy()
    off = wl.contig_off.cpu().numpy()
    return [(n, asm[off[i]:off[i + 1]].tobytes()) for i, n in enumerate(wl.contig_names)]


def _check_db(eng, contigs, k):
    exp = O.build_sunk_db(contigs, k)
    got = eng.db_export()
    assert len(got["kmer"]) == len(exp["kmer"])
    assert np.array_equal(got["contig"].astype(np.int64), exp["contig"].astype(np.int64))
    assert np.array_equal(got["start"].astype(np.int64), exp["start"])
    assert np.array_equal(got["kmer"], exp["kmer"])
    assert np.array_equal(got["group"].astype(np.int64), exp["group"])
    # dense group index: consecutive, in (contig, group) order
    key = exp["contig"].astype(np.int64) << 32 | exp["group"]
    _, inv = np.unique(key, return_inverse=True)
    assert np.array_equal(got["gidx"].astype(np.int64), inv)
    return exp


@pytest.mark.parametrize("k", [16, 20, 24, 31])
def test_db_build_synthetic(k):
    from gavisunk_b200.engine import Engine
    from gavisunk_b200 import workload as W
    eng = Engine(k)
    wl = W.make_assembly(eng, [70000, 41000, 333], snp_rate=2e-3, dup_frac=0.03, seed=7 + k)
    W.build_db(eng, wl)
    contigs = _contigs_from_workload(wl)
    assert any(b"N" in s for _, s in contigs)
    exp = _check_db(eng, contigs, k)
    assert len(exp["kmer"]) > 1000


def test_db_build_host_edge_cases():
    from gavisunk_b200.engine import Engine
    eng = Engine(4)
    contigs = [("a", b"ACGTACGTTT"), ("b", b"nnAAACnGTTT"), ("c", b"AC"), ("d", b""), ("e", b"acgtTGCAAGGCTTAACCGGTTAGCATCGA" * 3)]
    eng.build_db(contigs)
    _check_db(eng, contigs, 4)
    eng5 = Engine(5)
    contigs = [("x", b"ACGTN" * 7 + b"GATTACAGATTACCA"), ("y", b"TTTTTTTTTTTTTTTTTTTTTGATTACA")]
    eng5.build_db(contigs)
    _check_db(eng5, contigs, 5)


def _hg02723_contigs():
    from conftest import load_golden
    asm = load_golden("hg02723_asm")
    return gio.read_fastx(asm["h1"].encode()) + gio.read_fastx(asm["h2"].encode())


def test_db_build_readme_known_answer():
    """README.md:36-37: expected figure AMY_HG02723_hap1_AMY_h1_84861_524275 = gap AMY_h1 284861-324275
    +-200 kb (tagONT.smk:249) <=> 284861 and 324276 are adjacent SUNK-group starts on AMY_h1.
    Inputs: the reference's bundled HG02723 assemblies (tests/golden/hg02723_asm.json.gz)."""
    from gavisunk_b200.engine import Engine
    contigs = _hg02723_contigs()
    eng = Engine(20)
    eng.build_db(contigs)
    db = eng.db_export()
    names = [n for n, _ in contigs]
    ci = names.index("AMY_h1")
    groups = np.unique(db["group"][db["contig"] == ci])
    i = int(np.searchsorted(groups, 284861))
    assert groups[i] == 284861 and groups[i + 1] == 324276
    assert np.diff(groups).max() == 324276 - 284861
    per = {n: (int((db["contig"] == j).sum()), len(np.unique(db["group"][db["contig"] == j]))) for j, n in enumerate(names)}
    assert per == {"AMY_h1": (4102, 205), "AMY_orphan_h1": (216, 12), "nonuniq_kmers": (5384, 246), "AMY_h2": (3959, 196)}
    _check_db(eng, contigs, 20)


@pytest.mark.parametrize("k", [16, 24, 31])
def test_db_build_hg02723_k_sweep(k):
    """SURVEY A.2 [probe]: AMY_h1 SUNKs/groups for k = 16 / 24 / 31"""
    from gavisunk_b200.engine import Engine
    contigs = _hg02723_contigs()
    eng = Engine(k)
    eng.build_db(contigs)
    db = eng.db_export()
    ci = [n for n, _ in contigs].index("AMY_h1")
    sel = db["contig"] == ci
    exp = {16: (3253, 206), 24: (6660, 456), 31: (11395, 448)}[k]
    assert (int(sel.sum()), len(np.unique(db["group"][sel]))) == exp


def test_match_on_device_generated_reads():
    """device-generated assembly + reads, GPU db build + match vs the oracle on the same bytes"""
    import torch
    from gavisunk_b200.engine import Engine
    from gavisunk_b200 import workload as W
    k = 20
    eng = Engine(k)
    wl = W.make_assembly(eng, [150000, 90000], snp_rate=2e-3, dup_frac=0.02, seed=11)
    W.build_db(eng, wl)
    W.add_reads(eng, wl, coverage=6.0, n50=9000, sigma=0.7, len_min=15, len_max=60000, seed=5, nchunks=3)
    W.bind_reads(eng, wl)
    n = eng.match()
    rows = eng.rows(0)
    contigs = _contigs_from_workload(wl)
    db = O.build_sunk_db(contigs, k)
    names = [c for c, _ in contigs]
    loc = [(names[c], int(s), int(km), int(g)) for c, s, km, g in zip(db["contig"], db["start"], db["kmer"], db["group"])]
    reads = wl.reads.cpu().numpy()
    off = wl.read_off.cpu().numpy()
    assert off[-1] == wl.total_bases
    exp = []
    cf = wl.chunk_first
    for c in range(len(cf) - 1):
        chunk = [(int(r), reads[off[r]:off[r + 1]].tobytes()) for r in range(int(cf[c]), int(cf[c + 1]))]
        exp += O.match_chunk(chunk, db["kmer"], loc, k)
    got = list(zip(rows["read"].tolist(), rows["pos"].tolist(), [names[c] for c in rows["contig"]], rows["start"].tolist(),
                   rows["group"].tolist()))
    assert n == len(exp) and n > 500
    assert got == exp
    # reads carry both strands and N runs
    assert any(b"N" in reads[off[r]:off[r + 1]].tobytes() for r in range(wl.n_reads)) or wl.n_reads < 1000
