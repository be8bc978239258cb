"""GPU parity of the post-match stages against the oracle's line-by-line restatement of
badsunks_AR.py, process-by-contig_lowmem_AR.py and get_gaps.py (the scripts themselves cannot be
imported: graph_tool / pyranges are absent; DESIGN.md "Oracle")."""
from collections import defaultdict

import numpy as np
import pytest

import gavisunk_oracle as O

pytestmark = pytest.mark.gpu


def _pipeline_case(seed, contig_lens, cov, n50, min_len, k=20, snp=2e-3, dup=0.02, nchunks=3):
    import torch
    from gavisunk_b200.engine import Engine
    from gavisunk_b200 import workload as W
    eng = Engine(k)
    wl = W.make_assembly(eng, contig_lens, snp_rate=snp, dup_frac=dup, seed=seed)
    W.build_db(eng, wl)
    W.add_reads(eng, wl, coverage=cov, n50=n50, sigma=0.6, len_min=200, len_max=400000, seed=seed + 1, nchunks=nchunks)
    W.bind_reads(eng, wl)
    iv = eng.run_all(wl.contig_hap, min_read_len=min_len)
    return eng, wl, iv


def _check_against_oracle(eng, wl, iv, min_len):
    names = wl.contig_names
    kept = eng.rows(1)
    off = wl.read_off.cpu().numpy()
    rname = lambda r: f"r{int(r):09d}"
    rows = [(rname(r), int(p), names[c], int(s), int(g)) for r, p, c, s, g in
            zip(kept["read"], kept["pos"], kept["contig"], kept["start"], kept["group"])]
    rlen = {rname(r): int(off[r + 1] - off[r]) for r in range(wl.n_reads)}
    nreads_hap = wl.n_reads // 2
    hap_of_read = lambda r: 0 if r < nreads_hap else 1
    nc = len(names) // 2
    hapc = [set(names[:nc]), set(names[nc:])]
    rows_h = [[], []]
    for r, row in zip(kept["read"], rows):
        rows_h[hap_of_read(int(r))].append(row)
    # ---- bad SUNKs (badsunks_AR.py) ----
    exp_bad = O.bad_sunks(rows_h[0], hapc[0], rows_h[1], hapc[1])
    db = eng.db_export()
    gidx_to = {}
    for c, g, gi in zip(db["contig"], db["group"], db["gidx"]):
        gidx_to[int(gi)] = (names[c], int(g))
    got_bad = {gidx_to[int(g)] for g in eng.bad_list()}
    assert got_bad == exp_bad
    assert len(exp_bad) > 0
    # ---- per-contig validation (process-by-contig_lowmem_AR.py) ----
    pairs = eng.pairs()
    got_inter = defaultdict(list)
    for r, c, g in zip(pairs["read"], pairs["contig"], pairs["group"]):
        got_inter[names[c]].append((int(g), rname(r)))
    got_bed = defaultdict(list)
    for c, s, e in zip(iv["contig"], iv["start"], iv["end"]):
        got_bed[names[c]].append((names[c], int(s), int(e)))
    by_contig = defaultdict(list)
    for row in rows:
        by_contig[row[2]].append(row)
    n_valid_reads = 0
    beds = {}
    for ctg in names:
        inter, bed = O.process_by_contig(by_contig.get(ctg, []), rlen, exp_bad, ctg, minlen=min_len) if ctg in by_contig else (None, None)
        assert got_inter.get(ctg, []) == (inter or []), ctg
        assert got_bed.get(ctg, []) == (bed or []), ctg
        if bed is not None:
            beds[ctg] = [(s, e) for _, s, e in bed]
        n_valid_reads += len({r for _, r in (inter or [])})
    assert n_valid_reads > 20
    assert sum(len(v) for v in got_bed.values()) > 0
    # ---- gaps (get_gaps.py) ----
    gaps, nodata = eng.gaps(wl.contig_len.astype(np.uint32))
    for hap in range(2):
        fai = [(names[c], int(wl.contig_len[c])) for c in range(hap * nc, (hap + 1) * nc)]
        eg, en = O.get_gaps(fai, beds)
        gg = [(names[c], int(s), int(e)) for c, s, e in zip(gaps["contig"], gaps["start"], gaps["end"]) if names[c] in hapc[hap]]
        gn = [(names[c], 0, int(wl.contig_len[c])) for c in nodata if names[c] in hapc[hap]]
        assert gg == eg
        assert gn == en
    return n_valid_reads


def test_pipeline_small_reads():
    eng, wl, iv = _pipeline_case(seed=21, contig_lens=[400000, 250000, 3000], cov=24.0, n50=12000, min_len=3000)
    _check_against_oracle(eng, wl, iv, 3000)


def test_pipeline_reference_minlen():
    """the reference's hard-coded 10 kb minimum read length (Q13) with longer reads"""
    eng, wl, iv = _pipeline_case(seed=33, contig_lens=[900000], cov=20.0, n50=30000, min_len=10000, nchunks=4)
    _check_against_oracle(eng, wl, iv, 10000)


def test_pipeline_k31_dense():
    eng, wl, iv = _pipeline_case(seed=5, contig_lens=[300000, 200000], cov=30.0, n50=15000, min_len=2000, k=31, snp=4e-3)
    _check_against_oracle(eng, wl, iv, 2000)


def test_validate_big_read_path():
    """reads with more rows than fit shared memory go through the global-scratch kernel"""
    eng, wl, iv = _pipeline_case(seed=77, contig_lens=[1500000], cov=8.0, n50=400000, min_len=10000, snp=3e-3, dup=0.0)
    kept = eng.rows(1)
    _, counts = np.unique(kept["read"], return_counts=True)
    assert counts.max() > 512
    _check_against_oracle(eng, wl, iv, 10000)


def test_pinned_result_buffers_match_pageable():
    """rows()/pairs() into the engine's page-locked pool (views reused by the next call) == fresh arrays"""
    eng, wl, iv = _pipeline_case(seed=21, contig_lens=[200000, 90000], cov=12.0, n50=12000, min_len=3000)
    for which in (0, 1):
        a, b = eng.rows(which), eng.rows(which, pinned=True)
        assert all(np.array_equal(a[c], b[c]) for c in a) and len(a["read"]) > 100
    a, b = eng.pairs(), eng.pairs(pinned=True)
    assert all(np.array_equal(a[c], b[c]) for c in a) and len(a["read"]) > 100
    b2 = eng.pairs(pinned=True)  # same buffers again
    assert b2["read"].ctypes.data == b["read"].ctypes.data


@pytest.mark.parametrize("n_rows", [6, 40, 300])
def test_validate_wide_position_span_uses_64_bit_predicate(n_rows):
    """.sunkpos rows fed through gvs_rows_set may span >= 2^28 positions: 10*dpos then overflows 32 bits and the
    ratio test must run in 64-bit arithmetic for the whole read, whichever thread owns which row (the choice is
    made on the span of ALL rows of the read, in each of the three row-count tiers)"""
    from gavisunk_b200.engine import Engine
    rng = np.random.default_rng(n_rows)
    step = (1 << 29) // n_rows * 2  # total span ~2^30
    pos = np.arange(n_rows, dtype=np.int64) * step
    start = 5000 + pos + rng.integers(-40, 40, n_rows)
    # a few rows far off the diagonal must stay unvalidated
    off = rng.choice(n_rows, max(1, n_rows // 6), replace=False)
    start[off] += 3 * step
    rows = [("rd", int(pos[i]), "c1", int(start[i]), int(start[i])) for i in range(n_rows)]
    inter, bed = O.process_by_contig(rows, {"rd": int(pos.max()) + 100}, set(), "c1", minlen=10000)
    assert inter
    eng = Engine(20)
    eng.contig_names = ["c1"]
    eng.set_reads_meta(np.array([int(pos.max()) + 100], np.uint32))
    eng.set_rows(1, np.zeros(n_rows, np.uint32), [r[1] for r in rows], np.zeros(n_rows, np.uint32), [r[3] for r in rows],
                 [r[4] for r in rows], n_contigs=1)
    eng.set_contigs([0])
    eng.validate(10000)
    p = eng.pairs()
    assert [(int(g), "rd") for g in p["group"]] == inter
    eng.components_local()
    iv = eng.intervals()
    assert [("c1", int(s), int(e)) for s, e in zip(iv["start"], iv["end"])] == bed


def test_rows_set_rejects_rows_outside_the_tables():
    """read / contig indices index device tables of later stages: out-of-range rows are an argument error"""
    from gavisunk_b200.engine import Engine, GavisunkError
    eng = Engine(20)
    eng.contig_names = ["c1", "c2"]
    eng.set_reads_meta(np.array([20000, 20000], np.uint32))
    with pytest.raises(GavisunkError, match="read index"):
        eng.set_rows(1, [0, 2], [1, 2], [0, 0], [10, 20], [10, 20], n_contigs=2)
    with pytest.raises(GavisunkError, match="contig"):
        eng.set_rows(1, [0, 1], [1, 2], [0, 2], [10, 20], [10, 20], n_contigs=2)
    eng.set_rows(1, [0, 1], [1, 2], [0, 1], [10, 20], [10, 20], n_contigs=2)


def _rows_case(rng, m, kind):
    """rows of ONE read on one contig (read pos, assembly start, ID) shaped to hit the branches of k_validate: reads that
    follow one strand (positions monotone along the starts -> the first-partner search), with jitter small or large
    against short distances, equal positions / equal starts, outliers off the diagonal, repeated IDs"""
    gaps = rng.integers(15, 3000, m)
    start = 10_000 + np.cumsum(gaps)
    if kind in ("fwd", "fwd_jitter", "fwd_ties", "fwd_outlier", "fwd_dupid", "sparse_links"):
        pos = 500 + (start - start[0]) + rng.integers(-3, 4, m).cumsum()
    else:
        pos = 5_000_000 - (start - start[0]) + rng.integers(-3, 4, m).cumsum()
    if kind.endswith("jitter"):
        pos = pos + rng.integers(-40, 41, m)           # breaks short-range pairs, may break monotonicity
    if kind == "fwd_ties":
        j = rng.choice(m - 1, max(1, m // 8), replace=False)
        pos[j + 1] = pos[j]                            # equal positions, different starts: never a passing pair
    if kind == "fwd_outlier":
        j = rng.choice(m, max(1, m // 10), replace=False)
        pos[j] += rng.integers(20_000, 2_000_000, len(j)) * rng.choice([-1, 1], len(j))
    if kind == "sparse_links":
        pos = 500 + ((start - start[0]) * 1.3).astype(np.int64)   # ratio 1.3: no pair passes ...
        k = rng.choice(m, min(m, max(3, m // 3)), replace=False)
        pos[k] = 500 + (start[k] - start[0])                        # ... except among these rows
    pos = np.maximum(pos, 0)
    ids = start.copy()
    if kind == "fwd_dupid" and m > 3:
        j = rng.choice(m - 2, max(1, m // 12), replace=False)
        ids[j + 2] = ids[j]                            # the same group again two rows later (Q5, multipos clean-up)
    order = np.argsort(pos, kind="stable")             # a read's rows come in increasing read position
    return [("rd", int(pos[i]), "c1", int(start[i]), int(ids[i])) for i in order]


@pytest.mark.parametrize("kind", ["fwd", "rev", "fwd_jitter", "rev_jitter", "fwd_ties", "fwd_outlier", "fwd_dupid", "sparse_links"])
def test_validate_shaped_reads_against_oracle(kind):
    """per-read validation of hand-shaped reads in all three row-count tiers against the oracle's restatement of
    process-by-contig_lowmem_AR.py:136-198 (pairs in vertex order, then the contig's intervals)"""
    from gavisunk_b200.engine import Engine
    rng = np.random.default_rng(sum(map(ord, kind)))
    sizes = [2, 3, 5, 17, 33, 64, 65, 100, 257, 512, 513, 700]
    reads, all_rows, rlen = [], [], {}
    for ri, m in enumerate(sizes * 3):
        rows = _rows_case(rng, m, kind)
        name = f"r{ri:04d}"
        # every read on its own stretch of the contig, so that the contig-wide components keep them apart
        shift = ri * 20_000_000
        rows = [(name, p, c, s + shift, g + shift) for _, p, c, s, g in rows]
        all_rows += rows
        rlen[name] = max(p for _, p, _, _, _ in rows) + 10_000
    inter, bed = O.process_by_contig(all_rows, rlen, set(), "c1", minlen=10000)
    eng = Engine(20)
    eng.contig_names = ["c1"]
    names = sorted(rlen)
    ridx = {n: i for i, n in enumerate(names)}
    eng.set_reads_meta(np.array([min(rlen[n], 0xFFFFFFFF) for n in names], np.uint32))
    eng.set_rows(1, [ridx[r[0]] for r in all_rows], [r[1] for r in all_rows], np.zeros(len(all_rows), np.uint32),
                 [r[3] for r in all_rows], [r[4] for r in all_rows], n_contigs=1)
    eng.set_contigs([0])
    eng.validate(10000)
    p = eng.pairs()
    got = [(int(g), names[int(r)]) for g, r in zip(p["group"], p["read"])]
    assert got == (inter or []), kind
    eng.components_local()
    iv = eng.intervals()
    assert [("c1", int(s), int(e)) for s, e in zip(iv["start"], iv["end"])] == (bed or []), kind
    assert len(got) > 100
