"""BASELINE.json config 4 at FULL per-GPU size: 3.1 Gbp x2 diploid assembly in 23 contigs per haplotype
(1.2e8 SUNKs, whole-genome two-level filter with a saturated presence filter), one GPU's shard of the 30x
ultra-long reads (3.75x = 11.5 Gbp).  The reference executables cannot hold this database (33 GB and minutes
of text parsing per process), so the checks are the size-independent ones -- properties of the reference
algorithm, idempotence, equality of the pipelined host-copy path -- plus the oracle's match on a sample of
reads against the exported 1.2e8-SUNK database (membership vectorised in numpy, row rules by the oracle)."""
import numpy as np
import pytest

from test_gpu_fullsize import check_properties

pytestmark = pytest.mark.gpu

K = 20


@pytest.fixture(scope="module")
def full_h():
    import torch
    from gavisunk_b200 import workload as W
    from gavisunk_b200.engine import Engine
    eng = Engine(K)
    wl = W.make_assembly(eng, W.human_contigs(3100.0), snp_rate=1e-3, dup_frac=0.01, seed=1001, name="h3100")
    W.build_db(eng, wl)
    W.add_reads(eng, wl, coverage=3.75, n50=100000.0, sigma=0.8, len_min=1000, len_max=1000000, seed=2001, nchunks=10)
    W.bind_reads(eng, wl)
    iv = eng.run_all(wl.contig_hap, min_read_len=10000)
    gaps, nodata = eng.gaps(wl.contig_len.astype(np.uint32))
    res = dict(rows=eng.rows(0), kept=eng.rows(1), best=eng.best(), pairs=eng.pairs(), bad=eng.bad_list(), iv=iv, gaps=gaps,
               nodata=nodata, off=wl.read_off.cpu().numpy().astype(np.int64))
    yield eng, wl, res
    eng.close()
    del wl
    torch.cuda.empty_cache()


def test_sizes_are_config4(full_h):
    eng, wl, res = full_h
    n_sunks, n_groups = eng.db_size()
    assert n_sunks > 100_000_000 and n_groups > 5_000_000 and len(wl.contig_names) == 46
    assert wl.total_bases > 11_000_000_000 and len(res["rows"]["read"]) > 8_000_000
    assert len(res["iv"]["start"]) > 10_000 and len(res["gaps"]["start"]) > 10_000


def test_sample_of_reads_against_oracle_config4(full_h):
    """sunkpos rows of the first reads of one chunk per haplotype, produced by the LARGE-database probe
    instantiation on the whole-genome table, against oracle.match_chunk (kmerpos_annot3.nim:81-97).  The oracle
    gets the database restricted to the SUNKs the sample can hit at all (found with its own window encoder and a
    sorted-array membership over all 1.2e8 exported k-mers) -- the restriction cannot change its output."""
    import gavisunk_oracle as O
    eng, wl, res = full_h
    db = eng.db_export()
    order = np.argsort(db["kmer"], kind="stable")
    keys = db["kmer"][order]
    assert np.all(keys[1:] != keys[:-1])
    off = res["off"]
    cf = wl.chunk_first.astype(np.int64)
    rows = res["rows"]
    n_checked = 0
    for hap in (0, 1):
        chunk = int(np.nonzero(np.asarray(wl.chunk_hap) == hap)[0][1])
        r0 = int(cf[chunk])
        r1 = r0 + 60
        seq = wl.reads[int(off[r0]):int(off[r1])].cpu().numpy()
        reads = [(i, seq[off[i] - off[r0]:off[i + 1] - off[r0]].tobytes()) for i in range(r0, r1)]
        hit_rows = []
        for _, s in reads:
            canon = O.canonical_windows(O.codes_of(s), K)
            idx = np.minimum(np.searchsorted(keys, canon), len(keys) - 1)
            hit_rows.append(order[idx[keys[idx] == canon]])
        sub = np.unique(np.concatenate(hit_rows))
        loc = [(int(c), int(s), int(k), int(g)) for c, s, k, g in
               zip(db["contig"][sub], db["start"][sub], db["kmer"][sub], db["group"][sub])]
        want = O.match_chunk(reads, db["kmer"][sub], loc, K)
        lo, hi = np.searchsorted(rows["read"], [r0, r1])
        got = list(zip(rows["read"][lo:hi].tolist(), rows["pos"][lo:hi].tolist(), rows["contig"][lo:hi].tolist(),
                       rows["start"][lo:hi].tolist(), rows["group"][lo:hi].tolist()))
        # the chunk starts at r0, so the prevLoc carry of the full batch restarts exactly there
        assert got == want
        n_checked += len(want)
    assert n_checked > 3000


def test_properties_config4(full_h):
    check_properties(*full_h)


def test_idempotent_and_pipelined_config4(full_h):
    import torch
    from gavisunk_b200 import workload as W
    eng, wl, res = full_h
    h = torch.empty(wl.total_bases + 64, dtype=torch.uint8, pin_memory=True)
    h[:wl.total_bases].copy_(wl.reads[:wl.total_bases])
    torch.cuda.synchronize()
    eng.set_reads(h.numpy()[:wl.total_bases], res["off"].astype(np.uint64), wl.chunk_first, wl.chunk_hap)
    iv = eng.run_all(wl.contig_hap, min_read_len=10000)
    for c in ("read", "pos", "contig", "start", "group"):
        assert np.array_equal(eng.rows(0)[c], res["rows"][c]), c
        assert np.array_equal(eng.rows(1)[c], res["kept"][c]), c
    for c in ("read", "group", "gidx"):
        assert np.array_equal(eng.pairs()[c], res["pairs"][c]), c
    for c in ("contig", "start", "end"):
        assert np.array_equal(iv[c], res["iv"][c]), c
    del h
