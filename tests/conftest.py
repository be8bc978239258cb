import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return load_golden


# ---- helpers shared by the GPU parity tests --------------------------------------------------
def parse_case_db(case):
    db_lines = [l.encode() for l in case["db"].splitlines() if l]
    loc_rows = []
    for l in case["loc"].splitlines():
        c, s, km, g = l.split("\t")
        loc_rows.append((c, int(s), km.encode(), int(g)))
    return db_lines, loc_rows


def engine_from_case(case, contig_names=None, variant=0):
    """variant: 0 = probe instantiation by database size (small for every golden case), 2 = the large-database
    instantiation (whole-genome filter layout) forced onto the case"""
    from gavisunk_b200.engine import Engine
    eng = Engine(case["k"])
    eng.set_probe_variant(variant)
    db_lines, loc_rows = parse_case_db(case)
    eng.load_loc_text(db_lines, loc_rows, contig_names)
    return eng, eng.contig_names


def pack_chunks(chunk_texts):
    """list of FASTA/FASTQ texts -> (names, seq, off, chunk_first)"""
    from gavisunk_b200 import io as gio
    names, parts, lens, chunk_first = [], [], [], [0]
    for txt in chunk_texts:
        for n, s in gio.read_fastx(txt.encode("latin-1")):
            names.append(n)
            parts.append(s)
            lens.append(len(s))
        chunk_first.append(len(names))
    import numpy as np
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8)
    return names, seq, off, chunk_first


def format_rows(rows, read_names, contig_names, lo, hi):
    """sunkpos text of the rows whose read index lies in [lo, hi)"""
    out = []
    for r, p, c, s, g in zip(rows["read"], rows["pos"], rows["contig"], rows["start"], rows["group"]):
        if lo <= r < hi:
            out.append(f"{read_names[r]}\t{p}\t{contig_names[c]}\t{s}\t{g}\n")
    return "".join(out)


def run_match_chunks(eng, contig_names, chunk_texts, chunk_hap=None):
    names, seq, off, chunk_first = pack_chunks(chunk_texts)
    eng.set_reads(seq, off, chunk_first, chunk_hap if chunk_hap is not None else [0] * len(chunk_texts))
    eng.match()
    rows = eng.rows(0)
    eng._read_names = names
    eng._chunk_first = chunk_first
    return [format_rows(rows, names, contig_names, chunk_first[i], chunk_first[i + 1]) for i in range(len(chunk_texts))]
