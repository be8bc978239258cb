"""Native text egress (gvs_format_rows, csrc/format.cu) against plain Python formatting of the same rows: the
file formats of SURVEY Appendix C (.sunkpos, .rlen, kmer.loc, jellyfish.fa, inter_outs, BED).  No GPU."""
import numpy as np
import pytest

from gavisunk_b200 import io as gio
from gavisunk_b200.engine import encode_kmers


def test_sunkpos_rows_and_selection():
    rng = np.random.default_rng(5)
    n = 50000
    rnames = [f"read/{i:x} " if i % 7 else "" for i in range(300)]  # empty names and odd characters survive
    cnames = ["chr1", "h1tg#000001l", "AMY_h1", "c" * 70]
    rt, ct = gio.NameTable(rnames), gio.NameTable(cnames)
    r = rng.integers(0, len(rnames), n).astype(np.uint32)
    c = rng.integers(0, len(cnames), n).astype(np.uint32)
    pos = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    pos[:5] = [0, 9, 10, 4294967295, 1000000000]
    st = rng.integers(0, 3_000_000_000, n).astype(np.uint32)
    g = rng.integers(0, 3_000_000_000, n).astype(np.uint32)
    cols = [("name", r, rt), ("u32", pos), ("name", c, ct), ("u32", st), ("u32", g)]
    want = [f"{rnames[a]}\t{b}\t{cnames[cc]}\t{d}\t{e}\n" for a, b, cc, d, e in zip(r.tolist(), pos.tolist(), c.tolist(), st.tolist(), g.tolist())]
    for threads in (1, 3, 16):
        assert bytes(gio.format_rows(cols, threads=threads)) == "".join(want).encode("latin-1")
    sel = rng.permutation(n)[:777]
    assert bytes(gio.format_rows(cols, sel=sel)) == "".join(want[i] for i in sel).encode("latin-1")
    assert bytes(gio.format_rows(cols, sel=np.zeros(0, np.uint64))) == b""
    assert bytes(gio.format_rows([("u32", np.zeros(0, np.uint32))])) == b""


@pytest.mark.parametrize("k", [1, 16, 20, 31, 32])
def test_kmer_and_signed_columns(k):
    rng = np.random.default_rng(k)
    kms = ["".join(rng.choice(list("ACGT"), k)) for _ in range(200)]
    enc = encode_kmers([x.encode() for x in kms])
    assert bytes(gio.format_rows([("kmer", enc, k)])) == "".join(x + "\n" for x in kms).encode()
    fa = gio.format_rows([("kmer", enc, k, dict(prefix=b">", sep=b"\n")), ("kmer", enc, k)])
    assert bytes(fa) == "".join(f">{x}\n{x}\n" for x in kms).encode()
    v = np.array([-(2 ** 63), -200001, -1, 0, 7, 2 ** 63 - 1], np.int64)
    u = np.array([0, 1, 2 ** 64 - 1, 10, 99, 100], np.uint64)
    assert bytes(gio.format_rows([("i64", v, dict(sep=b":")), ("u64", u)])) == "".join(f"{a}:{b}\n" for a, b in zip(v.tolist(), u.tolist())).encode()


def test_name_table_concat_and_bytes_array():
    a, b = gio.NameTable(["x", "", "yz"]), gio.NameTable(["", "q"])
    c = gio.NameTable.concat([a, b, gio.NameTable([])])
    assert len(c) == 5 and c.as_bytes_array().tolist() == [b"x", b"", b"yz", b"", b"q"]
    assert bytes(gio.format_rows([("name", np.arange(5, dtype=np.uint32), c)])) == b"x\n\nyz\n\nq\n"
