"""The native FASTA/FASTQ(.gz) ingest (gvs_fastx_read, csrc/fastx.cu) against the Python restatement of
readfq in gavisunk_b200.io -- which tests/test_oracle_golden.py::test_rlen_b8 pins to the reference's
`rlen` executable -- on the golden vector and on adversarial / random inputs.  CPU only (no kernel)."""
import gzip
import random

import numpy as np
import pytest

from conftest import load_golden
from gavisunk_b200 import io as gio


def _check(tmp_path, datas, gz=()):
    paths = []
    for i, d in enumerate(datas):
        p = tmp_path / f"f{i}.fx{'.gz' if i in gz else ''}"
        p.write_bytes(gzip.compress(d) if i in gz else d)
        paths.append(str(p))
    nr = gio.NativeReads(paths, threads=3, pin=False)
    want, cf = [], [0]
    for d in datas:
        want += gio.read_fastx(d) if d else []
        cf.append(len(want))
    assert nr.n_reads == len(want)
    assert nr.names == [n for n, _ in want]
    off = nr.read_off.tolist()
    for i, (_, s) in enumerate(want):
        assert bytes(nr.seq[off[i]:off[i + 1]]) == s
    assert nr.chunk_first.tolist() == cf
    assert nr.total_bases == sum(len(s) for _, s in want)
    nr.close()


def test_golden_rlen_b8(tmp_path):
    case = load_golden("rlen_b8")
    data = case["reads"].encode("latin-1")
    _check(tmp_path, [data, data], gz={1})
    p = tmp_path / "r.fa"
    p.write_bytes(data)
    nr = gio.NativeReads([str(p)], pin=False)
    assert "".join(f"{n}\t{l}\n" for n, l in zip(nr.names, nr.lengths().tolist())) == case["rlen"]


def test_adversarial_records(tmp_path):
    cases = [
        b"", b"\n\n", b"garbage before\n>r1 desc\nACGT\nAC GT\r\n\n>r2\n>r3\t x\nNNN",
        b"@q1\nACGT\n+\nIIII\n@q2 c\nAC\nGT\n+q2\nII\nII\n>f1\nTTTT\n@q3\n\n+\n\n@q4\nAAAA\n+\nII",  # truncated quality
        b">\nACGT\n> leading blank\nGG\n>\t\nCC\n>a\r\nAC\r\r\n", b"@x\nACGT\n+\n@@@@\n@y\nTT\n+\n@@\n",
        b">only_header", b">h\nA+C\n+not a marker inside? it is\nACGT\n", b"@e\n\n+\n\n>z\nAC\n",
    ]
    for i, c in enumerate(cases):
        _check(tmp_path, [c], gz={0} if i % 2 else ())
    _check(tmp_path, cases)


def test_random_fuzz(tmp_path):
    rng = random.Random(7)
    alphabet = [b">", b"@", b"+", b"\n", b"\r\n", b"A", b"C", b"G", b"T", b"N", b" ", b"\t", b"ACGTACGTAC", b"x1", b"\n\n"]
    datas = []
    for _ in range(300):
        datas.append(b"".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 60))))
    for lo in range(0, 300, 50):
        _check(tmp_path, datas[lo:lo + 50], gz=set(range(0, 50, 3)))


def test_missing_file(tmp_path):
    with pytest.raises(IOError):
        gio.NativeReads([str(tmp_path / "nope.fq.gz")], pin=False)


def test_large_multiline_fasta_and_fastq(tmp_path):
    rng = np.random.default_rng(3)
    recs = [(f"r{i}", bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(n)))) for i, n in enumerate(rng.integers(0, 90000, 40))]
    fa = b"".join(b">" + n.encode() + b" d\n" + b"\n".join(s[i:i + 61] for i in range(0, len(s), 61)) + b"\n" for n, s in recs)
    fq = b"".join(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"@" * len(s) + b"\n" for n, s in recs)
    _check(tmp_path, [fa, fq, fa], gz={1, 2})


def test_pack_2bit_is_kmer_encode_byte_map(tmp_path):
    """gvs_pack_2bit / NativeReads.pack() against the byte map the reference's kmer.encode applies (pinned
    by tests/golden/kat_bytes through the ELF): all 256 byte values, ragged tails, multi-threaded split"""
    import gavisunk_oracle as O
    from gavisunk_b200.engine import pack_2bit
    rng = np.random.default_rng(5)
    for n in (0, 1, 15, 16, 17, 127, 128, 129, 255, 256, 4099, 1_200_003, 700_001):
        seq = rng.integers(0, 256, n, dtype=np.uint8) if n < 5000 else rng.choice(np.frombuffer(b"ACGTNacgtnU\x01\x02\x03-", np.uint8), n)
        if n == 700_001:  # read-like: letters only except for a sprinkle of arbitrary bytes (the 128-base AVX2 step
            seq = rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), n)  # of hostpack.cpp and its table path side by side)
            at = rng.integers(0, n, n // 1000)
            seq[at] = rng.integers(0, 256, len(at), dtype=np.uint8)
        got = pack_2bit(seq, threads=3)
        codes = O.codes_of(seq.tobytes()).astype(np.uint64)
        pad = np.zeros((-n) % 16, np.uint64)
        c = np.concatenate([codes, pad]).reshape(-1, 16)
        want = np.zeros(len(c), np.uint64)
        for i in range(16):
            want = (want << np.uint64(2)) | c[:, i]
        assert np.array_equal(got, want.astype(np.uint32)), n
    fa = b">a\nACGTNNacgtu\n>b\n\n>c\n" + b"GATTACA" * 9 + b"\n"
    p = tmp_path / "x.fa"
    p.write_bytes(fa)
    nr = gio.NativeReads([str(p)], pin=False)
    assert np.array_equal(nr.pack(pin=False), pack_2bit(nr.seq))
