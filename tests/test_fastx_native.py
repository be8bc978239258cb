"""The native FASTA/FASTQ(.gz) ingest (gvs_fastx_read, csrc/fastx.cu) against the Python restatement of
readfq in gavisunk_b200.io -- which tests/test_oracle_golden.py::test_rlen_b8 pins to the reference's
`rlen` executable -- on the golden vector and on adversarial / random inputs.  CPU only (no kernel)."""
import gzip
import random

import numpy as np
import pytest

from conftest import load_golden
from gavisunk_b200 import io as gio


_LUT = np.zeros(256, np.uint8)
for _ch, _v in (("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _LUT[ord(_ch)] = _LUT[ord(_ch.lower())] = _v
_LUT[1], _LUT[2], _LUT[3] = 1, 2, 3


def _check(tmp_path, datas, gz=(), blocks=(0, 1, 3, 7, 64)):
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
    # the fused parse + 2-bit pack path (gvs_fastx_read_packed) with block sizes that cut lines, CR LF pairs, headers
    # and words anywhere: same records, words == kmer.encode's byte map of the same bases
    allseq = np.frombuffer(b"".join(s for _, s in want), np.uint8)
    tot = len(allseq)
    nw = (tot + 15) // 16
    pad = np.zeros(nw * 16, np.uint8)
    pad[:tot] = _LUT[allseq]
    exp = np.zeros(nw, np.uint32)
    for i in range(16):
        exp |= pad[i::16].astype(np.uint32) << np.uint32(30 - 2 * i)
    for block in blocks:
        pk = gio.NativeReads(paths, threads=3, pin=False, packed=True, block_bytes=block)
        assert pk.seq is None and pk.n_reads == len(want) and pk.total_bases == tot
        assert pk.names == [n for n, _ in want]
        assert np.diff(pk.read_off).tolist() == [len(s) for _, s in want]
        assert pk.chunk_first.tolist() == cf
        assert len(pk.words) == nw and np.array_equal(pk.words, exp), block
        pk.close()


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
        b">cr\nAC\rGT\r\r\nTT\r", b"@p\nAC\n+", b"@p\nAC\n+\n", b">x\n\r\n\r\nAC\n",
        # header white space as the reference's rlen treats it (names "", "a", "", "", ""): no leading-blank skipping
        b">\r@\rT\nC\rNUN\n>a\rb c\nAC\n>\rx\nGG\n> y\nTT\n>\x0bz\nAA\n",
    ]
    for i, c in enumerate(cases):
        _check(tmp_path, [c], gz={0} if i % 2 else ())
    _check(tmp_path, cases)


def test_random_fuzz(tmp_path):
    rng = random.Random(7)
    alphabet = [b">", b"@", b"+", b"\n", b"\r\n", b"A", b"C", b"G", b"T", b"N", b" ", b"\t", b"ACGTACGTAC", b"x1", b"\n\n", b"\r", b"U", b"\x01"]
    datas = []
    for _ in range(300):
        datas.append(b"".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 60))))
    for lo in range(0, 300, 50):
        _check(tmp_path, datas[lo:lo + 50], gz=set(range(0, 50, 3)), blocks=(0, 1, 4, 11))


def test_missing_file(tmp_path):
    with pytest.raises(IOError):
        gio.NativeReads([str(tmp_path / "nope.fq.gz")], pin=False)


def test_large_multiline_fasta_and_fastq(tmp_path):
    rng = np.random.default_rng(3)
    recs = [(f"r{i}", bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(n)))) for i, n in enumerate(rng.integers(0, 90000, 40))]
    fa = b"".join(b">" + n.encode() + b" d\n" + b"\n".join(s[i:i + 61] for i in range(0, len(s), 61)) + b"\n" for n, s in recs)
    fq = b"".join(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"@" * len(s) + b"\n" for n, s in recs)
    _check(tmp_path, [fa, fq, fa], gz={1, 2}, blocks=(0, 4096, 100003))


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


def test_header_white_space_as_the_reference_rlen(tmp_path):
    """output of the reference's `rlen` executable for this input (run in the build container, SURVEY B.8 style
    probe): the name ends at the first white-space byte -- space, tab, CR, VT, FF -- and leading white space is not
    skipped; an internal CR of a sequence line counts as a base"""
    data = b">\r@\rT\nC\rNUN\n>a\rb c\nAC\n>\rx\nGG\n> y\nTT\n>\x0bz\nAA\n"
    want = "\t5\na\t2\n\t2\n\t2\n\t2\n"
    p = tmp_path / "ws.fa"
    p.write_bytes(data)
    for packed in (False, True):
        nr = gio.NativeReads([str(p)], pin=False, packed=packed)
        assert "".join(f"{n}\t{l}\n" for n, l in zip(nr.names, nr.lengths().tolist())) == want
    assert "".join(f"{n}\t{len(s)}\n" for n, s in gio.read_fastx(data)) == want
