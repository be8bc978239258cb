"""File formats at the drop-in boundary: FASTA/FASTQ(.gz) reads, .fai,
jellyfish.db, kmer.loc, .sunkpos, .rlen, BED.  Host-side plumbing only -- no compute here.

The FASTA/FASTQ reader mirrors the `readfq` port the reference's Nim tools use
(workflow/src/kmerpos_annot3.nim:85, workflow/src/rlen.nim:13; behaviour pinned by probe B.8 /
tests/golden/rlen_b8.json.gz): the record name ends at the first space or tab, multi-line
sequences are concatenated, blank lines are ignored, a trailing CR is stripped, internal
spaces are kept as sequence characters, empty records are emitted, FASTA and FASTQ records may
be mixed in one file.
"""
from __future__ import annotations

import gzip
import io
import os
from typing import Iterable, Iterator, List, Tuple

import numpy as np


def open_maybe_gz(path: str):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rb")
    return open(path, "rb")


def iter_fastx(fp) -> Iterator[Tuple[str, bytes]]:
    """Heng Li's readfq state machine over a binary line iterator."""
    last = None
    it = iter(fp)
    while True:
        if last is None:
            for l in it:
                if l[:1] in (b">", b"@"):
                    last = l.rstrip(b"\r\n")
                    break
            if last is None:
                return
        hdr = last[1:]
        # name = header up to the first white-space byte, without skipping leading white space (pinned through the
        # reference's `rlen`: ">\r@\rT" -> "", ">a\rb c" -> "a", "> y" -> "", ">\x0bz" -> "")
        name = hdr
        for i, ch in enumerate(hdr):
            if ch in b" \t\n\r\x0b\x0c":
                name = hdr[:i]
                break
        seqs: List[bytes] = []
        last = None
        for l in it:
            if l[:1] in (b"@", b"+", b">"):
                last = l.rstrip(b"\r\n")
                break
            seqs.append(l.rstrip(b"\r\n"))
        seq = b"".join(seqs)
        if last is None or last[:1] != b"+":
            yield name.decode("latin-1"), seq
            if last is None:
                return
        else:  # FASTQ: skip quality lines
            qlen = 0
            last = None
            for l in it:
                qlen += len(l.rstrip(b"\r\n"))
                if qlen >= len(seq):
                    break
            yield name.decode("latin-1"), seq
            # (truncated quality: record still yielded, as readfq does)


def read_fastx(path_or_bytes) -> List[Tuple[str, bytes]]:
    if isinstance(path_or_bytes, (bytes, bytearray)):
        data = bytes(path_or_bytes)
        if data[:2] == b"\x1f\x8b":
            data = gzip.decompress(data)
        return list(iter_fastx(io.BytesIO(data)))
    with open_maybe_gz(path_or_bytes) as f:
        return list(iter_fastx(f))


class NativeReads:
    """Reads of one or more FASTA/FASTQ(.gz) files parsed by libgavisunk_b200.so (gvs_fastx_read): one
    chunk per file, sequences back to back in one (page-locked when possible) host buffer.  The numpy
    arrays are views of library memory: keep this object alive while they are in use."""

    def __init__(self, paths, threads: int = 0, pin: bool = True, packed: bool = False, block_bytes: int = 0):
        """packed=True: the fused parse + 2-bit pack path (gvs_fastx_read_packed): `words` instead of `seq`, the
        ASCII bases never materialise -- what Engine.set_reads_packed takes."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self._lib = lib
        self._fx = _lib.FastxStruct()
        paths = [os.fspath(p) for p in paths]
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        err = C.create_string_buffer(512)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        nthr = threads or min(len(paths), ncpu)
        if nthr < len(paths) <= 2 * nthr:
            nthr = len(paths)  # one thread per file (a file is parsed by one thread): 20 files on 16 cores would take two rounds
        if packed:
            rc = lib.gvs_fastx_read_packed(arr, len(paths), nthr, 1 if pin else 0, int(block_bytes), C.byref(self._fx), err, 512)
        else:
            rc = lib.gvs_fastx_read(arr, len(paths), nthr, 1 if pin else 0, C.byref(self._fx), err, 512)
        if rc != 0:
            self._fx = None
            raise IOError(err.value.decode("utf-8", "replace"))
        fx = self._fx
        n, tot = int(fx.n_reads), int(fx.total_bases)
        as_np = lambda ptr, ctype, cnt: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(cnt,)) if cnt else np.zeros(0, ctype)
        self.seq = None if packed else as_np(fx.seq, C.c_uint8, tot)
        self.words = as_np(fx.words, C.c_uint32, int(fx.n_words)) if packed else None
        self.read_off = as_np(fx.read_off, C.c_uint64, n + 1)
        self.chunk_first = as_np(fx.chunk_first, C.c_uint64, int(fx.n_files) + 1).copy()
        self.name_off = as_np(fx.name_off, C.c_uint64, n + 1).copy()
        self.name_blob = C.string_at(fx.names, int(self.name_off[-1])) if n else b""
        self._names = None
        self.n_reads, self.total_bases, self.pinned = n, tot, bool(fx.pinned)

    @property
    def names(self):
        """read names as Python strings (built on first use; the formatter works on name_blob / name_off)"""
        if self._names is None:
            raw, no = self.name_blob, self.name_off.tolist()
            self._names = [raw[no[i]:no[i + 1]].decode("latin-1") for i in range(self.n_reads)]
        return self._names

    def lengths(self) -> np.ndarray:
        return np.diff(self.read_off).astype(np.uint64)

    def pack(self, threads: int = 0, pin: bool = True) -> np.ndarray:
        """2-bit words of the batch (gvs_fastx_pack) for Engine.set_reads_packed; view of library memory"""
        import ctypes as C
        rc = self._lib.gvs_fastx_pack(C.byref(self._fx), threads or (os.cpu_count() or 1), 1 if pin else 0)
        if rc != 0:
            raise MemoryError("gvs_fastx_pack failed")
        n = int(self._fx.n_words)
        self.words = np.ctypeslib.as_array(C.cast(self._fx.words, C.POINTER(C.c_uint32)), shape=(n,)) if n else np.zeros(0, np.uint32)
        return self.words

    def close(self):
        if getattr(self, "_fx", None) is not None:
            self.seq = self.read_off = self.words = None  # (name_blob / name_off / chunk_first are copies)
            self._lib.gvs_fastx_free(self._fx)
            self._fx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_reads(reads: Iterable[Tuple[str, bytes]]):
    """-> (names, seq uint8[total], off uint64[n+1])"""
    names, parts, lens = [], [], []
    for n, s in reads:
        names.append(n)
        parts.append(s)
        lens.append(len(s))
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    if lens:
        off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8)
    return names, seq, off


def read_fai(path: str) -> List[Tuple[str, int]]:
    out = []
    with open(path) as f:
        for l in f:
            p = l.rstrip("\n").split("\t")
            if len(p) >= 2:
                out.append((p[0], int(p[1])))
    return out


def read_db(path: str) -> List[str]:
    with open(path) as f:
        return [l.rstrip("\n") for l in f if l.rstrip("\n") != ""]


def read_loc(path: str):
    """kmer.loc: contig \t start0 \t kmer \t group_start  (defineSUNKs.smk:126)"""
    contig, start, kmer, group = [], [], [], []
    with open(path) as f:
        for l in f:
            p = l.rstrip("\n").split("\t")
            if len(p) < 4:
                continue
            contig.append(p[0])
            start.append(int(p[1]))
            kmer.append(p[2])
            group.append(int(p[3]))
    return contig, start, kmer, group


def parse_sunkpos(text: str):
    """rows (read, pos, contig, start, group) of sunkpos TEXT (tests, in-memory data)"""
    rows = []
    for l in text.splitlines():
        p = l.split("\t")
        if len(p) < 5:
            continue
        rows.append((p[0], int(p[1]), p[2], int(p[3]), int(p[4])))
    return rows


def read_sunkpos(path: str):
    """rows (read, pos, contig, start, group) of a .sunkpos FILE (plain or gz).  The path must open: a missing
    or unreadable file raises (FileNotFoundError / OSError) exactly where the reference's tools quit with
    'cannot open the file' (nim `open`) or pandas raises -- never an empty result."""
    with open_maybe_gz(os.fspath(path)) as f:
        return parse_sunkpos(f.read().decode())


def format_sunkpos(rows) -> str:
    return "".join(f"{r[0]}\t{r[1]}\t{r[2]}\t{r[3]}\t{r[4]}\n" for r in rows)


def write_bed(path: str, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write("\t".join(str(x) for x in r) + "\n")


# ------------------------------------------------------------------------------------------------
# native text egress (gvs_format_rows, csrc/format.cu)
# ------------------------------------------------------------------------------------------------
class NameTable:
    """names back to back + offsets: what a GVS_COL_NAME column indexes"""

    def __init__(self, names=None, blob: bytes = None, off: np.ndarray = None):
        if blob is None:
            enc = [n.encode("latin-1") for n in names]
            blob = b"".join(enc)
            off = np.zeros(len(enc) + 1, np.uint64)
            if enc:
                off[1:] = np.cumsum(np.fromiter((len(e) for e in enc), np.uint64, len(enc)))
        self.blob = blob
        self.off = np.ascontiguousarray(off, np.uint64)

    @staticmethod
    def concat(tables):
        blob = b"".join(t.blob for t in tables)
        offs, base = [np.zeros(1, np.uint64)], 0
        for t in tables:
            offs.append(t.off[1:] + np.uint64(base))
            base += len(t.blob)
        return NameTable(blob=blob, off=np.concatenate(offs))

    def __len__(self):
        return len(self.off) - 1

    def as_bytes_array(self) -> np.ndarray:
        """numpy 'S' array of the names (sorting, uniqueness)"""
        no = self.off.tolist()
        return np.array([self.blob[no[i]:no[i + 1]] for i in range(len(no) - 1)], dtype=object).astype("S") if len(no) > 1 else np.zeros(0, "S1")


def format_rows(cols, n_rows: int = None, sel: np.ndarray = None, threads: int = 0) -> bytes:
    """cols: list of ('u32'|'u64'|'i64', array) | ('name', index array, NameTable) | ('kmer', u64 array, k), each
    optionally followed by a dict(prefix=..., sep=...); cells are tab-separated, the last one ends the line.
    sel: row indices to write (in that order) instead of all rows.  -> the text as bytes"""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    kinds = {"u32": (0, np.uint32), "u64": (1, np.uint64), "i64": (2, np.int64), "name": (3, np.uint32), "kmer": (4, np.uint64)}
    arr = (_lib.ColStruct * len(cols))()
    keep = []
    n = n_rows
    for i, c in enumerate(cols):
        opt = c[-1] if isinstance(c[-1], dict) else {}
        kind, dt = kinds[c[0]]
        data = np.ascontiguousarray(c[1], dtype=dt)
        keep.append(data)
        if n is None:
            n = len(data)
        arr[i].kind = kind
        arr[i].data = data.ctypes.data
        arr[i].k = int(c[2]) if c[0] == "kmer" else 0
        if c[0] == "name":
            nt = c[2]
            keep.append(nt)
            arr[i].names = C.cast(C.c_char_p(nt.blob), C.c_void_p)
            arr[i].name_off = nt.off.ctypes.data
        arr[i].prefix = opt.get("prefix", b"\0")
        arr[i].sep = opt.get("sep", b"\n" if i == len(cols) - 1 else b"\t")
    sel_a = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint64)
    n_sel = 0 if sel_a is None else len(sel_a)
    sp = None if sel_a is None else C.c_void_p(sel_a.ctypes.data)
    thr = threads or min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    need = lib.gvs_format_rows(arr, len(cols), n or 0, sp, n_sel, None, 0, thr)
    if need < 0:
        raise ValueError(f"gvs_format_rows: error {need}")
    buf = bytearray(need)
    if need:
        got = lib.gvs_format_rows(arr, len(cols), n or 0, sp, n_sel, (C.c_char * need).from_buffer(buf), need, thr)
        if got != need:
            raise ValueError(f"gvs_format_rows: error {got}")
    return bytes(buf) if need < (1 << 20) else buf
