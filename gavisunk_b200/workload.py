"""Synthetic workloads of BASELINE.json's configs, generated on the device.

Used by bench.py and the GPU tests only.  torch is plumbing here: it owns the device buffers
(assembly, reads, offsets) that are handed to libgavisunk_b200.so as raw pointers.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from .engine import Engine

# human chromosome lengths (Mbp, T2T-CHM13-like, chr1..22 + X), scaled to sum 3.1 Gbp
_HUMAN_MBP = [248, 242, 201, 193, 182, 172, 160, 146, 150, 135, 135, 133, 114, 101, 100, 96, 84, 80, 62, 66, 45, 51, 154]


@dataclass
class Workload:
    name: str
    k: int
    contig_names: List[str]
    contig_len: np.ndarray          # per contig (both haps), uint64
    contig_hap: np.ndarray          # uint8 per contig
    asm: torch.Tensor               # uint8 device, hap1 contigs then hap2 contigs (+64 pad)
    contig_off: torch.Tensor        # uint64 (as int64 tensor) device, n_contigs+1
    reads: torch.Tensor = None      # uint8 device (+64 pad)
    read_off: torch.Tensor = None   # int64 device, n_reads+1
    n_reads: int = 0
    total_bases: int = 0
    chunk_first: np.ndarray = None
    chunk_hap: np.ndarray = None
    meta: dict = field(default_factory=dict)


def _dev_ptr(t: torch.Tensor) -> int:
    return int(t.data_ptr())


def make_assembly(eng: Engine, contig_len_per_hap, snp_rate=1e-3, dup_frac=0.01, seed=1001, name="syn", device=None) -> Workload:
    device = device or torch.device("cuda", eng.device)
    cl = np.asarray(contig_len_per_hap, dtype=np.uint64)
    nc = len(cl)
    hap_len = int(cl.sum())
    asm = torch.zeros(2 * hap_len + 64, dtype=torch.uint8, device=device)
    eng._ck(eng.lib.gvs_synth_assembly(eng.ctx, C.c_void_p(_dev_ptr(asm)), cl.ctypes.data_as(C.c_void_p), nc,
                                       float(snp_rate), float(dup_frac), int(seed)))
    both = np.concatenate([cl, cl])
    off = np.zeros(2 * nc + 1, dtype=np.int64)
    off[1:] = np.cumsum(both.astype(np.int64))
    names = [f"synH1_chr{i + 1}" for i in range(nc)] + [f"synH2_chr{i + 1}" for i in range(nc)]
    hap = np.array([0] * nc + [1] * nc, dtype=np.uint8)
    return Workload(name=name, k=eng.k, contig_names=names, contig_len=both, contig_hap=hap, asm=asm,
                    contig_off=torch.from_numpy(off).to(device))


@dataclass
class ReadBatch:
    """one batch of reads resident in HBM (a group of chunk files: `nchunks` per haplotype)"""
    reads: torch.Tensor             # uint8 device (+64 pad)
    read_off: torch.Tensor          # int64 device, n_reads+1
    n_reads: int
    total_bases: int
    chunk_first: np.ndarray
    chunk_hap: np.ndarray
    seed: int = 0


def new_batch(eng: Engine, wl: Workload, coverage: float, n50: float = 50000.0, sigma: float = 0.8, len_min=1000,
              len_max=1000000, seed=2001, nchunks=10, device=None, contig_range=None) -> ReadBatch:
    """a batch of reads drawn from wl's assembly (same arguments as add_reads); wl itself is not changed"""
    import copy
    tmp = copy.copy(wl)
    tmp.meta = dict(wl.meta)
    add_reads(eng, tmp, coverage, n50, sigma, len_min, len_max, seed, nchunks, device, contig_range)
    return ReadBatch(tmp.reads, tmp.read_off, tmp.n_reads, tmp.total_bases, tmp.chunk_first, tmp.chunk_hap, seed)


def bind_batch(eng: Engine, b: ReadBatch):
    if getattr(b, "words", None) is not None:
        eng.set_reads_packed_device(_dev_ptr(b.words), _dev_ptr(b.read_off), b.n_reads, b.chunk_first, b.chunk_hap)
    else:
        eng.set_reads_device(_dev_ptr(b.reads), _dev_ptr(b.read_off), b.n_reads, b.chunk_first, b.chunk_hap)


def pack_batch_on_device(b: ReadBatch, drop_ascii: bool = True):
    """the batch's bases as resident 2-bit words (kmer.encode's byte map: the layout the ingest emits and
    gvs_reads_set_packed takes), made with torch in slices -- bench / test set-up only"""
    dev = b.reads.device
    lut = torch.zeros(256, dtype=torch.uint8, device=dev)
    for ch, v in (("C", 1), ("G", 2), ("T", 3), ("U", 3)):
        lut[ord(ch)] = v
        lut[ord(ch.lower())] = v
    lut[1], lut[2], lut[3] = 1, 2, 3
    nw = (b.total_bases + 15) // 16
    words = torch.zeros(nw + 16, dtype=torch.int32, device=dev)
    sh = torch.tensor([30 - 2 * i for i in range(16)], dtype=torch.int64, device=dev)
    step = 1 << 24  # words per slice
    for w0 in range(0, nw, step):
        w1 = min(nw, w0 + step)
        seg = b.reads[16 * w0:min(16 * w1, b.total_bases)]
        codes = lut[seg.long()].long()
        if codes.numel() < 16 * (w1 - w0):
            codes = torch.cat([codes, torch.zeros(16 * (w1 - w0) - codes.numel(), dtype=torch.int64, device=dev)])
        v = (codes.view(-1, 16) << sh).sum(dim=1)
        words[w0:w1] = (v & 0xFFFFFFFF).to(torch.int64).where(v < (1 << 31), v - (1 << 32)).to(torch.int32)
    b.words = words
    if drop_ascii:
        b.reads = None
    torch.cuda.synchronize(dev)
    return b


def add_reads(eng: Engine, wl: Workload, coverage: float, n50: float = 50000.0, sigma: float = 0.8, len_min=1000,
              len_max=1000000, seed=2001, nchunks=10, device=None, contig_range=None) -> Workload:
    """coverage = total read bases / haploid size of the sampled contigs, split evenly across the two
    haps.  contig_range=(lo, hi): draw reads only from contigs lo..hi-1 of each haplotype."""
    device = device or wl.asm.device
    nc = len(wl.contig_names) // 2
    c_lo, c_hi = contig_range if contig_range is not None else (0, nc)
    hap_len = int(wl.contig_len[c_lo:c_hi].sum())
    mu = math.log(n50) - sigma * sigma           # length-weighted median of a log-normal = exp(mu + sigma^2)
    mean_len = math.exp(mu + sigma * sigma / 2)
    per_hap_bases = coverage * hap_len / 2
    n_per_hap = max(1, int(round(per_hap_bases / mean_len)))
    offs, totals = [], []
    for hap in range(2):
        ro = torch.zeros(n_per_hap + 1, dtype=torch.int64, device=device)
        tot = C.c_uint64()
        eng._ck(eng.lib.gvs_synth_reads_plan(eng.ctx, C.c_void_p(_dev_ptr(wl.contig_off)), hap * nc + c_lo, hap * nc + c_hi,
                                             n_per_hap, mu, sigma, len_min, len_max, seed + hap,
                                             C.c_void_p(_dev_ptr(ro)), C.byref(tot)))
        offs.append(ro)
        totals.append(int(tot.value))
        if hap == 0:
            # hap-1 reads are filled right away: the plan state is per call
            pad0 = (-totals[0]) % 16
            total0 = totals[0]
            est = int(total0 * 2.2) + 4096
            reads = torch.zeros(est + 64, dtype=torch.uint8, device=device)
            eng._ck(eng.lib.gvs_synth_reads_fill(eng.ctx, C.c_void_p(_dev_ptr(wl.asm)), C.c_void_p(_dev_ptr(reads)),
                                                 C.c_void_p(_dev_ptr(ro)), n_per_hap))
    total = totals[0] + totals[1]
    if total + 64 > reads.numel():
        bigger = torch.zeros(total + 64, dtype=torch.uint8, device=device)
        bigger[:totals[0]] = reads[:totals[0]]
        reads = bigger
    ro2 = offs[1] + totals[0]
    eng._ck(eng.lib.gvs_synth_reads_fill(eng.ctx, C.c_void_p(_dev_ptr(wl.asm)), C.c_void_p(_dev_ptr(reads)),
                                         C.c_void_p(_dev_ptr(ro2)), n_per_hap))
    read_off = torch.cat([offs[0][:-1], ro2]).contiguous()
    n_reads = 2 * n_per_hap
    cf, ch = [], []
    for hap in range(2):
        for c in range(nchunks):
            cf.append(hap * n_per_hap + (c * n_per_hap) // nchunks)
            ch.append(hap)
    cf.append(n_reads)
    if reads.numel() > total + 64 + (64 << 20):  # the estimate left > 64 MiB of slack: a run keeps many batches resident
        reads = reads[:total + 64].clone()
    wl.reads = reads[:total + 64] if reads.numel() > total + 64 else reads
    wl.read_off = read_off
    wl.n_reads = n_reads
    wl.total_bases = total
    wl.chunk_first = np.asarray(cf, dtype=np.uint64)
    wl.chunk_hap = np.asarray(ch, dtype=np.uint8)
    wl.meta.update(coverage=coverage, n50=n50, sigma=sigma, reads_per_hap=n_per_hap, nchunks=nchunks)
    torch.cuda.synchronize(device)
    return wl


def human_contigs(total_mbp=3100.0):
    s = sum(_HUMAN_MBP)
    return [int(x / s * total_mbp * 1e6) for x in _HUMAN_MBP]


def bind_reads(eng: Engine, wl: Workload):
    eng.set_reads_device(_dev_ptr(wl.reads), _dev_ptr(wl.read_off), wl.n_reads, wl.chunk_first, wl.chunk_hap)


def build_db(eng: Engine, wl: Workload):
    eng.build_db_device(_dev_ptr(wl.asm), _dev_ptr(wl.contig_off), wl.contig_names)
