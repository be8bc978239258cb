"""Runs the reference's own prebuilt executables (oracle/_ref, copied from
/root/reference/workflow/scripts by oracle/Makefile) the way the Snakemake rules do
(workflow/rules/tagONT.smk:36 SUNK_annot, :92 diag_filter_step, :110 diag_filter_final).

TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline / --impl reference legs and
by tests; never by the product path.
"""
from __future__ import annotations

import os
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
BINS = ("kmerpos_annot3", "rlen", "diag_filter_v3", "diag_filter_step2")


def available() -> bool:
    return all(os.access(os.path.join(REF_DIR, b), os.X_OK) for b in BINS)


def run_chunk(workdir: str, tag: str, reads: str, db: str, loc: str, fai: str):
    """one scatter item: kmerpos_annot3 -> diag_filter_v3 -> diag_filter_step2; returns timings"""
    sp = os.path.join(workdir, f"{tag}.sunkpos")
    dg = os.path.join(workdir, f"{tag}_diag.sunkpos")
    d2 = os.path.join(workdir, f"{tag}_diag2.sunkpos")
    t0 = time.perf_counter()
    subprocess.run([os.path.join(REF_DIR, "kmerpos_annot3"), reads, db, loc, sp], check=True, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    with open(dg, "wb") as f:
        subprocess.run([os.path.join(REF_DIR, "diag_filter_v3"), sp, fai], check=True, stdout=f)
    with open(d2, "wb") as f:
        subprocess.run([os.path.join(REF_DIR, "diag_filter_step2"), sp, dg], check=True, stdout=f)
    t2 = time.perf_counter()
    return dict(tag=tag, sunkpos=sp, diag=dg, diag2=d2, t_match=t1 - t0, t_diag=t2 - t1)


def run_chunks_parallel(jobs, n_procs: int):
    """jobs: list of kwargs for run_chunk; all chunks concurrently on n_procs cores
    (mirrors `snakemake --cores N`)."""
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=n_procs) as ex:
        res = list(ex.map(lambda kw: run_chunk(**kw), jobs))
    return res, time.perf_counter() - t0


def table_load_seconds(workdir: str, db: str, loc: str) -> float:
    """kmerpos_annot3 on an empty read file = its per-process table load (nim:20-69)"""
    empty = os.path.join(workdir, "empty.fa")
    open(empty, "w").close()
    t0 = time.perf_counter()
    subprocess.run([os.path.join(REF_DIR, "kmerpos_annot3"), empty, db, loc, os.path.join(workdir, "empty.sunkpos")],
                   check=True, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0
