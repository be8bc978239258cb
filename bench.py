#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: ONT read Gbp/s SUNK-matched + distance-validated.

One "step" = one pass of the whole hot path (match -> best-contig filter -> bad-SUNK histogram ->
inter-SUNK distance validation -> contig-wide intervals -> gaps) over one batch of
synthetic reads that is already resident in HBM, against a SUNK database built on the GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload s150|h3100|tiny] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; reads are sharded (each rank owns its own shard of
the same size: weak scaling), the SUNK table is replicated; NCCL all-reduces the group-hit
histogram and all-gathers the union-find forests.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ont_read_gbp_per_s_sunk_matched_validated"
UNIT = "Gbp/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s150", choices=["tiny", "s150", "h3100"])
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--asm-mbp", type=float, default=None, help="override haploid assembly size (Mbp)")
    ap.add_argument("--coverage", type=float, default=None, help="override read coverage per GPU shard")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-mbp", type=float, default=None)
    return ap.parse_args()


WORKLOADS = {
    # name: (haploid Mbp, contigs, coverage per shard, N50, description)
    "tiny": dict(mbp=2.0, human=False, cov=30.0, n50=20000.0,
                 desc="synthetic 2 Mbp x2 diploid + 30x ONT-like reads (smoke size)"),
    "s150": dict(mbp=150.0, human=False, cov=30.0, n50=50000.0,
                 desc="synthetic 150 Mbp single-chromosome x2 diploid assembly + 30x simulated ONT reads (N50 ~50 kb)"),
    "h3100": dict(mbp=3100.0, human=True, cov=3.75, n50=100000.0,
                  desc="synthetic 3.1 Gbp x2 diploid assembly, 23 contigs + ONT-like reads N50 ~100 kb, 3.75x (=30x/8) per GPU shard"),
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=(float(np.median(sm)) if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=sorted(reasons), samples=len(sm))


def build_workload(args, eng, rank):
    from gavisunk_b200 import workload as W
    spec = dict(WORKLOADS[args.workload])
    if args.asm_mbp is not None:
        spec["mbp"] = args.asm_mbp
    if args.coverage is not None:
        spec["cov"] = args.coverage
    L = int(spec["mbp"] * 1e6)
    contigs = W.human_contigs(spec["mbp"]) if spec["human"] else [L]
    t0 = time.time()
    wl = W.make_assembly(eng, contigs, snp_rate=1e-3, dup_frac=0.01, seed=1001, name=args.workload)
    eng.set_profiling(True)
    W.build_db(eng, wl)
    db_ms = eng.stage_ms("dbbuild")
    # every rank draws its own shard of reads (different seed) from the same assembly
    W.add_reads(eng, wl, coverage=spec["cov"], n50=spec["n50"], sigma=0.8, len_min=1000, len_max=1000000,
                seed=2001 + 16 * rank, nchunks=10)
    wl.meta.update(db_build_ms=db_ms, setup_s=time.time() - t0, desc=spec["desc"], asm_mbp=spec["mbp"])
    return wl


def run_step(eng, wl, bind, dist_ctx=None):
    """one pass of the hot path; returns a small dict of result sizes"""
    bind()
    n_rows = eng.match()
    n_best, n_kept = eng.diag_filter(wl.contig_hap)
    return dict(rows=n_rows, best=n_best, kept=n_kept)


def main_b200(args):
    import torch
    import torch.distributed as dist
    from gavisunk_b200 import workload as W
    from gavisunk_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(args.k, device=local, stream=torch.cuda.current_stream().cuda_stream)
    wl = build_workload(args, eng, rank)
    n_sunks, n_groups = eng.db_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bind = lambda: W.bind_reads(eng, wl)
    # ---- HBM-resident timing ----
    for _ in range(args.warmup):
        res = run_step(eng, wl, bind)
    launches0 = eng.launches
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    e0.record()
    for _ in range(args.steps):
        res = run_step(eng, wl, bind)
        for st in ("probe", "emit", "diag"):
            stage_ms.setdefault(st, []).append(eng.stage_ms(st))
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    bases = torch.tensor([wl.total_bases], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(bases, op=dist.ReduceOp.SUM)
    total_bases = float(bases.item())
    value = total_bases * args.steps / (ms_max * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (k_probe) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k = args.k
    off = wl.read_off.cpu().numpy()
    lens = np.diff(off)
    windows = int(np.maximum(lens - k + 1, 0).sum())
    probe_bytes = float(wl.total_bases) * 1.0 + windows * 16.0
    probe_ms = float(np.mean(stage_ms["probe"]))
    achieved = probe_bytes / (probe_ms * 1e-3) / 1e9
    roofline = dict(bound="hbm", kernel="k_probe", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=None, peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                    alg_bytes_per_launch=probe_bytes, kernel_ms=probe_ms,
                    kernel_share_of_step=probe_ms * args.steps / ms)

    # ---- end to end through the public API with HOST buffers ----
    e2e = None
    if args.e2e_steps > 0:
        h_reads = torch.empty(wl.total_bases + 64, dtype=torch.uint8, pin_memory=True)
        h_reads[:wl.total_bases].copy_(wl.reads[:wl.total_bases])
        h_off = torch.empty(wl.n_reads + 1, dtype=torch.int64, pin_memory=True)
        h_off.copy_(wl.read_off)
        torch.cuda.synchronize()
        np_reads = h_reads.numpy()[:wl.total_bases]
        np_off = h_off.numpy().view(np.uint64)
        bind_h = lambda: eng.set_reads(np_reads, np_off, wl.chunk_first, wl.chunk_hap)
        run_step(eng, wl, bind_h)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        d2h = 0
        for _ in range(args.e2e_steps):
            r = run_step(eng, wl, bind_h)
            kept = eng.rows(1)
            d2h = sum(a.nbytes for a in kept.values())
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ems = max(f0.elapsed_time(f1), wall_ms)
        t = torch.tensor([ems], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = dict(value=total_bases * args.e2e_steps / (float(t.item()) * 1e-3) / 1e9, unit=UNIT,
                   h2d_bytes_per_step=int(wl.total_bases + 8 * (wl.n_reads + 1)), d2h_bytes_per_step=int(d2h),
                   steps=args.e2e_steps)
        del h_reads, h_off

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = cpu_reference_sample(args, wl, eng)
        except Exception as ex:  # the baseline must never take the bench down
            cpu_baseline = dict(error=str(ex)[:200])

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl.meta["desc"], "k": k, "asm_haploid_mbp": wl.meta["asm_mbp"],
                       "read_gbp_per_gpu": wl.total_bases / 1e9, "reads_per_gpu": wl.n_reads, "n_sunks": n_sunks,
                       "n_groups": n_groups, "rows_per_step": res, "l2": "inputs larger than L2 (reads >> 126 MB)",
                       "stages": ["match", "diag_filter"], "db_build_ms": wl.meta["db_build_ms"],
                       "stage_ms": {s: float(np.mean(v)) for s, v in stage_ms.items()}},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_reference_sample(args, wl, eng):
    """placeholder until the reference arm lands"""
    return None


def main_reference(args):
    print(json.dumps({"impl": "reference", "unavailable": "reference arm not wired yet"}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
