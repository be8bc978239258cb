#!/usr/bin/env python
"""Golden vectors for the PYTHON stages of the reference (badsunks_AR.py, split_locs.py,
process-by-contig_lowmem_AR.py, get_gaps.py, covprob.py), produced by running the reference's OWN scripts from
/root/reference/workflow/scripts -- unmodified -- on sunkpos rows that the reference's own executables
produced.  graph-tool, pyranges, matplotlib and seaborn cannot be installed here, so they are replaced by the
minimal stand-ins under refpy_stubs/ (only the calls the scripts make), and pandas-1.3 behaviours the scripts
rely on are restored by refpy_stubs/refpy_compat.py; sympy, scipy, numpy and pandas are the real packages.
What this pins: every line of the scripts themselves (filters, sort orders, thresholds, output formats).
What it cannot pin: tie rules that live inside graph-tool / pyranges / pandas 1.3 (documented in the stubs).

  python tests/golden/make_golden_py.py        (needs /root/reference; writes tests/golden/pystages_*.json.gz)
"""
import glob
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as M  # noqa: E402

SCRIPTS = "/root/reference/workflow/scripts"
STUBS = os.path.join(HERE, "refpy_stubs")


def run_script(script, argv=(), snakemake=None, cwd=None):
    code = ("import sys, json, runpy, types; sys.path.insert(0, %r); import refpy_compat\n"
            "g = {}\n"
            "sm = json.loads(%r)\n"
            "if sm is not None:\n"
            "    ns = lambda d: types.SimpleNamespace(**d)\n"
            "    g['snakemake'] = types.SimpleNamespace(input=ns(sm['input']), output=ns(sm['output']), config=sm['config'], wildcards=ns(sm['wildcards']))\n"
            "sys.argv = [%r] + %r\n"
            "runpy.run_path(%r, init_globals=g, run_name='__main__')\n") % (
        STUBS, json.dumps(snakemake), script, list(argv), os.path.join(SCRIPTS, script))
    p = subprocess.run([sys.executable, "-W", "ignore", "-c", code], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return p.returncode, p.stdout, p.stderr


def read(path):
    return open(path).read() if os.path.exists(path) else None


def elf_case(seed, k, n_contigs, L, n_reads, mean_len, tandem=0.0):
    """diploid toy assembly + reads through the reference executables; the read set is shaped so that the
    Python stages have something to decide: no read touches the window [40 %, 46 %) of contig 0 (two validated
    intervals with a gap between them), the last contig gets no reads at all when there are >= 3 (nodata), a
    pile-up of 60 reads on one 8 kb locus of contig 1 (SUNK groups far above mode + 4*sqrt(mode))"""
    rng = np.random.default_rng(seed)
    hap1, hap2 = M.make_asm(rng, n_contigs, L)
    db_txt, loc_txt = M.db_loc_text(hap1 + hap2, k)
    case = dict(loc=loc_txt, chunks=[],
                fai1="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap1),
                fai2="".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in hap2))
    for hapi, hap in ((1, hap1), (2, hap2)):
        reads = []
        usable = hap[:-1] if n_contigs >= 3 else hap
        i = 0
        while len(reads) < n_reads:
            ci = int(rng.integers(len(usable)))
            a = np.frombuffer(usable[ci][1], dtype=np.uint8)
            ln = int(min(len(a), max(2000, rng.lognormal(np.log(mean_len), 0.5))))
            st = int(rng.integers(0, len(a) - ln + 1))
            if ci == 0 and st < 0.46 * len(a) and st + ln > 0.40 * len(a):
                continue
            r = M.mutate(rng, a[st:st + ln])
            if rng.random() < tandem and ln > 6000:
                # the read repeats a stretch of itself (and once more, shifted): the same SUNK groups at two or three
                # read positions -> the "multipos" clean-up of process-by-contig_lowmem_AR.py:161-181
                s0 = int(rng.integers(0, ln // 2))
                seg = a[st + s0:st + s0 + int(rng.integers(2000, ln // 2))]
                r = np.concatenate([r, M.mutate(rng, seg)] + ([M.mutate(rng, seg[len(seg) // 3:])] if rng.random() < 0.5 else []))
            if rng.random() < 0.5:
                r = M.rc_bytes(r)
            reads.append((f"h{hapi}r{i:05d}", r.tobytes()))
            i += 1
        a = np.frombuffer(usable[-1][1], dtype=np.uint8)
        for j in range(60):
            st = int(0.6 * len(a)) + int(rng.integers(0, 2000))
            reads.append((f"h{hapi}p{j:05d}", M.mutate(rng, a[st:st + 8000 + int(rng.integers(0, 4000))]).tobytes()))
        for ci in range(2):
            txt = M.fasta_text(reads[ci::2], fastq=(ci == 1))
            rc, sunkpos = M.run_kmerpos(txt, db_txt, loc_txt)
            assert rc == 0
            diag, diag2 = M.run_diag(sunkpos, case["fai1"] if hapi == 1 else case["fai2"])
            case["chunks"].append(dict(hap=hapi, diag2=diag2, rlen=M.run_rlen(txt)))
    return case


def pipeline_case(seed, k, n_contigs, L, n_reads, mean_len, tandem=0.0):
    # match + diag through the reference executables, then the Python scripts
    return python_stages(elf_case(seed, k, n_contigs, L, n_reads, mean_len, tandem), k)


def python_stages(base, k):
    """base: loc, fai1, fai2 (texts) and chunks [{hap, diag2, rlen}] in scatter order -> the files the reference's
    Python scripts write for them (combine_ont's cat, badsunks_AR.py, split_locs.py, process-by-contig_lowmem_AR.py per
    contig, get_gaps.py, covprob.py)"""
    case = dict(k=k, loc=base["loc"], fai1=base["fai1"], fai2=base["fai2"], hap={})
    with tempfile.TemporaryDirectory() as d:
        for sub in ("sunkpos", "breaks", "inter_outs", "bed_files", "final_out"):
            os.makedirs(os.path.join(d, sub))
        P = lambda *a: os.path.join(d, *a)
        open(P("kmer.loc"), "w").write(base["loc"])
        for h in (1, 2):
            chunks = [c for c in base["chunks"] if c["hap"] == h]
            open(P("sunkpos", f"hap{h}.sunkpos"), "w").write("".join(c["diag2"] for c in chunks))
            open(P("sunkpos", f"hap{h}.rlen"), "w").write("".join(c["rlen"] for c in chunks))
            open(P(f"hap{h}.fai"), "w").write(base[f"fai{h}"])
            case["hap"][str(h)] = dict(sunkpos=read(P("sunkpos", f"hap{h}.sunkpos")), rlen=read(P("sunkpos", f"hap{h}.rlen")))
        rc, out, err = run_script("badsunks_AR.py", [P("hap1.fai"), P("hap2.fai"), P("sunkpos", "hap1.sunkpos"), P("sunkpos", "hap2.sunkpos"),
                                                     P("sunkpos", "bad_sunks.txt")])
        assert rc == 0, err
        case["bad_sunks"] = sorted(read(P("sunkpos", "bad_sunks.txt")).split())
        case["breaks"], case["inter_outs"], case["bed_files"] = {}, {}, {}
        for h in (1, 2):
            sm = dict(input=dict(ONT_pos=P("sunkpos", f"hap{h}.sunkpos"), kmer_loc=P("kmer.loc")), output=dict(flag=P("breaks", f"hap{h}_splits_pos.done")),
                      config={}, wildcards=dict(hap=f"hap{h}"))
            rc, out, err = run_script("split_locs.py", snakemake=sm)
            assert rc == 0, err
        for f in sorted(glob.glob(P("breaks", "*"))):
            case["breaks"][os.path.basename(f)] = read(f)
        for f in sorted(glob.glob(P("breaks", "*.sunkpos"))):
            stem = os.path.basename(f)[:-len(".sunkpos")]
            rc, out, err = run_script("process-by-contig_lowmem_AR.py", [P("breaks", stem + ".loc"), f, P("sunkpos", f"hap{stem[-1]}.rlen"),
                                                                        P("sunkpos", "bad_sunks.txt"), P("inter_outs", stem + ".tsv"),
                                                                        P("bed_files", stem + ".bed")])
            assert rc == 0, err
            case["inter_outs"][stem] = read(P("inter_outs", stem + ".tsv"))
            case["bed_files"][stem] = read(P("bed_files", stem + ".bed"))  # None: the rule's `touch` creates it empty
            if case["bed_files"][stem] is None:
                open(P("bed_files", stem + ".bed"), "w").close()
        rc, out, err = run_script("get_gaps.py", [P("hap1.fai"), P("hap2.fai"), "sample", P("bed_files") + "/", P("final_out") + "/"])
        assert rc == 0, err
        case["final_out"] = {}
        for h in (1, 2):
            for f in ("gaps.bed", "nodata.bed"):
                case["final_out"][f"hap{h}.{f}"] = read(P("final_out", f"hap{h}.{f}"))
            sm = dict(input=dict(bed=P("final_out", f"hap{h}.gaps.bed"), locs=P("kmer.loc"), rlen=P("sunkpos", f"hap{h}.rlen"), fai=P(f"hap{h}.fai")),
                      output=dict(tsv=P("final_out", f"hap{h}.gaps.covprob.tsv")), config=dict(SUNK_len=k), wildcards=dict(hap=f"hap{h}"))
            rc, out, err = run_script("covprob.py", snakemake=sm)
            case["final_out"][f"hap{h}.gaps.covprob.tsv"] = read(P("final_out", f"hap{h}.gaps.covprob.tsv")) if rc == 0 else None
            case["final_out"][f"hap{h}.covprob_rc"] = rc
            if rc != 0:
                case["final_out"][f"hap{h}.covprob_err"] = err.strip().splitlines()[-1][:200]
    return case


def main():
    only = sys.argv[1:]
    for name, args in (("pystages_a", (211, 20, 2, 90000, 60, 16000)), ("pystages_b", (212, 24, 3, 50000, 50, 14000)),
                       ("pystages_c", (213, 20, 2, 70000, 70, 15000, 0.5))):
        if only and name not in only:
            continue
        c = pipeline_case(*args)
        M.save(name, c)
        print(name, "bad", len(c["bad_sunks"]), "inter", {k: (v or "").count("\n") for k, v in c["inter_outs"].items()},
              "beds", {k: (v or "").count("\n") for k, v in c["bed_files"].items() if v},
              "gaps", {k: (v or "").count("\n") for k, v in c["final_out"].items() if isinstance(v, str)})


if __name__ == "__main__":
    main()
