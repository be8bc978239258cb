"""The CLI shims (gavisunk_b200.cli) reproduce the reference programs' files byte for byte: argv and
formats of workflow/scripts/{kmerpos_annot3,rlen,diag_filter_v3,diag_filter_step2} against the ELF
outputs in tests/golden, and the fused rule against the oracle's composition of all stages."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

import gavisunk_oracle as O
from conftest import load_golden
from gavisunk_b200 import cli, io as gio


def test_rlen_shim(tmp_path):
    case = load_golden("rlen_b8")
    p = tmp_path / "reads.fa"
    p.write_bytes(case["reads"].encode("latin-1"))
    assert cli.main(["rlen", str(p), str(tmp_path / "o.rlen")]) == 0
    assert (tmp_path / "o.rlen").read_text() == case["rlen"]


def test_cli_errors_exit_nonzero(tmp_path):
    assert cli.main(["rlen", str(tmp_path / "missing.fa"), str(tmp_path / "o.rlen")]) == 1
    assert not (tmp_path / "o.rlen").exists()
    assert cli.main(["nonsense"]) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat_b1", "rand_k24"])
def test_kmerpos_annot3_shim(tmp_path, name):
    case = load_golden(name)
    (tmp_path / "db.txt").write_text(case["db"])
    (tmp_path / "loc.txt").write_text(case["loc"])
    chunks = [dict(reads=case["reads"], sunkpos=case["out"])] if "reads" in case else case["chunks"][:2]
    for i, ch in enumerate(chunks):
        rp = tmp_path / (f"r{i}.fq" if i % 2 else f"r{i}.fa")
        rp.write_bytes(ch["reads"].encode("latin-1"))
        out = tmp_path / f"o{i}.sunkpos"
        assert cli.main(["kmerpos_annot3", str(rp), str(tmp_path / "db.txt"), str(tmp_path / "loc.txt"), str(out)]) == 0
        assert out.read_text() == ch["sunkpos"]


@pytest.mark.gpu
def test_diag_shims(tmp_path):
    for case in load_golden("diag_cases")[:3]:
        sp, fp, dp = tmp_path / "s.sunkpos", tmp_path / "h.fai", tmp_path / "s.diag"
        sp.write_text(case["sunkpos"])
        fp.write_text(case["fai"])
        buf = io.StringIO()
        with redirect_stdout(buf):
            assert cli.main(["diag_filter_v3", str(sp), str(fp)]) == 0
        assert buf.getvalue() == case["diag"], case["name"]
        dp.write_text(case["diag"])
        buf = io.StringIO()
        with redirect_stdout(buf):
            assert cli.main(["diag_filter_step2", str(sp), str(dp)]) == 0
        assert buf.getvalue() == case["diag2"], case["name"]


def _oracle_pipeline(contigs_h, reads_h, k):
    """oracle composition of every stage; returns the file contents the fused rule must produce"""
    contigs = contigs_h[0] + contigs_h[1]
    names = [n for n, _ in contigs]
    db = O.build_sunk_db(contigs, k)
    loc = [(names[c], int(s), int(km), int(g)) for c, s, km, g in zip(db["contig"], db["start"], db["kmer"], db["group"])]
    out = {"kmer.loc": "".join(f"{c}\t{s}\t{O.decode(km, k)}\t{g}\n" for c, s, km, g in loc)}
    kept_h, rlen = [], {}
    for hap in range(2):
        hapc = {n for n, _ in contigs_h[hap]}
        kept = []
        for chunk in reads_h[hap]:
            rows = O.match_chunk(chunk, db["kmer"], loc, k)
            kept += O.diag_filter_step2(rows, O.diag_filter_v3(rows, hapc))
            for n, s in chunk:
                rlen[n] = len(s)
        kept_h.append(kept)
        out[f"hap{hap + 1}.sunkpos"] = gio.format_sunkpos(kept)
    bad = O.bad_sunks(kept_h[0], {n for n, _ in contigs_h[0]}, kept_h[1], {n for n, _ in contigs_h[1]})
    out["bad"] = {f"{c}:{g}" for c, g in bad}
    for hap in range(2):
        byc, beds, valid = {}, {}, []
        for r in kept_h[hap]:
            byc.setdefault(r[2], []).append(r)
        for ctg, rr in byc.items():
            inter, bed = O.process_by_contig(rr, rlen, bad, ctg)
            out[f"inter:{ctg}"] = "".join(f"{g}\t{n}\n" for g, n in inter) if inter else ctg + "\n"
            if bed is not None:
                beds[ctg] = [(s, e) for _, s, e in bed]
                valid += [f"{c}\t{s}\t{e}\n" for c, s, e in bed]
        fai = [(n, len(s)) for n, s in contigs_h[hap]]
        gaps, nodata = O.get_gaps(fai, beds)
        out[f"hap{hap + 1}.valid.bed"] = sorted(valid)
        out[f"hap{hap + 1}.gaps.bed"] = "".join(f"{c}\t{s}\t{e}\n" for c, s, e in gaps)
        out[f"hap{hap + 1}.nodata.bed"] = "".join(f"{c}\t{s}\t{e}\n" for c, s, e in nodata)
        out[f"hap{hap + 1}.gaps.slop.bed"] = "".join(f"{c}\t{s}\t{e}\n" for c, s, e in O.slop_gaps(gaps, dict(fai)))
    return out


@pytest.fixture(scope="module")
def fused_case(tmp_path_factory):
    """one small diploid sample run through `cli fused` + the oracle's composition of every stage"""
    tmp_path = tmp_path_factory.mktemp("fused")
    from gavisunk_b200.engine import Engine
    from gavisunk_b200 import workload as W
    k = 20
    eng = Engine(k)
    wl = W.make_assembly(eng, [260000, 90000, 4000], snp_rate=2e-3, dup_frac=0.05, seed=91)
    W.add_reads(eng, wl, coverage=16.0, n50=16000, sigma=0.5, len_min=300, len_max=200000, seed=92, nchunks=2)
    asm = wl.asm.cpu().numpy()
    coff = wl.contig_off.cpu().numpy()
    nc = len(wl.contig_names) // 2
    contigs_h = [[(wl.contig_names[c], asm[coff[c]:coff[c + 1]].tobytes()) for c in range(h * nc, (h + 1) * nc)] for h in range(2)]
    reads = wl.reads.cpu().numpy()
    off = wl.read_off.cpu().numpy()
    cf = wl.chunk_first
    reads_h, files_h = [[], []], [[], []]
    for ci in range(len(cf) - 1):
        hap = int(wl.chunk_hap[ci])
        chunk = [(f"read{r:07d}", reads[off[r]:off[r + 1]].tobytes()) for r in range(int(cf[ci]), int(cf[ci + 1]))]
        reads_h[hap].append(chunk)
        fp = tmp_path / f"hap{hap + 1}_{ci}.fa"
        fp.write_bytes(b"".join(b">" + n.encode() + b"\n" + s + b"\n" for n, s in chunk))
        files_h[hap].append(str(fp))
    asm_files, fai_files = [], []
    for hap in range(2):
        fp = tmp_path / f"hap{hap + 1}.fa"
        fp.write_bytes(b"".join(b">" + n.encode() + b"\n" + s + b"\n" for n, s in contigs_h[hap]))
        asm_files.append(str(fp))
        fai = tmp_path / f"hap{hap + 1}.fa.fai"
        fai.write_text("".join(f"{n}\t{len(s)}\t0\t60\t61\n" for n, s in contigs_h[hap]))
        fai_files.append(str(fai))
    outdir = tmp_path / "results"
    rc = cli.main(["fused", "--k", str(k), "--hap1-asm", asm_files[0], "--hap2-asm", asm_files[1], "--hap1-reads", *files_h[0],
                   "--hap2-reads", *files_h[1], "--outdir", str(outdir)])
    assert rc == 0
    exp = _oracle_pipeline(contigs_h, reads_h, k)
    return dict(outdir=str(outdir), exp=exp, contigs_h=contigs_h, fai=fai_files, k=k, tmp=tmp_path)


@pytest.mark.gpu
def test_fused_rule_files(fused_case):
    outdir, exp = fused_case["outdir"], fused_case["exp"]
    rd = lambda *p: open(os.path.join(outdir, *p)).read()
    assert rd("mrsfast", "kmer.loc") == exp["kmer.loc"]
    assert sorted(rd("db", "jellyfish.db").split()) == sorted(l.split("\t")[2] for l in exp["kmer.loc"].splitlines())
    assert set(rd("sunkpos", "bad_sunks.txt").split()) == exp["bad"]
    n_inter = 0
    for hap in (1, 2):
        assert rd("sunkpos", f"hap{hap}.sunkpos") == exp[f"hap{hap}.sunkpos"]
        assert sorted(rd("final_out", f"hap{hap}.valid.bed").splitlines(True)) == exp[f"hap{hap}.valid.bed"]
        for f in ("gaps.bed", "nodata.bed", "gaps.slop.bed"):
            assert rd("final_out", f"hap{hap}.{f}") == exp[f"hap{hap}.{f}"], f
    for key, val in exp.items():
        if key.startswith("inter:"):
            ctg = key[6:]
            hap = 1 if ctg.startswith("synH1") else 2
            assert rd("inter_outs", f"{ctg}_hap{hap}.tsv") == val, ctg
            n_inter += val.count("\n")
    assert n_inter > 100


@pytest.mark.gpu
def test_rule_shims_chain(fused_case):
    """the per-rule shims (badsunks_AR, split_locs, process_by_contig, get_gaps, slop_gaps, covprob) chained
    through files exactly as workflow/rules/tagONT.smk:132-250 chains the reference scripts"""
    outdir, exp, fai, k = fused_case["outdir"], fused_case["exp"], fused_case["fai"], fused_case["k"]
    work = fused_case["tmp"] / "rules"
    for sub in ("sunkpos", "breaks", "inter_outs", "bed_files", "final_out"):
        os.makedirs(work / sub, exist_ok=True)
    src = lambda *p: os.path.join(outdir, *p)
    # rule bad_sunks
    assert cli.main(["badsunks_AR", fai[0], fai[1], src("sunkpos", "hap1.sunkpos"), src("sunkpos", "hap2.sunkpos"),
                     str(work / "sunkpos" / "bad_sunks.txt")]) == 0
    assert set((work / "sunkpos" / "bad_sunks.txt").read_text().split()) == exp["bad"]
    # checkpoint split_sunkpos + rule process_by_contig
    rlen_all = work / "sunkpos" / "all.rlen"
    n_bed = 0
    for hap in (1, 2):
        flag = work / "breaks" / f"hap{hap}_splits_pos.done"
        assert cli.main(["split_locs", "--ont-pos", src("sunkpos", f"hap{hap}.sunkpos"), "--kmer-loc", src("mrsfast", "kmer.loc"),
                         "--flag", str(flag), "--hap", f"hap{hap}"]) == 0
        assert flag.exists()
        for fn in sorted(os.listdir(work / "breaks")):
            if not fn.endswith(f"_hap{hap}.sunkpos"):
                continue
            ctg = fn[:-len(f"_hap{hap}.sunkpos")]
            assert open(work / "breaks" / fn).read() == open(src("breaks", fn)).read()
            assert open(work / "breaks" / f"{ctg}_hap{hap}.loc").read() == open(src("breaks", f"{ctg}_hap{hap}.loc")).read()
            tsv, bed = work / "inter_outs" / f"{ctg}_hap{hap}.tsv", work / "bed_files" / f"{ctg}_hap{hap}.bed"
            assert cli.main(["process_by_contig", str(work / "breaks" / f"{ctg}_hap{hap}.loc"), str(work / "breaks" / fn),
                             src("sunkpos", f"hap{hap}.rlen"), str(work / "sunkpos" / "bad_sunks.txt"), str(tsv), str(bed)]) == 0
            assert tsv.read_text() == exp[f"inter:{ctg}"], ctg
            if bed.exists():
                n_bed += 1
                want = sorted(l for l in exp[f"hap{hap}.valid.bed"] if l.split("\t")[0] == ctg)
                assert sorted(bed.read_text().splitlines(True)) == want
                assert bed.read_text() == open(src("bed_files", f"{ctg}_hap{hap}.bed")).read()
            else:
                bed.write_text("")  # `touch {output.bed}` (tagONT.smk:190)
    assert n_bed >= 2
    # rule get_gaps (paths are concatenated: trailing slashes) + slop_gaps
    assert cli.main(["get_gaps", fai[0], fai[1], "sample", str(work / "bed_files") + "/", str(work / "final_out") + "/"]) == 0
    for hap in (1, 2):
        for f in ("gaps.bed", "nodata.bed"):
            assert (work / "final_out" / f"hap{hap}.{f}").read_text() == exp[f"hap{hap}.{f}"], f
        assert cli.main(["slop_gaps", str(work / "final_out" / f"hap{hap}.gaps.bed"), fai[hap - 1],
                         str(work / "final_out" / f"hap{hap}.gaps.slop.bed")]) == 0
        assert (work / "final_out" / f"hap{hap}.gaps.slop.bed").read_text() == exp[f"hap{hap}.gaps.slop.bed"]
    # rule cov_prob (covprob.py): table + per-gap lookup against the oracle, 1e-6 relative
    n_gap_rows = 0
    for hap in (1, 2):
        gaps_p = work / "final_out" / f"hap{hap}.gaps.bed"
        tsv = work / "final_out" / f"hap{hap}.gaps.covprob.tsv"
        gaps = [(l.split("\t")[0], int(l.split("\t")[1]), int(l.split("\t")[2])) for l in gaps_p.read_text().splitlines()]
        rc = cli.main(["covprob", "--bed", str(gaps_p), "--locs", src("mrsfast", "kmer.loc"), "--rlen", src("sunkpos", f"hap{hap}.rlen"),
                       "--fai", fai[hap - 1], "--sunk-len", str(k), "--tsv", str(tsv)])
        contig, start, kmer, group = gio.read_loc(src("mrsfast", "kmer.loc"))
        loc_rows = list(zip(contig, start, kmer, group))
        rl = [(l.split("\t")[0], int(l.split("\t")[1])) for l in open(src("sunkpos", f"hap{hap}.rlen")).read().splitlines()]
        table = O.covprob_table(rl, sum(l for _, l in gio.read_fai(fai[hap - 1])) / 1000, k)
        try:
            want = O.covprob_gaps(gaps, loc_rows, table)
        except ValueError:
            assert rc == 1  # a gap without a SUNK group in range kills the reference as well
            continue
        assert rc == 0
        lines = tsv.read_text().splitlines()
        assert lines[0] == "index\tChromosome\tStart\tEnd\ttype\tmax_gap\tcovprob"
        got = {}
        for l in lines[1:]:
            i, c, s, e, t, mg, p = l.split("\t")
            assert t == "gap"
            got[int(i)] = (c, int(s), int(e), int(mg), float(p))
        assert len(got) == len(want)
        for i, (c, s, e, mg, p) in enumerate(want):
            assert got[i][:4] == (c, s, e, mg)
            assert got[i][4] == pytest.approx(p, rel=1e-6, abs=1e-15)  # tolerance of BASELINE.json's north_star
            n_gap_rows += 1


@pytest.mark.gpu
def test_fused_palindromic_sunks_option(tmp_path):
    """a SUNK that is its own reverse complement (even SUNK_len) would be reported by mrsfast on both
    strands, i.e. twice in kmer.loc.  Unpinned (mrsfast is absent): one row by default, two with --mrsfast-palindromes;
    nothing downstream changes (every consumer de-duplicates)."""
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    s1 = alpha[rng.integers(0, 4, 6000)].copy()
    s1[1000:1008] = np.frombuffer(b"ACGTACGT", np.uint8)  # its own reverse complement
    s2 = s1.copy()
    s2[1004] = ord("C")  # (ACGT-A-CGT -> ACGT-C-CGT: the palindrome exists in haplotype 1 only, once)
    (tmp_path / "h1.fa").write_bytes(b">c1\n" + s1.tobytes() + b"\n")
    (tmp_path / "h2.fa").write_bytes(b">c2\n" + s2.tobytes() + b"\n")
    reads = b"".join(b">r%d\n" % i + s1[i * 500:i * 500 + 3000].tobytes() + b"\n" for i in range(6))
    (tmp_path / "r1.fa").write_bytes(reads)
    (tmp_path / "r2.fa").write_bytes(reads.replace(b">r", b">q"))
    outs = {}
    for flag in (False, True):
        out = tmp_path / ("pal" if flag else "plain")
        argv = ["fused", "--k", "8", "--hap1-asm", str(tmp_path / "h1.fa"), "--hap2-asm", str(tmp_path / "h2.fa"), "--hap1-reads",
                str(tmp_path / "r1.fa"), "--hap2-reads", str(tmp_path / "r2.fa"), "--outdir", str(out)] + (["--mrsfast-palindromes"] if flag else [])
        assert cli.main(argv) == 0
        outs[flag] = out
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    loc0 = (outs[False] / "mrsfast" / "kmer.loc").read_text().splitlines()
    loc1 = (outs[True] / "mrsfast" / "kmer.loc").read_text().splitlines()
    pal = [l for l in loc0 if l.split("\t")[2] == "".join(comp[c] for c in reversed(l.split("\t")[2]))]
    assert len(loc0) == len(set(loc0)) and len(pal) >= 1  # default: one row per SUNK
    exp = []
    for l in loc0:
        exp += [l, l] if l in pal else [l]
    assert loc1 == exp  # palindromes twice, in place
    for f in ("sunkpos/hap1.sunkpos", "sunkpos/hap2.sunkpos", "final_out/hap1.valid.bed", "final_out/hap1.gaps.bed", "db/jellyfish.db"):
        assert (outs[False] / f).read_text() == (outs[True] / f).read_text(), f
