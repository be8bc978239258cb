"""Host-side file plumbing of the CLI shims (no GPU): error behaviour at the reference's boundary."""
import gzip

import pytest

from gavisunk_b200 import cli, io as gio


def test_read_sunkpos_needs_an_existing_file(tmp_path):
    """A missing / mistyped .sunkpos path must raise, never parse as zero rows (the reference's Nim tools quit
    with 'cannot open the file' / OSError and exit 1, pandas raises FileNotFoundError)."""
    with pytest.raises(FileNotFoundError):
        gio.read_sunkpos(str(tmp_path / "nope.sunkpos"))
    with pytest.raises(FileNotFoundError):
        gio.read_sunkpos("r1\t5\tc1\t100\t100\n")  # sunkpos TEXT is not a path: parse_sunkpos() is for text
    assert gio.parse_sunkpos("r1\t5\tc1\t100\t100\n") == [("r1", 5, "c1", 100, 100)]
    p = tmp_path / "a.sunkpos"
    p.write_text("r1\t5\tc1\t100\t100\nr1\t9\tc1\t300\t300\n")
    assert gio.read_sunkpos(str(p)) == [("r1", 5, "c1", 100, 100), ("r1", 9, "c1", 300, 300)]
    with gzip.open(tmp_path / "a.sunkpos.gz", "wt") as f:
        f.write("r2\t1\tc2\t7\t7\n")
    assert gio.read_sunkpos(str(tmp_path / "a.sunkpos.gz")) == [("r2", 1, "c2", 7, 7)]


@pytest.mark.parametrize("cmd", ["diag_filter_v3", "diag_filter_step2"])
def test_shims_exit_nonzero_on_missing_input(tmp_path, cmd, capsys):
    """exit 1 + message, nothing on stdout (diag_filter_v3 / diag_filter_step2 of the reference do the same)"""
    other = tmp_path / "x"
    other.write_text("c1\t1000\n")
    assert cli.main([cmd, str(tmp_path / "missing.sunkpos"), str(other)]) == 1
    out = capsys.readouterr()
    assert out.out == "" and "missing.sunkpos" in out.err


def test_badsunks_and_split_locs_missing_input(tmp_path):
    fai = tmp_path / "a.fai"
    fai.write_text("c1\t1000\n")
    out = tmp_path / "bad.txt"
    assert cli.main(["badsunks_AR", str(fai), str(fai), str(tmp_path / "m1"), str(tmp_path / "m2"), str(out)]) == 1
    assert not out.exists()
    flag = tmp_path / "breaks" / "hap1_splits_pos.done"
    loc = tmp_path / "k.loc"
    loc.write_text("c1\t1\tAAAA\t1\n")
    assert cli.main(["split_locs", "--ont-pos", str(tmp_path / "m1"), "--kmer-loc", str(loc), "--flag", str(flag), "--hap", "hap1"]) == 1
    assert not flag.exists()
