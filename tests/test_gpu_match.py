"""GPU parity of the match stage (gvs_match) against the reference ELF outputs recorded in
tests/golden (kmerpos_annot3) -- bit-exact rows, including quirks Q1-Q6."""
import numpy as np
import pytest

from conftest import load_golden, engine_from_case, run_match_chunks

pytestmark = pytest.mark.gpu


VARIANTS = [0, 2]  # probe instantiation: by size (= small-database for the golden cases) / large-database forced


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["kat_b1", "kat_bytes", "kat_k32", "kat_k32b"])
def test_match_kats(name, variant):
    case = load_golden(name)
    eng, names = engine_from_case(case, variant=variant)
    got = run_match_chunks(eng, names, [case["reads"]])
    assert got[0] == case["out"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_match_q7_duplicate_loc_rows_and_orphan_db_kmer(variant):
    """SURVEY Q7 against the executable's output (tests/golden/kat_q7): last .loc row of a k-mer wins, a .loc k-mer
    outside the db set never matches, hitting a db k-mer without a .loc row is the reference's KeyError"""
    from gavisunk_b200.engine import GavisunkError
    case = load_golden("kat_q7")
    eng, names = engine_from_case(case, variant=variant)
    assert run_match_chunks(eng, names, [case["reads"]])[0] == case["out"]
    with pytest.raises(KeyError) as ei:  # the reference's exception type; also a GavisunkError (code GVS_E_KEYERROR)
        run_match_chunks(eng, names, [case["reads_keyerror"]])
    assert isinstance(ei.value, GavisunkError) and ei.value.code == -3
    assert "KeyError" in str(ei.value)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("batched", [False, True])
def test_match_every_sunk_len(variant, batched):
    """ksweep: every probe instantiation besides K = 4 / 16 / 20 / 24 / 31 (K = 2, 3 have one window per filter
    group, the others four) and the one-kernel k = 32 path, per chunk file and with all chunk files in one batch"""
    for case in load_golden("ksweep"):
        eng, names = engine_from_case(case, variant=variant)
        if batched:
            got = run_match_chunks(eng, names, [ch["reads"] for ch in case["chunks"]])
            assert got == [ch["sunkpos"] for ch in case["chunks"]], case["k"]
        else:
            for ch in case["chunks"]:
                assert run_match_chunks(eng, names, [ch["reads"]])[0] == ch["sunkpos"], case["k"]
        eng.close()


@pytest.mark.parametrize("name", ["rand_k20", "rand_k16", "rand_k24", "rand_k31", "rand_k20_many", "ragged_k20", "ragged_k31", "ragged_k16"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_match_random_per_chunk(name, variant):
    case = load_golden(name)
    eng, names = engine_from_case(case, variant=variant)
    for ch in case["chunks"]:
        got = run_match_chunks(eng, names, [ch["reads"]])
        assert got[0] == ch["sunkpos"]


@pytest.mark.parametrize("name", ["rand_k20", "rand_k20_many", "ragged_k20", "ragged_k31"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_match_random_batched_chunks(name, variant):
    """all chunk files of a sample in ONE batch: the prevLoc carry must restart per chunk (Q4)"""
    case = load_golden(name)
    eng, names = engine_from_case(case, variant=variant)
    got = run_match_chunks(eng, names, [ch["reads"] for ch in case["chunks"]])
    for g, ch in zip(got, case["chunks"]):
        assert g == ch["sunkpos"]


@pytest.mark.parametrize("name", ["rand_k20", "rand_k31", "rand_k20_many", "kat_bytes", "ragged_k20", "ragged_k16"])
@pytest.mark.parametrize("segments", [2, 7, 64])
@pytest.mark.parametrize("pack", [0, 1, 2, 3])
def test_match_copy_pipeline(name, segments, pack):
    """host batch copied in segments on the copy stream, one probe launch per segment
    (gvs_set_copy_pipeline): rows identical to the reference ELF output whatever the split and whichever
    segments were 2-bit packed by the host threads on their way (gvs_set_host_pack: none / the library's
    choice / all / every other one, i.e. ASCII and packed probe variants side by side in one batch)"""
    case = load_golden(name)
    eng, names = engine_from_case(case, variant=2 if (pack + segments) % 2 else 0)
    eng.set_copy_pipeline(0, segments)
    eng.set_host_pack(pack, 3)
    chunks = [ch["reads"] for ch in case["chunks"]] if "chunks" in case else [case["reads"]]
    want = [ch["sunkpos"] for ch in case["chunks"]] if "chunks" in case else [case["out"]]
    for _ in range(2):  # second round re-uses the segment events and overwrites the resident batch
        got = run_match_chunks(eng, names, chunks)
        assert got == want
        nbytes, nseg, npk = eng.copy_stats()
        if nseg > 1:
            assert npk == {0: 0, 1: nseg, 2: nseg, 3: nseg // 2}[pack]  # 1: pageable source, always packed


@pytest.mark.parametrize("name", ["kat_b1", "kat_bytes", "kat_k32b", "rand_k20", "rand_k16", "rand_k31", "rand_k20_many", "ragged_k20", "ragged_k31"])
@pytest.mark.parametrize("segments", [1, 5])
@pytest.mark.parametrize("variant", VARIANTS)
def test_match_packed_host_reads(name, segments, variant):
    """bases packed to 2 bits on the host (gvs_pack_2bit = kmer.encode's byte map) and matched by the PACKED
    probe variant: rows identical to the reference ELF output, with and without the copy pipeline"""
    from conftest import pack_chunks, format_rows
    from gavisunk_b200.engine import pack_2bit
    case = load_golden(name)
    eng, names = engine_from_case(case, variant=variant)
    eng.set_copy_pipeline(0, segments)
    chunks = [ch["reads"] for ch in case["chunks"]] if "chunks" in case else [case["reads"]]
    want = [ch["sunkpos"] for ch in case["chunks"]] if "chunks" in case else [case["out"]]
    rnames, seq, off, chunk_first = pack_chunks(chunks)
    eng.set_reads_packed(pack_2bit(seq), off, chunk_first, [0] * len(chunks))
    eng.match()
    rows = eng.rows(0)
    got = [format_rows(rows, rnames, names, chunk_first[i], chunk_first[i + 1]) for i in range(len(chunks))]
    assert got == want
