"""covprob table kernel (gvs_covprob_table) vs the oracle port of covprob.py:56-100 on the read-length
distributions of the bundled ONT read indexes; tolerance 1e-6 relative (north_star), applied where
the probability is large enough for `1 - prod` not to cancel, 1e-13 absolute below."""
import numpy as np
import pytest

import gavisunk_oracle as O
from conftest import load_golden
from gavisunk_b200 import engine as E


def test_host_scalars_match_oracle_and_survey_anchors():
    anchors = {16: (1.0611898893549694, 0.6869962339684254), 20: (1.0384140075192687, 0.46589071253970926),
               24: (1.0255110516158625, 0.29090839959169), 31: (1.0135620658558933, 0.1050584920141705)}
    for r, (root, pn) in anchors.items():
        assert abs(E.covprob_root(r) - root) / root < 1e-12
        assert abs(E.covprob_pn(r) - pn) / pn < 1e-9
        assert abs(O.covprob_pn(r) - pn) / pn < 1e-9


def test_oracle_table_anchors():
    case = load_golden("covprob_amy")
    t1, t2 = case["hap1"]["table_r20"], case["hap2"]["table_r20"]
    for i, v in {1: 0.9999999999842043, 10: 0.9999999993513834, 39: 0.999994564253644, 100: 0.9884562569379078,
                 200: 0.6541321053984619, 400: 0.05135194402166521}.items():
        assert abs(t1[i] - v) <= 1e-9 * v
    assert t1[600] == 0.0 and abs(t2[400] - 0.02734217554277385) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("hap", ["hap1", "hap2"])
def test_covprob_table_gpu(hap):
    case = load_golden("covprob_amy")[hap]
    rl = [(n, l) for n, l in case["read_lens"]]
    eng = E.Engine(20)
    kbp, cnt = E.covprob_bins(rl)
    got = eng.covprob_table(kbp, cnt, case["genome_kbp"], E.covprob_pn(20))
    exp = np.asarray(case["table_r20"])
    assert got[0] == 1.0
    big = exp > 1e-7
    assert np.max(np.abs(got[big] - exp[big]) / exp[big]) < 1e-6
    assert np.max(np.abs(got[~big] - exp[~big])) < 1e-13
    # per-gap lookup as covprob.py:118,130 does it: table[int(max_gap / 1000)]
    assert abs(got[39] - 0.999994564253644) < 1e-9 or hap == "hap2"


@pytest.mark.gpu
def test_covprob_sweep_k():
    rng = np.random.default_rng(3)
    rl = [(f"r{i}", int(x)) for i, x in enumerate(rng.lognormal(10.2, 0.8, 4000))]
    for r in (16, 24, 31):
        eng = E.Engine(r)
        kbp, cnt = E.covprob_bins(rl)
        got = eng.covprob_table(kbp, cnt, 150000.0, E.covprob_pn(r))
        exp = np.asarray(O.covprob_table(rl, 150000.0, r))
        big = exp > 1e-7
        assert np.max(np.abs(got[big] - exp[big]) / exp[big]) < 1e-6
