"""The two covprob host scalars (root of the run-length polynomial, pn) of gavisunk_b200.engine against sympy's
solveset used EXACTLY as workflow/scripts/covprob.py:63-81 uses it -- an independent solver, for every SUNK_len
the engine accepts from 8 to 31 (the oracle's covprob_root is the same bisection as the engine's, so agreeing
with it would be a self-comparison).  No GPU."""
import pytest

from gavisunk_b200 import engine as E


def _reference_root_pn(r):
    from sympy import Reals, Symbol, solveset
    x = Symbol("x")
    p = 0.94
    q = 1 - p
    roots2 = solveset(1 - x + q * ((p) ** (r)) * ((x) ** (r + 1)), x, domain=Reals)
    pinv = 1 / p
    roots2 = [x for x in roots2 if x > 0]
    roots2.sort(key=lambda e: abs(pinv - e))
    assert len(roots2) == 2  # covprob.py:72
    roots3 = roots2[1]
    n = 30
    qn = ((1 - p * roots3) / (q * (r + 1 - r * roots3))) * (1 / (roots3 ** (n + 1)))
    return float(roots3), float(1 - qn)


@pytest.mark.parametrize("r", list(range(8, 32)))
def test_root_and_pn_match_sympy(r):
    root, pn = _reference_root_pn(r)
    got_root, got_pn = E.covprob_root(r), E.covprob_pn(r)
    assert abs(got_root - root) <= 1e-9 * abs(root), (r, got_root, root)
    # pn feeds covprob through pn**2: 1e-7 relative here keeps the table inside the 1e-6 bar
    assert abs(got_pn - pn) <= 1e-7 * max(abs(pn), 1e-12), (r, got_pn, pn)
