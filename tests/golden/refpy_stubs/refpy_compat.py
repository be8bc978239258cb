"""pandas 1.3.4 / numpy 1.20 behaviours (envs/viz.yaml:15-16) that the reference scripts rely on and that the
installed pandas 3 / numpy 2 changed or removed.  Imported before a reference script is run by
tests/golden/make_golden_py.py; touches nothing outside that subprocess."""
import numpy as np
import pandas as pd

_vc = pd.Series.value_counts


def _value_counts_13(self, *a, **k):
    # pandas < 2: the result carries the NAME of the counted series and an unnamed index
    # (badsunks_AR.py:24-27 renames columns on that basis); order: counts descending, ties in order of
    # first appearance for object dtype (hash-table order of pandas 1.3 value_counts + stable sort)
    res = _vc(self, *a, **k)
    res.name = self.name
    res.index.name = None
    return res


pd.Series.value_counts = _value_counts_13
if not hasattr(pd.Series, "iteritems"):
    pd.Series.iteritems = pd.Series.items  # covprob.py:90

_split = pd.core.strings.accessor.StringMethods.split


def _split_positional(self, pat=None, n=-1, expand=False, **k):  # badsunks_AR.py:29 passes n positionally
    return _split(self, pat=pat, n=n, expand=expand, **k)


pd.core.strings.accessor.StringMethods.split = _split_positional
if not hasattr(np, "row_stack"):
    np.row_stack = np.vstack  # process-by-contig_lowmem_AR.py:157

_gb_iter = pd.core.groupby.generic.DataFrameGroupBy.__iter__


def _gb_iter_13(self):  # pandas < 2: groupby(['col']) iterates scalar keys, not 1-tuples (split_locs.py:6-7)
    for key, grp in _gb_iter(self):
        if isinstance(key, tuple) and len(key) == 1:
            key = key[0]
        yield key, grp


pd.core.groupby.generic.DataFrameGroupBy.__iter__ = _gb_iter_13
