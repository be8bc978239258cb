"""Minimal stand-in for pyranges -- ONLY what get_gaps.py (PyRanges(df).merge().df) and covprob.py
(PyRanges(df).df) use.  Reproduced behaviour: rows grouped by Chromosome in natural order of the names,
original order inside a chromosome, column order kept; merge() joins overlapping and book-ended intervals
([Start, End) half-open, slack 0) and returns Chromosome / Start / End sorted by Start."""
import re

import pandas as pd


def _natkey(s):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]


class PyRanges:
    def __init__(self, df=None):
        self._df = df.copy() if df is not None else pd.DataFrame(columns=["Chromosome", "Start", "End"])

    @property
    def df(self):
        if len(self._df) == 0:
            return self._df.copy()
        parts = [g for _, g in sorted(self._df.groupby("Chromosome", sort=False, observed=True), key=lambda kv: _natkey(kv[0]))]
        return pd.concat(parts).reset_index(drop=True)

    def merge(self, strand=None, count=False, slack=0):
        rows = []
        for chrom, g in sorted(self._df.groupby("Chromosome", sort=False, observed=True), key=lambda kv: _natkey(kv[0])):
            cur = None
            for s, e in sorted(zip(g["Start"].tolist(), g["End"].tolist())):
                if cur is not None and s <= cur[1] + slack:
                    cur[1] = max(cur[1], e)
                else:
                    if cur is not None:
                        rows.append((chrom, cur[0], cur[1]))
                    cur = [s, e]
            if cur is not None:
                rows.append((chrom, cur[0], cur[1]))
        return PyRanges(pd.DataFrame(rows, columns=["Chromosome", "Start", "End"]))
