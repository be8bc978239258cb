"""The C-ABI library loads on a CPU-only box and exports every symbol include/gavisunk_b200.h
declares; the ctypes prototype table covers the same set.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gavisunk_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gvs_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from gavisunk_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gavisunk_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gavisunk_b200.engine import Engine, GavisunkError
    with pytest.raises(GavisunkError):
        Engine(20)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gavisunk_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "gavisunk_oracle" not in src and "oracle/" not in src, fn


def test_stage_ids_match_header():
    """gavisunk_b200._lib.STAGES (names used by Engine.stage_ms and bench.py) are the header's GVS_ST_* ids"""
    from gavisunk_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "gavisunk_b200.h")).read()
    ids = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"\bGVS_ST_([A-Z]+)\s*=\s*(\d+)", txt)}
    count = ids.pop("count")
    assert ids == _lib.STAGES
    assert max(ids.values()) < count
