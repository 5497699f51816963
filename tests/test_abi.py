"""The C-ABI library loads on a CPU-only box, exports every symbol include/b200sat.h
declares, and refuses to solve without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest


def header_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "b200sat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sat_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert header_symbols() == sorted(pkg.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    for s in header_symbols():
        assert hasattr(L, s), s


def test_struct_layouts_match_oracle(pkg, po):
    assert C.sizeof(pkg.Settings) == C.sizeof(po.Settings)
    assert [f[0] for f in pkg.Settings._fields_] == [f[0] for f in po.Settings._fields_]
    a, b = pkg.default_settings(), po.default_settings()
    assert bytes(a) == bytes(b)


def test_no_device_is_a_loud_error(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    assert pkg.lib().b200sat_device_count() == pkg.E_NO_DEVICE
    with pytest.raises(pkg.EngineError):
        pkg.Engine(0)
    s = pkg.CoreSolver()
    a = s.new_var()
    b = s.new_var()
    assert s.add_clause([pkg.mk_lit(a, False), pkg.mk_lit(b, True)])
    with pytest.raises(pkg.EngineError):
        s.solve_limited()


def test_invalid_settings_rejected(pkg):
    st = pkg.default_settings()
    st.var_decay = 0.0
    with pytest.raises(pkg.EngineError):
        pkg.CoreSolver(st)


def test_shard_range(pkg):
    for n, w in ((10, 3), (100000, 8), (7, 8), (0, 2)):
        ranges = [pkg.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert b == c
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
