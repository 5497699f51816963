"""CPU tests of the ORACLE (test infrastructure): golden pins, SATLIB verdicts, model
validity, Lit/LBool tables of the reference (src/sat/formula.rs:183-261), edge cases."""
import numpy as np
import pytest

from common import EDGE_CASES, csr_of, golden_instance, load_golden, synth_batch, SEED2


def test_lit_encoding_tables(pkg):
    # formula.rs:183-207 test_lits: var/sign round trip, !!l == l, !l flips only the sign
    for v in (0, 1, 5, 1000):
        for sign in (False, True):
            l = pkg.mk_lit(v, sign)
            assert pkg.lit_var(l) == v and pkg.lit_sign(l) == sign
            assert pkg.lit_not(pkg.lit_not(l)) == l
            assert pkg.lit_var(pkg.lit_not(l)) == v and pkg.lit_sign(pkg.lit_not(l)) != sign
    assert pkg.mk_lit(3, False) == 6 and pkg.mk_lit(3, True) == 7


def test_golden_pins(po):
    g = load_golden()
    for i, name in enumerate(g["names"]):
        off, lits = golden_instance(g, i)
        for kind, tag in ((po.CORE, "core"), (po.SIMP, "simp")):
            r, m = po.solve_csr(off, lits, kind=kind)
            assert r.verdict == int(g[tag + "_verdict"][i]), name
            assert list(r.stats) == [int(x) for x in g[tag + "_stats"][i]], (name, tag)
            assert r.trace_hash == int(g[tag + "_hash"][i]), (name, tag)
            if g["name_verdict"][i] >= 0:
                assert r.verdict == int(g["name_verdict"][i]), name    # SATLIB naming: uf SAT, uuf UNSAT
            if r.verdict == po.SAT:
                assert po.check_model(off, lits, m), name


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_edge_cases_consistent(po, case):
    off, lits = csr_of(EDGE_CASES[case])
    verdicts = set()
    for kind in (po.CORE, po.SIMP):
        r, m = po.solve_csr(off, lits, kind=kind)
        verdicts.add(r.verdict)
        if r.verdict == po.SAT:
            assert po.check_model(off, lits, m)
    assert len(verdicts) == 1
    expect_unsat = {"unit_conflict", "empty_clause", "unit_chain_unsat", "php_3_2", "late_unit"}
    assert (verdicts == {po.UNSAT}) == (case in expect_unsat)


def test_synthetic_generator_shape(po):
    lits = po.gen_ksat(SEED2, 100, 426)
    assert lits.shape == (1278,)
    v = (lits >> 1).reshape(426, 3)
    assert v.max() < 100
    assert (v[:, 0] != v[:, 1]).all() and (v[:, 0] != v[:, 2]).all() and (v[:, 1] != v[:, 2]).all()
    assert (po.gen_ksat(SEED2, 100, 426) == lits).all()
    assert (po.gen_ksat(SEED2 + 1, 100, 426) != lits).any()


def test_core_simp_agree_on_synthetic(po):
    ioff, coff, lits = synth_batch(po, SEED2, 60, 60, 256)
    rc, mc, _ = po.solve_batch(ioff, coff, lits, kind=po.CORE, threads=2, model_stride=60)
    rs, ms, _ = po.solve_batch(ioff, coff, lits, kind=po.SIMP, threads=2, model_stride=60)
    for i in range(60):
        assert rc[i].verdict == rs[i].verdict
        cb, ce = ioff[i], ioff[i + 1]
        if rc[i].verdict == po.SAT:
            assert po.check_model(coff[cb:ce + 1], lits, mc[i]) and po.check_model(coff[cb:ce + 1], lits, ms[i])


def test_dimacs_reader_matches_oracle_parser(pkg, po, tmp_path):
    text = "c hello\np cnf 5 4\n1 -2 0\nc mid comment\n 3 +4\n -5 0 2 0\n-1 0\nc % \n"
    p = tmp_path / "t.cnf"
    p.write_text(text)
    off, lits, hv, hc = pkg.dimacs.read_csr(str(p))
    o_off, o_lits, o_hv, o_hc = po.parse_dimacs(str(p))
    assert (off == o_off).all() and (lits == o_lits).all() and (hv, hc) == (o_hv, o_hc) == (5, 4)
    with pytest.raises(IOError):
        pkg.dimacs.read_csr_text("p cnf 1 1\n1 0\n%\n0\n")          # '%' -> "int expected" (dimacs.rs:319-336)
    with pytest.raises(IOError):
        pkg.dimacs.read_csr_text("1 2 0\n")                          # header is mandatory
    with pytest.raises(IOError):
        pkg.dimacs.read_csr_text("p cnf 2 3\n1 2 0\n", validate=True)  # --strict header check
    import gzip
    gz = tmp_path / "t.cnf.gz"
    with gzip.open(gz, "wb") as f:
        f.write(text.encode())
    off2, lits2, _, _ = pkg.dimacs.read_csr(str(gz))
    assert (off2 == off).all() and (lits2 == lits).all()


def test_write_result_format(pkg):
    res = pkg.SolveRes(pkg.SAT, pkg.Stats(), [pkg.mk_lit(0, False), pkg.mk_lit(1, True), pkg.mk_lit(2, False)])
    assert pkg.dimacs.result_string(res) == "SAT\n1 -2 3 0\n"
    assert pkg.dimacs.result_string(pkg.SolveRes(pkg.UNSAT, pkg.Stats())) == "UNSAT\n"
    assert pkg.dimacs.result_string(pkg.SolveRes(pkg.INTERRUPTED, pkg.Stats())) == "INDET\n"
    assert pkg.dimacs.validate_model_text("p cnf 3 2\n1 -2 0\n2 3 0\n", res.model)
    assert not pkg.dimacs.validate_model_text("p cnf 3 1\n-1 2 -3 0\n", res.model)


def test_cxx_dimacs_reader_quirks(pkg, tmp_path):
    """b200sat_read_dimacs (csrc/dimacs_io.cpp) keeps the reference parser's quirks (dimacs.rs:171-365)."""
    def rd(text, **kw):
        p = tmp_path / "q.cnf"
        p.write_text(text)
        return pkg.dimacs.read_csr(str(p), **kw)
    off, lits, hv, hc = rd("c x\nc y\np cnf 3 2\n+1 -2 0 3\n0\n")                 # '+' sign, clause split over lines
    assert list(off) == [0, 2, 3] and list(lits) == [0, 3, 4] and (hv, hc) == (3, 2)
    off, lits, hv, hc = rd("p cnf 1 1\n1 0\n")                                       # header counts are not checked ...
    off2, _, _, _ = rd("p cnf 9 9\n1 0\n")
    assert list(off) == list(off2)
    for bad in ("p cnf 1 1\n1 0\n%\n0\n", "1 2 0\n", "p cnf x 1\n", "p cnf 1 1\n1 a 0\n"):
        with pytest.raises(IOError):
            rd(bad)
    with pytest.raises(IOError):
        rd("p cnf 2 3\n1 2 0\n", validate=True)                                      # ... unless --strict
    with pytest.raises(IOError):
        rd("p cnf 1 1\n1 2 0\n", validate=True)
    with pytest.raises(IOError):
        pkg.dimacs.read_csr(str(tmp_path / "missing.cnf"))
    # text reader and C++ reader agree
    t = "c c\np cnf 4 3\n1 -2 0\n-3 4 -1 0\nc mid\n2 0\n"
    a = rd(t)
    b = pkg.dimacs.read_csr_text(t)
    assert list(a[0]) == list(b[0]) and list(a[1]) == list(b[1])


def test_hand_traces(po):
    """tests/golden/hand_traces.md: the counters derived by hand from the Rust source pin the oracle itself."""
    from common import HAND_TRACES, csr_of
    for name, (clauses, kind, over, verdict, stats, model, elim, ncl) in HAND_TRACES.items():
        st = po.default_settings()
        for k, v in over.items():
            setattr(st, k, v)
        off, lits = csr_of(clauses)
        r, m = po.solve_csr(off, lits, kind=kind, settings=st)
        assert r.verdict == verdict, name
        assert list(r.stats) == stats, (name, list(r.stats))
        assert r.eliminated_vars == elim and r.n_clauses_pre == ncl, name
        if model is not None:
            assert [int(x) for x in m] == model, name


def test_hand_traces_call_semantics(po):
    """H6 / H7 of tests/golden/hand_traces.md: a false assumption, a conflict budget with Interrupted and a second
    solve_limited -- constants derived by hand from the Rust source pin the oracle's round-2 additions."""
    from common import HAND_TRACES_CALLS, HAND_TRACES_PROGRESS, csr_of
    for name, (clauses, assumps, (cb, pb), first, second) in HAND_TRACES_CALLS.items():
        off, lits = csr_of(clauses)
        for rounds, want in ((0, first), (1, second)):
            if want is None:
                continue
            po.set_assumptions(assumps); po.set_budget(cb, pb, rounds)
            try:
                r, m = po.solve_csr(off, lits, kind=po.CORE)
            finally:
                po.set_assumptions(); po.set_budget()
            assert r.verdict == want[0] and list(r.stats) == want[1], (name, rounds, r.verdict, list(r.stats))
            if len(want) > 2:
                assert [int(x) for x in m] == want[2], name
            if rounds == 0 and name in HAND_TRACES_PROGRESS:
                assert r.progress == HAND_TRACES_PROGRESS[name], (name, r.progress)


def test_hand_trace_document_and_fixture_agree():
    """The constants asserted by the hand-trace tests are the ones written in tests/golden/hand_traces.md (the derivation
    is the source of truth; common.py must not drift from it)."""
    import os, re
    from common import HAND_TRACES, HAND_TRACES_CALLS
    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hand_traces.md")).read()
    doc = {}
    for sec in re.split(r"\n## ", text)[1:]:
        stats = [[int(x) for x in t.split(",")] for t in re.findall(r"[Ss]tats `\[([0-9, ]+)\]`", sec)]
        for key in re.findall(r"H\d+", sec.split("\n", 1)[0].split("\u2014")[0]):   # "H13, H14 \u2014 ...": the cases in front of the dash
            doc[key] = stats
    for name, t in HAND_TRACES.items():
        key = name.split("_")[0]
        assert key in doc and t[4] in doc[key], (name, t[4], doc.get(key))
    for name, (_, _, _, first, second) in HAND_TRACES_CALLS.items():
        key = name.split("_")[0]
        assert first[1] in doc[key], name
        if second is not None:
            assert second[1] in doc[key], name

