"""CLI front end: flag surface / ranges on the CPU, full run against the oracle on the GPU."""
import io
import os

import numpy as np
import pytest


def _cli(pkg):
    import importlib
    return importlib.import_module("minisat_rust_b200.cli")


def test_flag_surface_and_ranges(pkg):
    cli = _cli(pkg)
    ap = cli.build_parser()
    # every long flag of main.rs:20-54 parses; --rsync (README only) does not exist
    a = ap.parse_args(["--verb", "0", "--core", "--strict", "--no-pre", "--no-solve", "--var-decay", "0.9", "--cla-decay", "0.99",
                       "--rnd-freq", "0.1", "--rnd-seed", "7", "--ccmin-mode", "1", "--phase-saving", "0", "--rnd-init",
                       "--no-luby", "--rfirst", "50", "--rinc", "3", "--gc-frac", "0.3", "--min-learnts", "5", "x.cnf", "out.txt"])
    s = cli.settings_from_args(a)
    assert (s.var_decay, s.clause_decay, s.random_var_freq, s.random_seed) == (0.9, 0.99, 0.1, 7.0)
    assert (s.ccmin_mode, s.rnd_init_act, s.luby_restart, s.restart_first, s.restart_inc) == (1, 1, 0, 50.0, 3.0)
    assert s.phase_saving == 2                       # --phase-saving is a silent no-op in the reference (main.rs:144)
    with pytest.raises(SystemExit):
        ap.parse_args(["--rsync", "x.cnf"])
    # out-of-range values are silently ignored (main.rs:89-213)
    d = pkg.default_settings()
    s = cli.settings_from_args(ap.parse_args(["--var-decay", "1.5", "--rnd-freq", "-1", "--rinc", "0.5", "--gc-frac", "0", "x.cnf"]))
    assert (s.var_decay, s.random_var_freq, s.restart_inc, s.garbage_frac) == (d.var_decay, d.random_var_freq, d.restart_inc, d.garbage_frac)
    # simplifier flags are only read for SimpSolver
    s = cli.settings_from_args(ap.parse_args(["--grow", "4", "--cl-lim", "10", "--sub-lim", "-1", "--no-elim", "x.cnf"]))
    assert (s.grow, s.clause_lim, s.subsumption_lim, s.use_elim) == (4, 10, -1, 0)
    s = cli.settings_from_args(ap.parse_args(["--core", "--grow", "4", "x.cnf"]))
    assert s.grow == 0


@pytest.mark.gpu
@pytest.mark.parametrize("core", [False, True])
def test_cli_run_matches_oracle(pkg, po, tmp_path, core):
    from common import golden_instance, load_golden
    cli = _cli(pkg)
    g = load_golden()
    for name in ("uf50-02.cnf.gz", "uuf50-03.cnf.gz"):
        i = list(g["names"]).index(name)
        off, lits = golden_instance(g, i)
        cnf = tmp_path / (name.replace(".gz", ""))
        with open(cnf, "w") as f:
            f.write("c test\np cnf %d %d\n" % (int(g["n_vars"][i]), len(off) - 1))
            for c in range(len(off) - 1):
                f.write(" ".join(str(-(int(l >> 1) + 1) if l & 1 else int(l >> 1) + 1) for l in lits[off[c]:off[c + 1]]) + " 0\n")
        outp = tmp_path / "res.txt"
        so, se = io.StringIO(), io.StringIO()
        verdict = cli.run((["--core"] if core else []) + [str(cnf), str(outp)], out=so, errout=se)
        tag = "core" if core else "simp"
        assert verdict == int(g[tag + "_verdict"][i])
        assert so.getvalue().strip() == ("SATISFIABLE" if verdict == 1 else "UNSATISFIABLE")
        log = se.getvalue().splitlines()
        st = [int(x) for x in g[tag + "_stats"][i]]
        # the reference's own test reads the value after ':' of these lines (tests/minisat.rs:173-219)
        val = lambda key: int([l for l in log if l.startswith(key)][0].split(":")[1].split()[0])
        assert val("restarts") == st[1] and val("conflicts  ") == st[4] and val("decisions") == st[2]
        assert val("propagations") == st[5] and val("conflict literals") == st[6]
        body = open(outp).read()
        if verdict == 1:
            assert body.startswith("SAT\n") and body.rstrip().endswith(" 0")
            model = [int(x) for x in body.split("\n")[1].split()[:-1]]
            assert pkg.dimacs.validate_model_text(open(cnf).read(), [pkg.mk_lit(abs(x) - 1, x < 0) for x in model])
        else:
            assert body == "UNSAT\n"
