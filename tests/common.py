"""Shared helpers for the test-suite."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "corpus_sample.npz")
SEED2 = 0x5EED000000000000   # config 2 generator seed base
SEED3 = 0x5EED000100000000   # config 3


def load_golden():
    g = np.load(GOLDEN)
    return {k: g[k] for k in g.files}


def golden_instance(g, i):
    cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
    off = g["clause_off"][cb:ce + 1].astype(np.int64)
    lits = g["lits"][off[0]:off[-1]]
    return (off - off[0]).astype(np.uint32), lits


def synth_batch(po, seed_base, n_inst, n, m):
    lits = np.concatenate([po.gen_ksat(seed_base + i, n, m) for i in range(n_inst)])
    coff = np.arange(0, n_inst * m + 1, dtype=np.uint32) * 3
    ioff = np.arange(0, n_inst + 1, dtype=np.uint32) * m
    return ioff, coff, lits


def csr_of(clauses):
    """list of clauses (signed DIMACS ints) -> (clause_off, lits) Lit-encoded."""
    off, lits = [0], []
    for c in clauses:
        for x in c:
            lits.append(((abs(x) - 1) << 1) | (1 if x < 0 else 0))
        off.append(len(lits))
    return np.array(off, np.uint32), np.array(lits, np.uint32)


def batch_of(insts):
    ioff, coffs, alll, base = [0], [np.zeros(1, np.uint32)], [], 0
    for off, lits in insts:
        coffs.append((off[1:].astype(np.uint64) + base).astype(np.uint32))
        alll.append(lits)
        base += int(off[-1])
        ioff.append(ioff[-1] + len(off) - 1)
    lits = np.concatenate(alll) if alll else np.zeros(0, np.uint32)
    return np.array(ioff, np.uint32), np.concatenate(coffs), lits.astype(np.uint32)


# small hand-made cases for the add_clause branches the bundled corpus never exercises
# (empty clause, duplicate literals, tautologies, ingest-time units, var gaps)
EDGE_CASES = {
    "empty_formula": [],
    "single_unit": [[1]],
    "unit_conflict": [[1], [-1]],
    "empty_clause": [[1, 2], [], [3]],
    "dup_literals": [[1, 1, 2], [-1, -1], [2, 2, 2, 3]],
    "tautology": [[1, -1], [2, -2, 3], [3, 4], [-3, 4], [-4, 1]],
    "unit_chain": [[1], [-1, 2], [-2, 3], [-3, 4], [-4, 5, 6], [-5, -6], [5, 7]],
    "unit_chain_unsat": [[1], [-1, 2], [-2, 3], [-3, -1]],
    "var_gaps": [[1, 5], [-5, 9], [-1, -9], [12, 1]],
    "late_unit": [[1, 2, 3], [-1, 2], [-2, 3], [-3], [1, 2, 4]],
    "long_clause": [list(range(1, 41)), [-1], [-2], [-3, 40], [-40, 39], list(range(-40, -20)), [5, 6, 7]],
    "php_3_2": [[1, 2], [3, 4], [5, 6], [-1, -3], [-1, -5], [-3, -5], [-2, -4], [-2, -6], [-4, -6]],
    "binary_only": [[1, 2], [-1, 2], [1, -2], [3, 4], [-3, -4], [2, 3]],
}


def check_models(ioff, coff, lits, verdicts, models):
    """Vectorised: every SAT instance's model satisfies every one of its clauses.  Returns #violations."""
    n_inst = len(ioff) - 1
    n_cl = len(coff) - 1
    inst_of_clause = np.repeat(np.arange(n_inst), np.diff(ioff.astype(np.int64)))
    clause_of_lit = np.repeat(np.arange(n_cl), np.diff(coff.astype(np.int64)))
    li = lits[:int(coff[-1])].astype(np.int64)
    inst_of_lit = inst_of_clause[clause_of_lit]
    val = models[inst_of_lit, li >> 1]
    lit_true = (val != 2) & ((val == 1) == ((li & 1) == 0))
    sat_clause = np.zeros(n_cl, dtype=bool)
    np.logical_or.at(sat_clause, clause_of_lit, lit_true)
    bad = (~sat_clause) & (verdicts[inst_of_clause] == 1)
    return int(bad.sum())


# tests/golden/hand_traces.md: stats derived BY HAND from the Rust source (not from oracle.cpp) for five tiny formulas.
# name -> (clauses, solver kind (0 core / 1 simp), settings overrides, verdict, 8 stats, model (0/1 per var) or None,
#          eliminated_vars, n_clauses after preprocessing)
HAND_TRACES = {
    "H1_unit_chain": ([[1], [-1, 2], [-2, 3]], 0, {}, 1, [1, 1, 1, 0, 0, 3, 0, 0], [1, 1, 1], 0, 0),
    "H2_two_conflicts_unsat": ([[1, 2], [1, -2], [-1, 2], [-1, -2]], 0, {}, 0, [1, 1, 1, 0, 2, 2, 1, 0], None, 0, 4),
    "H3_binary_learnt": ([[1, 3, 2], [1, 3, -2]], 0, {}, 1, [1, 1, 4, 0, 1, 4, 2, 0], [0, 1, 1], 0, 2),
    "H4_luby_restart": ([[1, 3, 2], [1, 3, -2]], 0, {"restart_first": 1.0}, 1, [1, 2, 5, 0, 1, 6, 2, 0], [0, 1, 1], 0, 2),
    "H5_simp_elimination": ([[1, 2]], 1, {}, 1, [1, 1, 1, 0, 0, 0, 0, 0], [1, 0], 2, 0),
    "H8_simp_subsume_strengthen_unit": ([[1, 2], [1, 2, 3], [-1, 2]], 1, {}, 1, [1, 1, 1, 0, 0, 1, 0, 0], [0, 1, 0], 2, 0),
    "H9_simp_resolvent_and_heap_order": ([[1, 2], [-1, 3], [-2, -3], [2, 3]], 1, {}, 1, [1, 1, 1, 0, 0, 0, 0, 0], [0, 1, 0], 3, 0),
    "H10_deep_minimisation": ([[1, 2], [1, 3, 4], [-2, -3, 4]], 0, {}, 1, [1, 1, 4, 0, 1, 5, 2, 1], [0, 1, 1, 1], 0, 3),
    "H11_random_decisions": ([[1, 2, 3]], 0, {"random_var_freq": 1.0}, 1, [1, 1, 3, 1, 0, 3, 0, 0], [0, 1, 0], 0, 1),
    "H12_reduce_db_removes_the_learnt": ([[1, 2, 3, 4], [1, -2, 3, 4]], 0, {"restart_first": 1.0}, 1, [1, 2, 7, 0, 1, 8, 3, 0], [0, 1, 1, 0], 0, 2),
    "H13_basic_minimisation": ([[1, 2], [1, 3, 4], [-2, -3, 4]], 0, {"ccmin_mode": 1}, 1, [1, 1, 4, 0, 1, 5, 2, 1], [0, 1, 1, 1], 0, 3),
    "H14_no_minimisation": ([[1, 2], [1, 3, 4], [-2, -3, 4]], 0, {"ccmin_mode": 0}, 1, [1, 1, 4, 0, 1, 5, 3, 0], [0, 1, 1, 1], 0, 3),
}

# H6 / H7 of tests/golden/hand_traces.md: solve_limited(budget, assumptions) semantics, derived by hand from the Rust source.
# name -> (clauses, assumptions (Lit encoding), (conflict_budget, propagation_budget),
#          first call (verdict, stats), second call without budget (verdict, stats, model) or None)
HAND_TRACES_CALLS = {
    "H6_false_assumption": ([[1, 2], [-1, 2]], [3], (-1, -1), (0, [1, 1, 0, 0, 1, 2, 1, 0]), None),
    "H7_conflict_budget_then_resume": ([[1, 3, 2], [1, 3, -2]], [], (1, -1), (2, [1, 1, 2, 0, 1, 3, 2, 0]),
                                       (1, [2, 2, 5, 0, 1, 6, 2, 0], [0, 1, 1])),
    "H15_propagation_budget_then_resume": ([[1, 3, 2], [1, 3, -2]], [], (-1, 1), (2, [1, 1, 1, 0, 0, 1, 0, 0]),
                                           (1, [2, 2, 5, 0, 1, 5, 2, 0], [1, 1, 0])),
}
# the f64 carried by SolveRes::Interrupted after the FIRST call (search/util.rs:5-14; derivation in hand_traces.md H7)
HAND_TRACES_PROGRESS = {"H7_conflict_budget_then_resume": ((1.0 / 3.0) * (1.0 / 3.0)) * 2.0,
                        "H15_propagation_budget_then_resume": ((1.0 / 3.0) * (1.0 / 3.0)) * 1.0}

