"""DIMACS(.gz) reader / result writer / model validator -- host-side IO mirroring
the reference's sat/dimacs.rs (paths relative to /root/reference/src):

  parse_file   dimacs.rs:16-32    gz sniffing, then `parse`
  parse        dimacs.rs:35-43    character-level parser dimacs.rs:171-365 driving
                                  Solver::new_var / add_clause through `Subst` (dimacs.rs:131-168)
  write_result dimacs.rs:46-70    "UNSAT" | "INDET" | "SAT\\n<lits> 0"
  validate_model dimacs.rs:90-128

Quirks kept on purpose: comment lines are recognised only at token start, the
`p cnf` header is mandatory, header counts are ignored unless `validate`
(``--strict``), variables are created lazily in id order up to the largest id seen
so far, `%` is a parse error ("int expected").

`read_csr` produces the CSR arrays the batch engine ingests; the device replays
exactly the add_clause sequence the reference parser would issue.
"""
import gzip
import io
import os

import numpy as np

from . import mk_lit, lit_sign, lit_var, SAT, UNSAT


class ParseError(IOError):
    pass


def _read_text(path):
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:
            return f.read().decode("utf-8")
    with open(path, "rb") as f:
        return f.read().decode("utf-8")


class _Parser:
    """dimacs.rs:171-365"""

    def __init__(self, text):
        self.s = text
        self.n = len(text)
        self.pos = 0
        self.vars = set()
        self.clauses = 0

    def cur(self):
        return self.s[self.pos] if self.pos < self.n else None

    def skip_whitespace(self):
        s, n, p = self.s, self.n, self.pos
        while p < n and s[p].isspace():
            p += 1
        self.pos = p

    def skip_line(self):
        i = self.s.find("\n", self.pos)
        self.pos = self.n if i < 0 else i + 1

    def consume(self, target):
        for tc in target:
            if self.cur() == tc:
                self.pos += 1
            else:
                raise ParseError("failed to consume; expected '%s'" % target)

    def read_int_body(self):
        s, n, p = self.s, self.n, self.pos
        q = p
        while q < n and "0" <= s[q] <= "9":
            q += 1
        if q == p:
            raise ParseError("int expected")
        self.pos = q
        return int(s[p:q])

    def next_int(self):
        self.skip_whitespace()
        sign = 1
        c = self.cur()
        if c == "+":
            self.pos += 1
        elif c == "-":
            self.pos += 1
            sign = -1
        return sign * self.read_int_body()

    def next_uint(self):
        self.skip_whitespace()
        if self.cur() == "+":
            self.pos += 1
        return self.read_int_body()

    def parse(self, validate, on_clause):
        state = None
        while True:
            self.skip_whitespace()
            c = self.cur()
            if state is None:
                if c == "c":
                    self.skip_line()
                else:
                    self.consume("p cnf")
                    nv = self.next_uint()
                    nc = self.next_uint()
                    state = (nv, nc)
            else:
                if c == "c":
                    self.skip_line()
                elif c is None:
                    if validate:
                        if state[1] != self.clauses:
                            raise ParseError("PARSE ERROR! DIMACS header mismatch: %d clauses declared, %d found"
                                             % (state[1], self.clauses))
                        if state[0] < len(self.vars):
                            raise ParseError("PARSE ERROR! DIMACS header mismatch: %d vars declared, %d discovered"
                                             % (state[0], len(self.vars)))
                    return state
                else:
                    lits = []
                    while True:
                        lit = self.next_int()
                        if lit == 0:
                            self.clauses += 1
                            break
                        self.vars.add(abs(lit))
                        lits.append(lit)
                    on_clause(lits)


def parse_text(text, solver, validate=False):
    """dimacs.rs:35-43 with Subst (dimacs.rs:131-168).  Returns backward_subst: var -> DIMACS id."""
    backward = {}

    def on_clause(raw):
        lits = []
        for lit_id in raw:
            a = abs(lit_id)
            while a > solver.n_vars():
                idx = solver.n_vars() + 1
                v = solver.new_var(None, True)
                backward[v] = idx
            lits.append(mk_lit(a - 1, lit_id < 0))
        solver.add_clause(lits)

    _Parser(text).parse(validate, on_clause)
    return backward


def parse_file(path, solver, validate=False):
    """dimacs.rs:16-32"""
    return parse_text(_read_text(path), solver, validate)


def read_csr_text(text, validate=False):
    """Parse into CSR (clause_off uint32[m+1], lits uint32[L]) + (header vars, header clauses)."""
    off = [0]
    lits = []

    def on_clause(raw):
        for lit_id in raw:
            lits.append(mk_lit(abs(lit_id) - 1, lit_id < 0))
        off.append(len(lits))

    hdr = _Parser(text).parse(validate, on_clause)
    return np.asarray(off, dtype=np.uint32), np.asarray(lits, dtype=np.uint32), hdr[0], hdr[1]


def read_csr(path, validate=False):
    """File -> CSR through the library's C++ reader (b200sat_read_dimacs); same quirks as `read_csr_text`."""
    import ctypes as C
    from . import lib
    u32p = C.POINTER(C.c_uint32)
    offp, litp = u32p(), u32p()
    n, hv, hc = C.c_uint32(), C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(512)
    rc = lib().b200sat_read_dimacs(os.fsencode(path), int(validate), C.byref(offp), C.byref(litp), C.byref(n),
                                   C.byref(hv), C.byref(hc), err, 512)
    if rc != 0:
        raise ParseError(err.value.decode())
    off = np.ctypeslib.as_array(offp, shape=(n.value + 1,)).copy()
    nl = int(off[-1])
    lits = np.ctypeslib.as_array(litp, shape=(max(nl, 1),))[:nl].copy()
    lib().b200sat_free_buffer(offp)
    lib().b200sat_free_buffer(litp)
    return off, lits, hv.value, hc.value


def concat_csr(instances):
    """[(clause_off, lits), ...] -> (inst_clause_off, clause_off, lits) of one batch."""
    ioff = [0]
    offs = [np.zeros(1, np.uint32)]
    alll = []
    base = 0
    nc = 0
    for off, lits in instances:
        off = np.asarray(off, dtype=np.uint64)
        offs.append((off[1:] + base).astype(np.uint32))
        alll.append(np.asarray(lits, dtype=np.uint32))
        base += int(off[-1])
        nc += len(off) - 1
        ioff.append(nc)
    return (np.asarray(ioff, dtype=np.uint32), np.concatenate(offs),
            np.concatenate(alll) if alll else np.zeros(0, np.uint32))


def write_result(writer, res, backward_subst=None):
    """dimacs.rs:46-70.  `res` is a SolveRes; backward_subst maps var -> DIMACS id (identity+1 if None)."""
    if res.kind == UNSAT:
        writer.write("UNSAT\n")
    elif res.kind == SAT:
        writer.write("SAT\n")
        for lit in res.model:
            var_id = backward_subst[lit_var(lit)] if backward_subst is not None else lit_var(lit) + 1
            writer.write("%d " % (-var_id if lit_sign(lit) else var_id))
        writer.write("0\n")
    else:
        writer.write("INDET\n")


def validate_model_text(text, model, backward_subst=None):
    """dimacs.rs:90-128: no complementary pair in the model, every clause holds a model literal."""
    lits = set()
    for lit in model:
        var_id = backward_subst[lit_var(lit)] if backward_subst is not None else lit_var(lit) + 1
        lit_id = -var_id if lit_sign(lit) else var_id
        lits.add(lit_id)
        if -lit_id in lits:
            return False
    ok = [True]

    def on_clause(cl):
        if not any(l in lits for l in cl):
            ok[0] = False

    _Parser(text).parse(False, on_clause)
    return ok[0]


def validate_model_file(path, model, backward_subst=None):
    return validate_model_text(_read_text(path), model, backward_subst)


def result_string(res):
    buf = io.StringIO()
    write_result(buf, res)
    return buf.getvalue()
