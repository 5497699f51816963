/*
 * dimacs_io.cpp -- DIMACS(.gz) -> CSR reader and result writer of the C ABI (host side).
 *
 * Mirrors the reference's character-level parser src/sat/dimacs.rs:171-365 and its quirks:
 * comment lines are recognised only at token start, the "p cnf" header is mandatory, header counts
 * are ignored unless `strict` (dimacs.rs:230-240), '+' signs are accepted, anything else that is not
 * an integer is "int expected" (so the SATLIB '%' trailer is an error, as in the reference), gzip is
 * sniffed by its magic bytes (dimacs.rs:16-32).  Output is the CSR the batch engine ingests; the
 * device then replays Subst::add_clause order (dimacs.rs:146-167).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <unordered_set>
#include <vector>
#include <zlib.h>

#include "../../include/b200sat.h"

namespace {
struct Parser {
    const std::string &buf;
    size_t pos = 0;
    std::string err;
    explicit Parser(const std::string &b) : buf(b) {}
    int cur() const { return pos < buf.size() ? (unsigned char)buf[pos] : -1; }
    static bool ws(int c) { return c == ' ' || (c >= 9 && c <= 13); }
    void skip_ws() { while (pos < buf.size() && ws((unsigned char)buf[pos])) pos++; }
    void skip_line() { while (pos < buf.size()) { if (buf[pos++] == '\n') break; } }
    bool consume(const char *t) {
        for (const char *p = t; *p; p++) {
            if (cur() == (unsigned char)*p) pos++;
            else { err = std::string("failed to consume; expected '") + t + "'"; return false; }
        }
        return true;
    }
    bool int_body(unsigned long long &v) {
        size_t len = 0;
        v = 0;
        while (cur() >= '0' && cur() <= '9') { v = v * 10 + (unsigned)(cur() - '0'); len++; pos++; }
        if (!len) { err = "int expected"; return false; }
        return true;
    }
    bool next_int(long long &out) {
        skip_ws();
        int sign = 1;
        if (cur() == '+') pos++;
        else if (cur() == '-') { pos++; sign = -1; }
        unsigned long long v;
        if (!int_body(v)) return false;
        out = sign * (long long)v;
        return true;
    }
    bool next_uint(unsigned long long &out) {
        skip_ws();
        if (cur() == '+') pos++;
        return int_body(out);
    }
};

bool slurp(const char *path, std::string &out, std::string &err) {
    gzFile f = gzopen(path, "rb"); /* transparently reads plain files too */
    if (!f) { err = std::string("cannot open ") + path; return false; }
    char tmp[1 << 16];
    int n;
    while ((n = gzread(f, tmp, sizeof tmp)) > 0) out.append(tmp, (size_t)n);
    gzclose(f);
    if (n < 0) { err = "read error"; return false; }
    return true;
}
int fail(char *err, size_t cap, const std::string &m) {
    if (err && cap) snprintf(err, cap, "%s", m.c_str());
    return B200SAT_E_ARG;
}
} // namespace

extern "C" int b200sat_read_dimacs(const char *path, int strict, uint32_t **clause_off, uint32_t **lits, uint32_t *n_clauses,
                                   uint32_t *hdr_vars, uint32_t *hdr_clauses, char *err, size_t errcap) {
    std::string buf, e;
    if (!slurp(path, buf, e)) return fail(err, errcap, e);
    Parser p(buf);
    std::vector<uint32_t> off{0}, ls;
    std::unordered_set<uint32_t> seen;
    bool parsing = false;
    unsigned long long vars = 0, clauses = 0, found = 0, distinct = 0;
    for (;;) {
        p.skip_ws();
        int c = p.cur();
        if (!parsing) {
            if (c == 'c') p.skip_line();
            else {
                if (!p.consume("p cnf") || !p.next_uint(vars) || !p.next_uint(clauses)) return fail(err, errcap, p.err);
                parsing = true;
            }
        } else if (c == 'c') p.skip_line();
        else if (c < 0) {
            if (strict) {
                if (clauses != found)
                    return fail(err, errcap, "PARSE ERROR! DIMACS header mismatch: " + std::to_string(clauses) + " clauses declared, " +
                                                 std::to_string(found) + " found");
                if (vars < distinct)
                    return fail(err, errcap, "PARSE ERROR! DIMACS header mismatch: " + std::to_string(vars) + " vars declared, " +
                                                 std::to_string(distinct) + " discovered");
            }
            break;
        } else {
            for (;;) {
                long long lit;
                if (!p.next_int(lit)) return fail(err, errcap, p.err);
                if (lit == 0) { found++; off.push_back((uint32_t)ls.size()); break; }
                unsigned long long v = (unsigned long long)(lit < 0 ? -lit : lit);
                if (v > 0x7FFFFFF0ull) return fail(err, errcap, "variable id too large");
                if (seen.insert((uint32_t)v).second) distinct++; /* HashSet in the reference (dimacs.rs): size follows the distinct ids */
                ls.push_back((uint32_t)(((v - 1) << 1) | (lit < 0 ? 1u : 0u)));
            }
        }
    }
    *n_clauses = (uint32_t)found;
    if (hdr_vars) *hdr_vars = (uint32_t)vars;
    if (hdr_clauses) *hdr_clauses = (uint32_t)clauses;
    *clause_off = (uint32_t *)malloc(off.size() * 4);
    memcpy(*clause_off, off.data(), off.size() * 4);
    *lits = (uint32_t *)malloc((ls.size() ? ls.size() : 1) * 4);
    if (!ls.empty()) memcpy(*lits, ls.data(), ls.size() * 4);
    return 0;
}

extern "C" void b200sat_free_buffer(void *p) { free(p); }

/* dimacs.rs:46-70: "UNSAT" | "INDET" | "SAT\n<signed ids in var order> 0".  model bytes: 1 true, 0 false, 2 absent. */
extern "C" int b200sat_write_result(const char *path, int verdict, const uint8_t *model, uint32_t n_vars) {
    FILE *f = fopen(path, "w");
    if (!f) return B200SAT_E_ARG;
    if (verdict == B200SAT_UNSAT) fprintf(f, "UNSAT\n");
    else if (verdict == B200SAT_SAT) {
        fprintf(f, "SAT\n");
        for (uint32_t v = 0; v < n_vars; v++)
            if (model[v] != 2) fprintf(f, "%s%u ", model[v] == 0 ? "-" : "", v + 1);
        fprintf(f, "0\n");
    } else fprintf(f, "INDET\n");
    fclose(f);
    return 0;
}
