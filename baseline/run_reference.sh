#!/bin/bash
# Cross-check hook: when a Rust toolchain (or a `minisat` binary) IS available, build the
# unmodified reference and dump per-instance {verdict, restarts, conflicts, decisions, propagations, conflict literals}
# so that the oracle's trace counters can be pinned.  Not runnable in the build image (no cargo / rustc / minisat).
#   usage: baseline/run_reference.sh /path/to/minisat-rust  file1.cnf.gz [file2.cnf.gz ...]
set -e
REF="$1"; shift
command -v cargo >/dev/null || { echo "cargo not found: the reference cannot be built here" >&2; exit 2; }
(cd "$REF" && cargo build --release)
for f in "$@"; do
  out=$("$REF/target/release/minisat-rust" --verb=1 "$f" 2>&1 || true)
  python3 - "$f" <<PY
import re, sys, json
text = """$out"""
def val(k):
    m = re.search(r"^%s\s*:\s*(\d+)" % re.escape(k), text, re.M)
    return int(m.group(1)) if m else None
verdict = "SAT" if "SATISFIABLE" in text and "UNSATISFIABLE" not in text else "UNSAT" if "UNSATISFIABLE" in text else "INDET"
print(json.dumps({"file": sys.argv[1], "verdict": verdict, "restarts": val("restarts"), "conflicts": val("conflicts"),
                  "decisions": val("decisions"), "propagations": val("propagations"), "conflict_literals": val("conflict literals")}))
PY
done
