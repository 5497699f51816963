"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line and per function of
cdcl_warp.cuh (profiling aid, CPU only).  usage: ncu_lines2.py <export.csv> [lo-hi line range to list]"""
import csv, re, sys, collections, bisect
rows = list(csv.reader(open(sys.argv[1])))
cur_file = None; hdr = None
per_line = collections.Counter(); samp = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = {}; [ci.setdefault(n, i) for i, n in enumerate(hdr)]; ci_all = [(i, n) for i, n in enumerate(hdr)]; continue
    if hdr is None or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    def val(name):
        try: return int(r[ci[name]])
        except Exception: return 0
    k = (cur_file, ln)
    per_line[k] += val("Instructions Executed"); samp[k] += val("# Samples")
    for i, n in ci_all:
        if n.startswith("stall_") and "Not Issued" not in n:
            try: v = int(r[i])
            except Exception: v = 0
            if v: stalls[k][n] += v
tot = sum(per_line.values()); tots = sum(samp.values())
funcs = []
for i, l in enumerate(open("/root/repo/minisat-rust_b200/csrc/cdcl_warp.cuh"), 1):
    m = re.match(r'\s+DEV(?:NI)?\s+(?:static\s+)?[\w:<>\*&\s]+?\s+\**(\w+)\(', l)
    if m: funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
def fn_of(k):
    f, l = k
    if f != "cdcl_warp.cuh": return f
    i = bisect.bisect_right(starts, l) - 1
    return funcs[i][1] if i >= 0 else f
pf = collections.Counter(); sf = collections.Counter(); stf = collections.defaultdict(collections.Counter)
for k, v in per_line.items(): pf[fn_of(k)] += v
for k, v in samp.items(): sf[fn_of(k)] += v
for k, d in stalls.items():
    for c, v in d.items(): stf[fn_of(k)][c] += v
print("total warp instructions %d, samples %d" % (tot, tots))
for f, v in pf.most_common(28):
    st = ", ".join("%s %.0f%%" % (c.replace("stall_", ""), 100.0 * n / max(1, sf[f])) for c, n in stf[f].most_common(4))
    print("%-26s %7.2f%% %7.2f%%  %s" % (f, 100.0 * v / tot, 100.0 * sf[f] / max(1, tots), st))
if len(sys.argv) > 2:
    lo, hi = [int(x) for x in sys.argv[2].split("-")]
    srcl = open("/root/repo/minisat-rust_b200/csrc/cdcl_warp.cuh").read().split("\n")
    for l in range(lo, hi + 1):
        k = ("cdcl_warp.cuh", l)
        if per_line.get(k, 0) == 0 and samp.get(k, 0) == 0: continue
        st = ", ".join("%s %d" % (c.replace("stall_", ""), n) for c, n in stalls[k].most_common(2))
        print("%5d %6.2f%% %6.2f%%  %-36s | %s" % (l, 100.0 * per_line[k] / tot, 100.0 * samp[k] / tots, st, srcl[l - 1].strip()[:90]))
