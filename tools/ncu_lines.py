"""Join an `ncu --page source --csv` SASS export with nvdisasm line info and aggregate executed warp
instructions / stall samples per source line and per function (profiling aid, CPU only).
usage: ncu_lines.py <src.csv> <nvdisasm -g -c output> <kernel mangled name> [top]"""
import csv, re, sys, collections, bisect
src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: offset -> (file,line)
off2line = {}
cur = None
inside = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kname
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*)', ln)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
base = int(rows[2][0], 16)
per_line = collections.Counter(); samp_line = collections.Counter(); tot = 0; tot_s = 0
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
stalls_line = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    off = int(r[0], 16) - base
    ie = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    key = off2line.get(off)
    per_line[key] += ie; samp_line[key] += s; tot += ie; tot_s += s
    for c in stall_cols:
        v = int(r[ci[c]] or 0)
        if v: stalls_line[key][c] += v
# function map for cdcl_warp.cuh by line ranges
funcs = []
try:
    path = "/root/repo/minisat-rust_b200/csrc/cdcl_warp.cuh"
    for i, l in enumerate(open(path), 1):
        m = re.match(r'\s+DEV(?:NI)?\s+(?:static\s+)?[\w:<>\*&\s]+?\s+\**(\w+)\(', l)
        if m: funcs.append((i, m.group(1)))
except Exception: pass
starts = [f[0] for f in funcs]
def fn_of(key):
    if key is None: return "?"
    f, l = key
    if f != "cdcl_warp.cuh": return f
    i = bisect.bisect_right(starts, l) - 1
    return funcs[i][1] if i >= 0 else f
per_fn = collections.Counter(); samp_fn = collections.Counter(); stalls_fn = collections.defaultdict(collections.Counter)
for k, v in per_line.items(): per_fn[fn_of(k)] += v
for k, v in samp_line.items(): samp_fn[fn_of(k)] += v
for k, d in stalls_line.items():
    for c, v in d.items(): stalls_fn[fn_of(k)][c] += v
print("total warp instructions %d, samples %d" % (tot, tot_s))
print("%-28s %8s %8s  top stalls" % ("function", "inst%", "samp%"))
for f, v in per_fn.most_common(30):
    st = ", ".join("%s %.0f%%" % (c.replace("stall_", ""), 100.0 * n / max(1, samp_fn[f])) for c, n in stalls_fn[f].most_common(4))
    print("%-28s %7.2f%% %7.2f%%  %s" % (f, 100.0 * v / tot, 100.0 * samp_fn[f] / tot_s, st))
print()
print("top lines by samples")
for k, v in samp_line.most_common(top):
    print("%-22s %6s inst %6.2f%% samp %6.2f%%" % (k[0] if k else "?", k[1] if k else "", 100.0 * per_line[k] / tot, 100.0 * v / tot_s))
if len(sys.argv) > 5:
    lo, hi = [int(x) for x in sys.argv[5].split("-")]
    print("\nlines %d-%d of cdcl_warp.cuh (inst%%, samp%%, top stall)" % (lo, hi))
    srcl = open("/root/repo/minisat-rust_b200/csrc/cdcl_warp.cuh").read().split("\n")
    for l in range(lo, hi + 1):
        k = ("cdcl_warp.cuh", l)
        if per_line.get(k, 0) == 0 and samp_line.get(k, 0) == 0: continue
        st = ", ".join("%s %d" % (c.replace("stall_", ""), n) for c, n in stalls_line[k].most_common(2))
        print("%5d %6.2f%% %6.2f%%  %-40s | %s" % (l, 100.0 * per_line[k] / tot, 100.0 * samp_line[k] / tot_s, st, srcl[l - 1].strip()[:90]))
