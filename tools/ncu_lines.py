"""Developer script: join an `ncu --page source --csv` SASS table with `nvdisasm -g -c` line info and
aggregate warp-stall samples per source line of fpm_update*.cuh (outermost inlining site).
usage: ncu_lines.py <ncu_sass.csv> <nvdisasm_listing.txt> [top]"""
import csv, re, sys, collections
src_csv, lst, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
off2line, cur = {}, None
for ln in open(lst):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        cur = chain
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m and cur is not None:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rows[2:]:
    off = int(r[ia], 16) - base
    chain, _ = off2line.get(off, ([("?", 0)], ""))
    key = next(((f.split("/")[-1], l) for f, l in reversed(chain) if "fpm_update" in f), (chain[-1][0].split("/")[-1], chain[-1][1]))
    s = int(r[isamp] or 0)
    agg[key][0] += s
    agg[key][1] += int(r[iex] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: agg[key][2][hdr[i][6:]] += v
    tot += s
print("total samples", tot)
srcs = {}
for key, (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key
    if f not in srcs:
        try: srcs[f] = open("/root/repo/fpm-opencv_b200/csrc/" + f).read().split("\n")
        except OSError: srcs[f] = []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
    print("%5.1f%% %9d inst  %s:%d  %-70s  %s" % (100.0 * s / tot, ex, f, l, text, " ".join("%s=%d" % kv for kv in st.most_common(4))))

# ---- per-stage totals: instructions of inlined intrinsics (no inlining chain in the line table) inherit the last
# fpm_update*.cuh line seen in address order
if len(sys.argv) > 4:
    bounds = [(int(a), b) for a, b in (x.split(":") for x in sys.argv[4].split(","))]   # "296:S1,359:S2,..."
    def stage(l):
        name = "pre"
        for lo, nm in bounds:
            if l >= lo: name = nm
        return name
    st_s, st_i, st_st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    last = 0
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        chain, _ = off2line.get(off, ([("?", 0)], ""))
        own = [l for f, l in chain if "fpm_update" in f]
        if own: last = own[-1]
        sg = stage(last)
        st_s[sg] += int(r[isamp] or 0); st_i[sg] += int(r[iex] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: st_st[sg][hdr[i][6:]] += v
    print("\nper stage: samples%, warp instructions, top stalls")
    for nm in ["pre"] + [b for _, b in bounds]:
        print("  %-4s %5.1f%% %11d  %s" % (nm, 100.0 * st_s[nm] / tot, st_i[nm], " ".join("%s=%.1f%%" % (k, 100.0 * v / tot) for k, v in st_st[nm].most_common(6))))
