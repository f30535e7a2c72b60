"""Instruction mix + stall share of one kernel from an .ncu-rep source page: python tools/ncu_mix.py rep regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
ie, src, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot, agg, stall, lines = 0, {}, {}, []
for r in rows[start:]:
    try:
        n, s = int(r[ie]), int(r[ss])
    except (ValueError, IndexError):
        continue
    tot += n
    t = r[src].split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
    agg[op] = agg.get(op, 0) + n
    stall[op] = stall.get(op, 0) + s
    lines.append((s, n, r[src]))
st = max(sum(stall.values()), 1)
print("total warp instructions", tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
    print("%-12s %12d %5.1f%%  stall %5.1f%%" % (k, v, 100 * v / tot, 100 * stall[k] / st))
print("--- top stalled instructions")
for s, n, text in sorted(lines, reverse=True)[:12]:
    print("%5.1f%%  n=%9d  %s" % (100 * s / st, n, text[:110]))
