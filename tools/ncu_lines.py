"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
per = collections.OrderedDict()
cur = None
seen = set()
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        ix = hdr.index("Instructions Executed")
        sx = hdr.index("# Samples")
        ax = 2
        continue
    if hdr is None or len(r) <= ix:
        continue
    if r[0].strip():
        cur = (r[0], r[1])
        per.setdefault(cur, [0, 0, 0])
    try:
        n, s = int(r[ix]), int(r[sx])
    except ValueError:
        continue
    if cur is not None and r[ax].strip() and r[ax] not in seen:     # every SASS address once
        seen.add(r[ax])
        per[cur][0] += n
        per[cur][1] += s
        per[cur][2] += 1
tot = sum(v[0] for v in per.values())
tots = sum(v[1] for v in per.values())
print("total warp instructions %d, samples %d" % (tot, tots))
for (ln, src), (n, s, k) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%9d %5.1f%%  smp %5.1f%%  sass %3d  L%-4s %s" % (n, 100.0 * n / max(tot, 1), 100.0 * s / max(tots, 1), k, ln, src.strip()[:100]))
