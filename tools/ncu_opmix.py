"""Opcode mix of one kernel from an .ncu-rep source page: python tools/ncu_opmix.py rep kernel_regex [voxels] [--lines N]"""
import collections, csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
vox = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 0
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
tot, n, lines, ix, kernels = collections.Counter(), 0, [], None, 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        continue
    if r and r[0] == "Address":
        ix = {h: i for i, h in enumerate(r)}
        continue
    if ix is None or len(r) < len(ix):
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    e = int(r[ix["Instructions Executed"]])
    tot[op] += e
    n += e
    lines.append((e, int(r[ix["# Samples"]]), src))
print("total warp instructions", n, ("per voxel %.2f" % (n * 32 / vox)) if vox else "")
for op, c in tot.most_common(24):
    print("%-28s %12d %5.1f%% %s" % (op, c, 100.0 * c / n, ("%.2f/voxel" % (c * 32 / vox)) if vox else ""))
if nlines:
    for i, (e, smp, src) in enumerate(lines):
        if e >= sorted([l[0] for l in lines])[-nlines]:
            print(i, e, smp, src[:110])
