"""Which Blackwell-specific instructions each kernel of libtsp_b200.so contains (cuobjdump -sass):
UTMALDG = TMA tensor loads, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32, VIMNMX = packed integer min/max,
ACQBULK / PREEXIT = programmatic dependent launch, REDG = no-return global atomics.
    python tools/sass_proof.py > profiles/r2_sass_tma.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tissue_image_processing_b200", "libtsp_b200.so")
WATCH = ("UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "VIMNMX", "ACQBULK", "PREEXIT", "REDG", "IDP")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        total[fn] += 1
        if m.group(1) in WATCH:
            counts[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
print("# %s: sm_100a SASS, instruction counts per kernel (only kernels with at least one of %s)" %
      (os.path.basename(lib), ", ".join(WATCH)))
grand = collections.Counter()
for fn, name in sorted(zip(total, names), key=lambda p: p[1]):
    if counts[fn]:
        short = re.sub(r"\(.*", "", name)
        print("%-58s total=%-6d %s" % (short[:58], total[fn], "  ".join("%s=%d" % kv for kv in sorted(counts[fn].items()))))
        grand.update(counts[fn])
print("# library totals: " + "  ".join("%s=%d" % kv for kv in sorted(grand.items())))
