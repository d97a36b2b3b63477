"""Per-kernel SASS opcode histogram of the built library: the Blackwell-native instructions (UTCHMMA = tcgen05.mma,
UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, REDG.*x4 = vector
reductions) per kernel.  python tools/sass_hist.py [path/to/libmonosdf_b200.so] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "monosdf_b200", "libmonosdf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "USETMAXREG", "MUFU", "REDG", "HMMA"]
kern, hist, total, arch = None, {}, {}, None
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEY:
            if op.startswith(k):
                hist[kern][k + ("x4" if "x4" in op else "")] += 1
print("# %s (%s): kernels with tensor-core / TMA / TMEM instructions" % (os.path.basename(lib), arch))
dem = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
for name, k in zip(dem, hist):
    h = hist[k]
    if not (h["UTCHMMA"] or h["UTMALDG"] or h["UTMASTG"] or h["LDTM"]):
        continue
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    short = re.sub(r"\(.*", "", short)[:110]
    print("%-112s %6d instr  %s" % (short, total[k], " ".join("%s=%d" % kv for kv in sorted(h.items()))))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("# whole library: " + " ".join("%s=%d" % kv for kv in sorted(tot.items())))
