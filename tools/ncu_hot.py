"""Hot SASS instructions of an ncu source page dump: python tools/ncu_hot.py src.csv [n]
(ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
totinst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("samples", tot, "warp-instructions", totinst)
byop = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    byop[op][0] += int(r[ix["# Samples"]] or 0)
    byop[op][1] += int(r[ix["Instructions Executed"]] or 0)
print("--- by opcode (samples%, inst%)")
for k, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:22]:
    print("%-10s %5.1f%% %5.1f%%" % (k, 100.0 * v[0] / tot, 100.0 * v[1] / totinst))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("--- top instructions")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("%5.1f%% inst %8s  %-70s %s" % (100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                          " ".join("%s=%d" % (s[6:], c) for c, s in st)))
