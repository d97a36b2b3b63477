"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_summary.py file.csv [n]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.defaultdict(lambda: [0, 0.0])
for d in data:
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", d["Kernel Name"])
    name = re.sub(r"\(.*", "", name).replace("__nv_bfloat16", "bf16")[:100]
    v = float(d["Metric Value"])
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.1f us" % (len(data), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%9.1f us %5.1f%% n=%5d avg %7.1f  %s" % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k))
