"""Top SASS instructions of an ncu source page by stall samples: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[1]
ci, cs, ce = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
data = []
for r in rows[2:]:
    try:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i] not in ("", "0")), reverse=True)[:2]
        data.append((int(r[cs]), int(r[ce]), r[ci].strip(), st, len(data)))
    except (ValueError, IndexError):
        pass
tot, tote = sum(d[0] for d in data), sum(d[1] for d in data)
print("total samples %d, warp instructions executed %d" % (tot, tote))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for d in sorted(data, reverse=True)[:n]:
    print("%6d %5.1f%% exec %9d  #%-5d %-70s %s" % (d[0], 100.0 * d[0] / max(tot, 1), d[1], d[4], d[2][:70], d[3]))
