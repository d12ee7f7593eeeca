"""Top-N SASS instructions by stall samples from an ncu source-page CSV (ncu -i rep --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += s
    data.append((s, r))
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:n]
for i in sorted(order):
    s, r = data[i]
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:3]
    print("%5d %5.1f%% %-58s %s" % (s, 100.0 * s / tot, r[ix["Source"]][:58], " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
