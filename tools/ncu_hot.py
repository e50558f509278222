"""Top stall sites of one kernel from an ncu report's source page:  ncu -i X.ncu-rep --page source --csv [--kernel-id ...] > f.csv;  python tools/ncu_hot.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {c: i for i, c in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
num = lambda r, c: int(r[ix[c]] or 0)
tot = sum(num(r, "# Samples") for r in data)
print("kernel:", rows[h - 1][1] if h else "?", " instructions:", len(data), " samples:", tot,
      " warp-instructions executed:", sum(num(r, "Instructions Executed") for r in data))
agg = {c: sum(num(r, c) for r in data) for c in stall_cols}
print("stall totals:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / max(1, tot)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n_top]:
    st = sorted(((c[6:], num(r, c)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%%  exec %8d  %-72s %s" % (num(r, "# Samples"), 100.0 * num(r, "# Samples") / max(1, tot), num(r, "Instructions Executed"),
                                            r[ix["Source"]].strip()[:72], st))
