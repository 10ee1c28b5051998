"""Summarise an ncu source-page CSV: opcode mix, stall reasons, hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python tools/ncu_hot.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].replace(".", "").isdigit()]
ci = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(float(r[ci[k]] or 0))
tot_s = sum(I(r, "# Samples") for r in data) or 1
tot_i = sum(I(r, "Instructions Executed") for r in data) or 1
print("kernel:", rows[0][1][:100] if rows and len(rows[0]) > 1 else "?")
print("total samples", tot_s, "warp-instructions", tot_i)
agg, aggs = collections.Counter(), collections.Counter()
for r in data:
    src = r[ci["Source"]].split()
    op = src[0] if src and not src[0].startswith("@") else (src[1] if len(src) > 1 else "?")
    op = op.split(".")[0]
    agg[op] += I(r, "Instructions Executed")
    aggs[op] += I(r, "# Samples")
print("opcode        instr%  samples%")
for op, c in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{op:12s} {100 * c / tot_i:6.2f}% {100 * aggs[op] / tot_s:6.2f}%")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tots = {h: sum(I(r, h) for r in data) for h in st}
print("stalls:", [(k, v) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]])
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:16]:
    print(I(r, "# Samples"), I(r, "Instructions Executed"), r[ci["Source"]][:100])
