"""Collapse the output of tools/prof_timeline.py into runs of torch glue between the libia_b200 kernels."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().split("\n\n== aten")[0].splitlines()
GLUE = ("elementwise", "Memcpy", "Memset", "reduce_kernel", "cub::", "CatArray", "index", "at_cuda_detail", "indexFunc", "tensor_kernel_scan", "distribution")
run = []
def flush():
    global run
    if run:
        c = Counter(r[2] for r in run)
        tot = sum(r[0] + r[1] for r in run)
        print(f"   [glue x{len(run)} {tot:7.1f} us incl gaps] " + ", ".join(f"{k[:60]} x{v}" for k, v in c.most_common(8)))
        run = []
for l in lines:
    m = re.match(r"\s*([\d.]+) us  gap\s+([\d.]+)  (.*)", l)
    if not m:
        continue
    d, g, n = float(m.group(1)), float(m.group(2)), m.group(3).strip()
    if any(k in n for k in GLUE):
        run.append((d, g, n))
    else:
        flush()
        print(f"{d:9.1f} us  {n[:60]} (gap {g:.1f})")
flush()
