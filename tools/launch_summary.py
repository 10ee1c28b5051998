"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 1 --warmup 3 --no-cpu-baseline`:
shares over the whole process and over ONE training step (the launches between two consecutive pairs of adamw_kernel
launches; the step profiled is the timed one).  usage: python tools/launch_summary.py launches.csv > profiles/rNN_launches_summary.md"""
import collections, csv, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v, u = float(r[vi].replace(",", "")), r[ui]
    launches.append((r[ki], v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v))
mine = lambda k: not any(s in k for s in ("at::", "at_cuda_detail", "cub::", "nccl", "Memcpy", "Memset", "cutlass", "cublas", "gemm"))
short = lambda k: re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", k).split("(")[0][:72]


def table(ls, top):
    agg = collections.OrderedDict()
    for k, ms in ls:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(ms for _, ms in ls)
    out = [f"Total kernel time {tot:.2f} ms over {len(ls)} launches; kernels of libia_b200.so: "
           f"{100 * sum(ms for k, ms in ls if mine(k)) / tot:.1f} % of it ({sum(1 for k, _ in ls if mine(k))} launches).", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| `{short(k)}`{' (libia_b200)' if mine(k) else ''} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")
    return out


adam = [i for i, (k, _) in enumerate(launches) if "adamw_kernel" in k]
# two adamw launches per step (main arena, variance arena); steps: 3 warm-up, 1 timed, 1 e2e
step = launches[adam[-5] + 1: adam[-3] + 1] if len(adam) >= 6 else launches
print("# ncu launch list summary\n")
print("Command (on the B200 box): `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "
      "gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline`\n")
print("Per-launch times under ncu are cold-cache and serialised: only SHARES are meaningful.\n")
print("## One training step (the timed step: launches between two consecutive optimizer launches)\n")
print("\n".join(table(step, 30)))
print("\n## Whole process (16 occupancy warm-up refreshes, 5 steps, hash-grid micro-benchmark with its sector-gather peaks)\n")
print("\n".join(table(launches, 25)))
