"""Summarise the ncu launch list of the bench command (gpu__time_duration.sum per launch) into a markdown table:
per kernel the launches and time inside the LAST step of the run and its share of that step.
usage: python scripts/launch_list_summary.py gpurun_out/launches.csv 101 > profiles/r01_launches_rb3d256_final.md"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; rows = rows[1:]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 101
name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
import re
def short(n):
    n = re.sub(r"\(.*$", "", n).replace("void ", "").replace("<unnamed>::", "")
    return n[:70]
launches = [(short(r[name_i]), float(r[val_i]) * 1e-6) for r in rows]          # ms
print(f"# ncu launch list of `bench.py --gpus 1 --steps 1 --warmup 3` (rb3d 256^3, 1 GPU): {len(launches)} launches in total")
print(f"\n`ncu --metrics gpu__time_duration.sum --clock-control none`; times are cold-cache and serialised: only the SHARE of the step is comparable with bench.py's per-kernel accounting.  Last step = last {per_step} launches.\n")
last = launches[-per_step:]
agg = collections.OrderedDict()
for n, ms in last:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("| kernel | launches/step | ms/step (ncu) | share |\n|---|---|---|---|")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {ms:.3f} | {ms / tot:.3f} |")
print(f"| total | {per_step} | {tot:.3f} | 1.000 |")
print("\nSet-up launches before the first step (assembly, factorisation, fills):\n")
agg2 = collections.OrderedDict()
for n, ms in launches[:len(launches) - 4 * per_step]:
    a = agg2.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
for n, (c, ms) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"* `{n}`: {c} launches, {ms:.3f} ms")
