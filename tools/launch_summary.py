"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).
    python tools/launch_summary.py profiles/r02_launches_bench_c3.csv "<command line that was profiled>" > ..._summary.txt"""
import collections
import csv
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [ln for ln in open(path) if ln.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ci, cn, cu, cv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    tot[r[cn]] += float(r[cv].replace(",", "")) * scale[r[cu]]
    cnt[r[cn]] += 1
total = sum(tot.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}")
print("(launch list of the default bench command, one GPU; per-launch times are cold-cache and serialised: compare shares)")
print(f"{'launches':>8s}  {'total ms':>10s}  {'share':>5s}  kernel")
for name, ms in tot.most_common():
    print(f"{cnt[name]:8d}  {ms:10.3f}  {100 * ms / total:4.1f}%  {name[:100]}")
