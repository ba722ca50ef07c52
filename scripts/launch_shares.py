"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/launch_shares.py gpurun_out/launches.csv [skip_launches] > profiles/rN_launch_shares.csv
Only the library's kernels (namespace xm) are listed; shares are of their total."""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows:
    name = r[ki]
    if "at::" in name or "nccl" in name.lower() or "elementwise" in name or "cub::" in name:
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "").replace("xm::", "")
    t = float(r[vi].replace(",", "")) / 1e3
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_us,avg_us,share_of_our_kernels")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.1f,%.1f%%" % (k, n, t, t / n, 100 * t / tot))
