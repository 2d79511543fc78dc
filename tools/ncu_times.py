"""ncu_times.py log.csv -- per-kernel launch counts and total / mean gpu__time_duration from an `ncu --csv --metrics gpu__time_duration.sum` log"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns": v /= 1000.0
    elif r[ui] == "ms": v *= 1000.0
    name = re.sub(r"^void ovd_kernel<(.*)>\(.*$", r"\1", r[ki])
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in acc.values())
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-34s launches %3d  total %9.1f us  mean %8.1f us  share %5.1f %%" % (k[:34], n, t, t / n, 100 * t / tot))
print("total %.1f us" % tot)
