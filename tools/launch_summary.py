"""per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    k = r[ki].split("(")[0]
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
print(f"{sys.argv[1]}: {sum(cnt.values())} launches, {s/1e3:.2f} ms of kernel time")
for k, v in tot.most_common():
    print(f"  {k[:70]:70s} {cnt[k]:6d} x {v/cnt[k]:9.1f} us = {v/1e3:9.2f} ms  {100*v/s:5.1f} %")
