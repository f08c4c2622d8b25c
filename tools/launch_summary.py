"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total time, share of the step).
    python tools/launch_summary.py gpurun_out/<tag>/launches.csv "<header comment>" > profiles/<name>.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ci["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").strip()
    tot[name] += float(r[ci["Metric Value"]].replace(",", ""))
    cnt[name] += 1
total = sum(tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print("# per-launch times are cold-cache and serialised: compare SHARES")
print("kernel,launches,total_ns,share")
for k, v in tot.most_common():
    print("%s,%d,%d,%.3f" % (k.replace(",", ";"), cnt[k], v, v / total))
print("# total,%d,%d,1.000" % (sum(cnt.values()), total))
