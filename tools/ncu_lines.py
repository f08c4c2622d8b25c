"""Rank source lines of one kernel in an .ncu-rep by warp-stall samples (uses `ncu --page source --csv`)."""
import csv
import subprocess
import sys

rep, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pattern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hi = his[0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
end = his[1] if len(his) > 1 else len(rows)
tot, per = 0, []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:end]:
    if not r or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    num = lambda v: int(v) if v not in ("", "-") else 0  # noqa: E731
    s = num(r[ci["# Samples"]])
    ie = num(r[ci["Instructions Executed"]])
    st = sorted(((num(r[ci[c]]), c) for c in stall_cols), reverse=True)[:2]
    tot += s
    per.append((s, ie, ln, r[1][:100], st))
per.sort(reverse=True)
print("total samples", tot)
for s, ie, ln, src, st in per[:top]:
    print("%6d %5.1f%% inst=%10d L%-4d %-100s %s" % (s, 100 * s / max(tot, 1), ie, ln, src, " ".join("%s=%d" % (c[6:], v) for v, c in st)))
