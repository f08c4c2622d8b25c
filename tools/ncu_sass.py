"""Opcode / stall histogram of one kernel from an .ncu-rep (SASS view of `ncu --page source --csv`).
    python tools/ncu_sass.py <rep> [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
per, inst, stalls, wf = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        s = int(r[ci["# Samples"]] or 0); e = int(r[ci["Instructions Executed"]] or 0)
    except ValueError:
        continue
    toks = r[ci["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    opb = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "UTC", "LDS", "STS")) else op.split(".")[0]
    per[opb] += s; inst[opb] += e
    for h in stallcols:
        stalls[h] += int(r[ci[h]] or 0)
    for h in ("L1 Wavefronts Shared", "L1 Tag Requests Global"):
        if h in ci:
            try:
                wf[(h, opb)] += int(r[ci[h]] or 0)
            except ValueError:
                pass
    lines.append((s, e, r[ci["Source"]].strip()))
tot, ti = sum(per.values()), sum(inst.values())
print("samples %d, warp instructions %d" % (tot, ti))
for op, s in per.most_common(top):
    print("%-12s samples %6d %5.1f%%   inst %11d %5.1f%%" % (op, s, 100.0 * s / max(tot, 1), inst[op], 100.0 * inst[op] / max(ti, 1)))
print("stalls:", ", ".join("%s=%d" % (k[6:], v) for k, v in stalls.most_common(8)))
print("L1:", ", ".join("%s/%s=%d" % (k[0].split()[-1], k[1], v) for k, v in wf.most_common(8)))
print("hottest instructions:")
for s, e, src in sorted(lines, reverse=True)[:top]:
    print("  %5d  x%-9d %s" % (s, e, src[:90]))
