"""SASS opcode histogram per CUDA source of libetch_b200.so (cuobjdump -sass over etch_b200/build/*.o): the mnemonics that prove
the Blackwell-native paths (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit,
UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UCGABAR = cluster barrier, REDUX = redux.sync, FFMA2 = fma.rn.f32x2.
    python tools/sass_histogram.py > profiles/<round>_sass_histogram.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR", "REDUX", "FFMA", "FFMA2", "DFMA", "MUFU", "LDS", "STS",
        "LDG", "STG", "ATOMG", "RED", "HMMA", "LDL", "STL"]
print("# SASS opcode histogram (sm_100a), one row per CUDA source\n")
print("| source | kernels | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|---|" + "---|" * len(KEYS))
for obj in sorted(glob.glob(os.path.join(ROOT, "etch_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ops = collections.Counter()
    nk = sass.count("Function :")
    n = 0
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            n += 1
            hits = [k for k in KEYS if m.group(1).startswith(k)]
            if hits:                    # longest prefix wins: FFMA2 (packed fma.rn.f32x2) is not FFMA, REDUX is not RED
                ops[max(hits, key=len)] += 1
    if nk == 0:
        continue
    print("| %s | %d | %d | " % (os.path.basename(obj)[:-2] + ".cu", nk, n) + " | ".join(str(ops.get(k, 0)) for k in KEYS) + " |")
