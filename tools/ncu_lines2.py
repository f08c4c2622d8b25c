"""Per-source-line view of one kernel launch in an .ncu-rep: shared-memory wavefronts (total / excessive = bank conflicts a layout can
fix) and warp-stall samples, over ALL source files of the kernel (tools/ncu_lines.py looks at the first file only).
    python tools/ncu_lines2.py <report.ncu-rep> <kernel regex> <launch-skip>"""
import csv, subprocess, sys
rep, pattern, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+pattern,"--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
his=[i for i,r in enumerate(rows) if r and r[0]=='Line No' and len(r)>5]
res=[]
for k,hi in enumerate(his):
    h=rows[hi]; ci={}
    for i,n in enumerate(h):
        ci.setdefault(n,i)
    fname = rows[hi-2][1] if hi>=2 else ''
    end = his[k+1] if k+1<len(his) else len(rows)
    for r in rows[hi+1:end]:
        if len(r)<len(h) or r[0]=='' : continue
        try: ln=int(r[0])
        except: continue
        def f(n):
            try: return float(r[ci[n]])
            except: return 0
        res.append((f('L1 Wavefronts Shared Excessive'), f('L1 Wavefronts Shared'), f('# Samples'), fname.split('/')[-1], ln, r[1][:100]))
print('samples',sum(x[2] for x in res),'excess',sum(x[0] for x in res),'wavefronts',sum(x[1] for x in res))
print('--- by excess')
for x in sorted(res,reverse=True)[:12]: print(x)
print('--- by samples')
for x in sorted(res,key=lambda x:-x[2])[:25]: print(x)
