#!/usr/bin/env python3
"""Segment the SASS of one kernel (ncu source page exported as CSV) into runs of equal execution count:
instructions per warp, share of executed instructions / of stall samples, active lanes, opcode mix.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_source_segments.py src.csv seg 0.3 <number of warps launched>
"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); ismp=hdr.index("# Samples"); ith=hdr.index("Thread Instructions Executed")
data=[(r[isrc].strip(), int(r[ia]), int(r[ismp]), int(r[ith])) for r in rows[2:] if len(r)>ia]
tot=sum(d[1] for d in data); tots=sum(d[2] for d in data)
print("total inst",tot,"samples",tots, "ninstr", len(data))
# segment by exec count changes >3%
seg=[]; start=0
W=int(sys.argv[4]) if len(sys.argv)>4 else 1911612
def flush(a,b):
    ins=sum(d[1] for d in data[a:b]); sm=sum(d[2] for d in data[a:b]); th=sum(d[3] for d in data[a:b])
    ops={}
    for d in data[a:b]:
        op=d[0].split()[0] if not d[0].startswith('@') else d[0].split()[1]
        op=op.split('.')[0]
        ops[op]=ops.get(op,0)+1
    top=sorted(ops.items(), key=lambda x:-x[1])[:6]
    print(f"{a:5d}-{b:5d} n={b-a:4d} exec/warp={data[a][1]/W:6.3f} inst%={100*ins/tot:5.1f} smp%={100*sm/tots:5.1f} lanes={th/max(ins,1):5.1f} {top}")
mode=sys.argv[2] if len(sys.argv)>2 else "seg"
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.15
for i in range(1,len(data)+1):
    if i==len(data) or abs(data[i][1]-data[start][1])>thr*max(data[start][1],W*0.01):
        flush(start,i); start=i
