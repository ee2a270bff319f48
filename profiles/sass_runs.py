import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix=hdr.index('Instructions Executed'); src=hdr.index('Source'); sm=hdr.index('# Samples')
tot=sum(int(r[ix]) for r in data if r[ix].isdigit()); tots=sum(int(r[sm]) for r in data if r[sm].isdigit())
print('total warp inst', tot, 'sass lines', len(data), 'samples', tots)
# contiguous runs with the same execution count
runs=[]
for i,r in enumerate(data):
    if not r[ix].isdigit(): continue
    c=int(r[ix]); s=int(r[sm]) if r[sm].isdigit() else 0
    if runs and runs[-1][0]==c: runs[-1][1]+=1; runs[-1][2]+=s; runs[-1][4]=i
    else: runs.append([c,1,s,i,i])
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
for c,n,s,i0,i1 in runs:
    if c*n/tot>thr or s/tots>thr:
        print(f"lines {i0:5d}-{i1:5d} count {c:10d} n {n:4d} inst share {c*n/tot:.3f} samples {s/tots:.3f}   {data[i0][src].strip()[:50]}")
op=collections.Counter()
for r in data:
    if not r[ix].isdigit(): continue
    t=r[src].split()
    o=t[1] if t[0].startswith('@') else t[0]
    op[o.split('.')[0]]+=int(r[ix])
print(' '.join(f"{o}:{c/tot:.3f}" for o,c in op.most_common(22)))
