import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
which=int(sys.argv[2]) if len(sys.argv)>2 else -1
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
st=starts[which]; en=starts[starts.index(st)+1] if starts.index(st)+1<len(starts) else len(rows)
print(rows[st][1], "section", which, "of", len(starts))
H=rows[st+1]; data=[r for r in rows[st+2:en] if len(r)==len(H)]
ia=H.index('Source'); ie=H.index('Instructions Executed'); istall=H.index('Warp Stall Sampling (All Samples)')
tot=sum(int(r[ie]) for r in data)
print("total warp instr", tot, "static instrs", len(data))
ops=collections.Counter(); stalls=collections.Counter()
for r in data:
    toks=r[ia].split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    op='.'.join(op.split('.')[:2])
    ops[op]+=int(r[ie]); stalls[op]+=int(r[istall])
for op,c in ops.most_common(30): print(f"{op:16s} {c:12d} {100*c/tot:5.1f}%  stalls {stalls[op]}")
print("---- execution count by address region")
cnts=[int(r[ie]) for r in data]
i=0
while i<len(data):
    j=i
    while j<len(data) and abs(cnts[j]-cnts[i])<=0.15*max(cnts[i],1): j+=1
    print(f"[{i:4d},{j:4d}) n={j-i:4d} exec~{cnts[i]:10d} total={sum(cnts[i:j]):12d} stalls={sum(int(r[istall]) for r in data[i:j])}")
    i=j
