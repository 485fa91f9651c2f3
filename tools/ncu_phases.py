import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
ph=0; tot=collections.Counter(); ops=collections.defaultdict(collections.Counter); stall=collections.Counter(); samples=collections.Counter()
for r in data:
    src=r[ix['Source']].strip(); n=int(r[ix['Instructions Executed']] or 0); s=int(r[ix['# Samples']] or 0)
    t=re.sub(r'^@!?U?P\d+\s+','',src); op=re.split(r'[ .]',t)[0]
    tot[ph]+=n; ops[ph][op]+=n; samples[ph]+=s
    if src.startswith('BAR.SYNC') or ' BAR.SYNC' in src: ph+=1
W=float(sys.argv[2]) if len(sys.argv)>2 else 840960.0
T=sum(tot.values()); S=sum(samples.values())
for p in sorted(tot):
    print(f"phase {p}: {tot[p]/W:8.1f} instr/warp ({100*tot[p]/T:4.1f} %)  samples {100*samples[p]/S:4.1f} %  top:", ", ".join(f"{o} {c/W:.0f}" for o,c in ops[p].most_common(12)))
print("total per warp", T/W)
