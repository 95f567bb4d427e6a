"""Basic-block level profile from an ncu source-page CSV (development aid)."""
import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
H=rows[hdr]; si,ei,wi=H.index("Source"),H.index("Instructions Executed"),H.index("Warp Stall Sampling (All Samples)")
seen=set(); blocks=[]; cur=None
for r in rows[hdr+1:]:
    if len(r)<=ei or not r[ei].isdigit(): continue
    if r[0] in seen: continue
    seen.add(r[0])
    n=int(r[ei]); op=re.sub(r"^@!?U?P\d+\s+","",r[si].strip()).split()[0].split('.')[0]
    if cur is None or cur[0]!=n: cur=[n,collections.Counter(),r[0],0]; blocks.append(cur)
    cur[1][op]+=1; cur[3]+=int(r[wi] or 0)
tot=sum(b[0]*sum(b[1].values()) for b in blocks); stot=sum(b[3] for b in blocks)
print('total warp-inst',tot,'stall samples',stot)
for b in blocks:
    w=b[0]*sum(b[1].values())
    if w/tot>0.01 or b[3]/max(stot,1)>0.02: print(f"{100*w/tot:5.1f}% inst {100*b[3]/max(stot,1):5.1f}% samples exec={b[0]:9d} n={sum(b[1].values()):4d}  "+' '.join(f'{k}:{v}' for k,v in b[1].most_common(9)))
