import csv,sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
exits=[n for n,r in enumerate(data) if 'EXIT' in r[ix['Source']]]
print('exits',exits)
ntop=int(sys.argv[2]) if len(sys.argv)>2 else 14
def top(lo,hi,name):
    tot=Counter(); total=0; lines=[]
    for n in range(lo,hi):
        r=data[n]
        s=int(r[ix['# Samples']] or 0); total+=s
        lines.append((s,n,r))
        for st in stalls: tot[st]+=int(r[ix[st]] or 0)
    print('==',name,lo,hi,'samples',total, [(k,v) for k,v in tot.most_common(6)])
    for s,n,r in sorted(lines,reverse=True)[:ntop]:
        t=sorted(((int(r[ix[st]] or 0),st) for st in stalls),reverse=True)[:2]
        print('   ',n,s,r[ix['Source']].strip()[:80],t)
b=[0]+exits+[len(data)]
for i in range(len(b)-1):
    if b[i+1]-b[i]>50: top(b[i],b[i+1],'seg%d'%i)
