import sys
ev=[]
for l in open(sys.argv[1]):
    if l.startswith("PH "):
        _,h,lab,t0,t1=l.split(); ev.append((h,lab,float(t0),float(t1)))
calls=[e for e in ev if e[1]=="call"]
T1=max(e[3] for e in calls)
hs=sorted(set(e[0] for e in ev)); hid={h:i for i,h in enumerate(hs)}
lo=T1-0.75; hi=T1-0.35
W=int((hi-lo)/0.004)
rows={i:[' ']*W for i in range(len(hs))}
sym={'scan':'s','select':'e','chain':'c','dp':'D','fetch':'f'}
for h,lab,a,b in ev:
    if lab=='call':
        for x in range(max(0,int((a-lo)/0.004)), min(W,int((b-lo)/0.004)+1)):
            if rows[hid[h]][x]==' ': rows[hid[h]][x]='.'
for h,lab,a,b in ev:
    if lab in sym:
        for x in range(max(0,int((a-lo)/0.004)), min(W,int((b-lo)/0.004)+1)): rows[hid[h]][x]=sym[lab]
for i in range(len(hs)): print(f"H{i} "+''.join(rows[i]))
import collections
d=collections.defaultdict(list)
for h,lab,a,b in ev:
    if b>lo-1.0: d[lab].append(b-a)
print({k:(round(1e3*sum(v)/len(v),1), len(v)) for k,v in d.items()})
