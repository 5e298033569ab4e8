import numpy as np, sys, random
sys.path.insert(0,'/root/repo')
from oracle import tdo
MAXG=int(sys.argv[1]) if len(sys.argv)>1 else 20
def closed_form(P,T,clip):
    n,m=len(P),len(T); L=abs(m-n)
    if L>MAXG: return None
    p=0; lim=min(n,m)
    while p<lim and P[p]==T[p]: p+=1
    if L==0:
        if p>=n: return None
        if P[p+1:]!=T[p+1:]: return None
        cols=n; return 8, (-8 if (clip<=p<cols-clip) else 0)
    if m>n:
        if P[p:]!=T[p+L:]: return None
        cols=m
    else:
        if P[p+L:]!=T[p:]: return None
        cols=n
    pen=min(4+2*L,24+L)
    a=max(p,clip); b=min(p+L,cols-clip); ln=b-a
    return pen, (-min(4+2*ln,24+ln) if ln>0 else 0)
rng=random.Random(11)
a=tdo.Aligner(); tot=bad=0
def rs(n): return "".join(rng.choice("ACGT") for _ in range(n))
for it in range(60000):
    k=rng.randint(1,6); motif=rs(k)
    mode=rng.random()
    lf=rs(rng.randint(0,80)); rf=rs(rng.randint(0,80))
    if mode<0.5:   # STR expansion/contraction by whole units, long repeats so the heuristic can act
        u=rng.randint(0,120); d=rng.randint(1,max(1,MAXG//k))
        T=lf+motif*u+rf; P=lf+motif*(u+d if rng.random()<0.5 else max(0,u-d))+rf
    elif mode<0.7: # compound repeat
        m2=rs(rng.randint(1,6)); u=rng.randint(5,60); u2=rng.randint(5,60); d=rng.randint(1,max(1,MAXG//k))
        T=lf+motif*u+m2*u2+rf; P=lf+motif*(u+d)+m2*u2+rf
    elif mode<0.85: # random insertion / deletion anywhere
        T=lf+motif*rng.randint(0,80)+rf or "A"; i=rng.randrange(len(T)+1); d=rng.randint(1,MAXG)
        P=T[:i]+rs(d)+T[i:] if rng.random()<0.5 else T[:i]+T[i+d:]
    else: # gap of repeat-like content inside a long homopolymer-ish region
        b=rng.choice("ACGT"); u=rng.randint(20,200); d=rng.randint(1,MAXG)
        T=lf+b*u+rf; P=lf+b*(u+d)+rf
    if not P or not T or P==T: continue
    P=P.encode(); T=T.encode()
    r0=closed_form(P,T,50)
    if r0 is None: continue
    a.align_end_to_end(P,T)
    for clip in (50,0,7):
        pen,sc=closed_form(P,T,clip); tot+=1
        if -a.score()!=pen or a.cigar_score_clipped(clip)!=sc:
            bad+=1
            if bad<6: print("MISMATCH",clip,pen,sc,-a.score(),a.cigar_score_clipped(clip),a.cigar(),len(P),len(T))
print("checked",tot,"bad",bad)
