import numpy as np, sys, random
sys.path.insert(0,'/root/repo')
from importlib import import_module
import trgt_denovo_b200
from oracle import tdo
from tests.util import make_pairs
S=import_module("trgt_denovo_b200.synth")

def closed_form(P,T,clip):
    n,m=len(P),len(T); L=abs(m-n)
    if L>5: return None
    p=0; lim=min(n,m)
    while p<lim and P[p]==T[p]: p+=1
    if L==0:
        if p>=n: return None
        if P[p+1:]!=T[p+1:]: return None
        cols=n; pen=8
        sc = -8 if (clip<=p<cols-clip) else 0
        return pen,sc
    if m>n:
        if P[p:]!=T[p+L:]: return None
        cols=m
    else:
        if P[p+L:]!=T[p:]: return None
        cols=n
    pen=4+2*L
    a=max(p,clip); b=min(p+L,cols-clip); ln=b-a
    sc = -min(4+2*ln,24+ln) if ln>0 else 0
    return pen,sc

def run(pairs, clips=(50,0,3)):
    a=tdo.Aligner(); tot=0; bad=0
    for P,T in pairs:
        r0=closed_form(P,T,50)
        if r0 is None: continue
        a.align_end_to_end(P,T)
        for clip in clips:
            pen,sc=closed_form(P,T,clip)
            tot+=1
            if -a.score()!=pen or a.cigar_score_clipped(clip)!=sc:
                bad+=1
                if bad<5: print("MISMATCH",clip,pen,sc,-a.score(),a.cigar_score_clipped(clip),a.cigar(),len(P),len(T))
    return tot,bad

b=S.generate(S.config("trio_genome"),0,2500,8)
blob=b.blob.tobytes()
def seq(i):
    o=int(b.seq_off[i]); return blob[o:o+int(b.seq_len[i])]
seen=set()
for j in range(b.n_pairs):
    P=seq(int(b.pair_pattern[j])); T=seq(int(b.pair_text[j]))
    if P!=T: seen.add((P,T))
print("synthetic trio:", run(seen))
rng=random.Random(5); 
for err in (0.0005,0.003,0.01):
    seqs,pp,pt=make_pairs(rng,400,12,max_units=30,err=err)
    print("make_pairs err",err, run({(seqs[p],seqs[t]) for p,t in zip(pp,pt)}))
# adversarial: homopolymers / short / gaps near ends
adv=set()
for _ in range(30000):
    k=rng.randint(1,6); motif="".join(rng.choice("ACGT") for _ in range(k))
    lf="".join(rng.choice("ACGT") for _ in range(rng.randint(0,60))); rf="".join(rng.choice("ACGT") for _ in range(rng.randint(0,60)))
    u=rng.randint(0,30); T=lf+motif*u+rf
    mode=rng.random()
    if mode<0.4:
        d=rng.randint(-5,5); u2=max(0,u+ (d//k if k else 0)); P=lf+motif*u2+rf
    elif mode<0.7 and len(T)>0:
        i=rng.randrange(len(T)); P=T[:i]+rng.choice("ACGT")+T[i+1:]
    else:
        i=rng.randrange(len(T)+1); d=rng.randint(1,5)
        P = T[:i]+"".join(rng.choice("ACGT") for _ in range(d))+T[i:] if rng.random()<0.5 else T[:i]+T[i+d:]
    if P and T and P!=T: adv.add((P.encode(),T.encode()))
print("adversarial:", run(adv))
