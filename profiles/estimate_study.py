import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import importlib, numpy as np, pickle
synth = importlib.import_module("trgt-denovo_b200.synth")
import tdo
def run(P,T,v,h,lim=10**9):
    n,m=len(P),len(T); e=0
    while e<lim and v+e<n and h+e<m and P[v+e]==T[h+e]: e+=1
    return e
def gapc(g): return min(4+2*g, 24+g)
def estimate(P,T,mode,look=16):
    n,m=len(P),len(T)
    v=h=0; est=0; ev=0; ncmp=0
    while True:
        e=run(P,T,v,h); v+=e; h+=e; ncmp+=e//16+1
        if v>=n or h>=m:
            rest=(n-v)+(m-h)
            if rest: est+=gapc(rest); ev+=1
            return est,ev,ncmp
        need=(m-h)-(n-v)
        cands=[(8,v+1,h+1)]
        gs=set()
        if mode==0:
            for g in range(1,min(abs(m-n),13)+3): gs.add(g); gs.add(-g)
        else:
            if need: gs.add(need)
            for g in (1,2,-1,-2): gs.add(g)
            if mode>=2:
                for g in (3,-3,4,-4,5,-5,6,-6): gs.add(g)
        for g in gs:
            if g>0: cands.append((4+2*g,v,h+g))
            else: cands.append((4-2*g,v-g,h))
        best=None
        for c,v2,h2 in cands:
            if v2>n or h2>m: continue
            r=run(P,T,v2,h2,look); ncmp+=1
            fin=(v2+r>=n and h2+r>=m)
            key=(r*2 - c*0.5 + (100 if fin else 0))
            if best is None or key>best[0]: best=(key,c,v2,h2)
        _,c,v,h=best; est+=c; ev+=1
        if ev>12: return 99,ev,ncmp
if os.path.exists('/tmp/est_pairs.pkl'):
    rows=pickle.load(open('/tmp/est_pairs.pkl','rb'))
else:
    b = synth.generate(synth.config("trio_genome"), 0, 6000)
    seen = {}; al = tdo.Aligner(); rows=[]
    for i in range(b.n_pairs):
        a, c = int(b.pair_pattern[i]), int(b.pair_text[i])
        s1, s2 = b.seq(a), b.seq(c)
        if s1 == s2: continue
        k=(s1,s2)
        if k in seen: continue
        seen[k]=1
        L = abs(len(s1)-len(s2))
        al.align_end_to_end(s1, s2); pen=-al.score(); st=al.stats()
        if (L==0 and pen==8) or (L>0 and pen==4+2*L and L<=20): continue
        rows.append((s1,s2,L,pen,st.cells,i))
    pickle.dump(rows,open('/tmp/est_pairs.pkl','wb'))
print("pairs",len(rows))
def eff(order,cost,g=32):
    tot=0;mx=0
    for j in range(0,len(order),g):
        c=cost[order[j:j+g]]; tot+=c.sum(); mx+=c.max()*len(c)
    return tot/mx, mx
for mode in (0,1,2):
    r=[]
    for s1,s2,L,pen,cells,i in rows:
        est,ev,ncmp=estimate(s1,s2,mode)
        r.append((L,pen,cells,i,est,min(len(s1),len(s2)),max(len(s1),len(s2)),ncmp))
    r=np.array(r); d=r[:,4]-r[:,1]
    print("mode",mode,"est==pen %.3f est<pen %.4f |d|<=4 %.3f d>8 %.3f  mean compares %.1f"%((d==0).mean(),(d<0).mean(),(np.abs(d)<=4).mean(),(d>8).mean(), r[:,7].mean()))
    th=(r[:,0]<=10)&(r[:,5]>=16)&(r[:,6]<=304)
    x=r[th]; cost=np.where(x[:,1]<=30,x[:,2],200).astype(float)
    base=np.lexsort((x[:,3],-x[:,0]))
    print("  thread current: eff %.3f cost %.0f"%eff(base,cost))
    keep=x[:,4]<=30; x2=x[keep]; cost2=x2[:,2].astype(float)
    o=np.lexsort((x2[:,3],-x2[:,4]))
    print("  thread est-sorted, est>30 routed away (%d of %d; truly <=30: %d): eff %.3f cost %.0f"%((~keep).sum(),len(x),((~keep)&(x[:,1]<=30)).sum(),*eff(o,cost2)))
    # quad: L<14 not in thread; cost ~ pen (steps) for groups of 4
    q=r[(r[:,0]<14)&~(th&(r[:,4]<=30))]
    cq=q[:,1].astype(float)
    print("  quad n %d: current(L-sorted) eff %.3f ; est-sorted eff %.3f"%(len(q),eff(np.lexsort((q[:,3],-q[:,0])),cq,4)[0],eff(np.lexsort((q[:,3],-q[:,4])),cq,4)[0]))
    du=r[r[:,0]>=14]; cd=du[:,1].astype(float)
    print("  duo n %d: current eff %.3f ; est-sorted eff %.3f"%(len(du),eff(np.lexsort((du[:,3],-du[:,0])),cd,2)[0],eff(np.lexsort((du[:,3],-du[:,4])),cd,2)[0]))
