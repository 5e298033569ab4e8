import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import importlib, numpy as np
synth = importlib.import_module("trgt-denovo_b200.synth")
import tdo
b = synth.generate(synth.config("trio_genome"), 0, 8000)
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
    if L>10 or min(len(s1),len(s2))<16 or max(len(s1),len(s2))>304: continue
    # prefix / suffix
    n,m=len(s1),len(s2); p=0
    while p<min(n,m) and s1[p]==s2[p]: p+=1
    q=0
    while q<min(n,m)-p and s1[n-1-q]==s2[m-1-q]: q+=1
    rows.append((L,min(pen,32),st.cells if pen<=30 else 200,i,p,q,n,m,st.ext16))
r=np.array(rows)
print("thread-tier pairs",len(r))
def eff(order,col=2):
    x=r[order]; tot=0; mx=0
    for j in range(0,len(x),32):
        g=x[j:j+32,col]; tot+=g.sum(); mx+=g.max()*32
    return tot/mx
base=np.lexsort((r[:,3], -r[:,0]))   # by L desc, then original order
print("current order (L, locus): cells eff %.3f  pen eff %.3f ext eff %.3f"%(eff(base),eff(base,1),eff(base,8)))
o2=np.lexsort((r[:,1], -r[:,0]))
print("oracle (L, pen): cells eff %.3f pen eff %.3f"%(eff(o2),eff(o2,1)))
o3=np.argsort(r[:,2])
print("oracle by cells: %.3f"%eff(o3))
# estimator: middle length
mid=np.minimum(r[:,6],r[:,7])-r[:,4]-r[:,5]
o4=np.lexsort((mid, -r[:,0]))
print("(L, mid): cells eff %.3f pen eff %.3f"%(eff(o4),eff(o4,1)))
import collections
for L in (0,1,2,4,6):
    x=r[r[:,0]==L]
    print("L",L,"n",len(x),"pen hist",sorted(collections.Counter(x[:,1].tolist()).items()))

# multi-pass simulation: caps per L; truncated cost ~ cells * (cap/pen)^2
def sim(capsf):
    total=0.0
    cur=r.copy()
    pen=cur[:,1].astype(float); cells=cur[:,2].astype(float); L=cur[:,0]
    alive=np.ones(len(cur),bool)
    for pi in range(3):
        caps=capsf(L,pi)
        idx=np.nonzero(alive)[0]
        idx=idx[np.lexsort((cur[idx,3], -L[idx]))]
        cost=np.where(pen[idx]<=caps[idx], cells[idx], cells[idx]*(caps[idx]/np.maximum(pen[idx],1))**2)
        for j in range(0,len(idx),32):
            total+=cost[j:j+32].max()*32
        alive[idx[pen[idx]<=caps[idx]]]=False
    return total
cur_total=sum(r[base][j:j+32,2].max()*32 for j in range(0,len(r),32))
print("current cost",cur_total, "ideal", r[:,2].sum())
for name,f in [("2L+8,2L+12,30", lambda L,pi: [np.where(L==0,12,2*L+8), np.where(L==0,16,2*L+12), np.full(len(L),30)][pi]),
               ("2L+12,30,30", lambda L,pi: [np.where(L==0,16,2*L+12), np.full(len(L),30), np.full(len(L),30)][pi]),
               ("2L+8,30,30", lambda L,pi: [np.where(L==0,12,2*L+8), np.full(len(L),30), np.full(len(L),30)][pi])]:
    print(name, sim(f))
