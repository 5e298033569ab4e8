"""CPU (oracle) study: the pairs of tier 1 (|n-m| >= 16): penalty in excess of the single-gap cost."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import importlib, numpy as np
synth = importlib.import_module("trgt-denovo_b200.synth")
import tdo
b = synth.generate(synth.config("trio_genome"), 0, 6000)
seen = {}
al = tdo.Aligner()
rows = []
for i in range(b.n_pairs):
    a, c = int(b.pair_pattern[i]), int(b.pair_text[i])
    s1, s2 = b.seq(a), b.seq(c)
    if s1 == s2: continue
    L = abs(len(s1) - len(s2))
    if L < 16: continue
    k = (s1, s2)
    if k in seen: continue
    seen[k] = 1
    al.align_end_to_end(s1, s2)
    pen = -al.score(); st = al.stats()
    rows.append((L, pen, st.cells, st.max_width, st.steps, pen - min(4+2*L, 24+L)))
r = np.array(rows)
print("n", len(r))
print("L pct", np.percentile(r[:,0],[10,50,90,99]), "pen pct", np.percentile(r[:,1],[10,50,90,99]))
print("cells mean", r[:,2].mean(), "width pct", np.percentile(r[:,3],[10,50,90,99]), "steps", np.percentile(r[:,4],[50,90]))
ex = r[:,5]
for v in (0, 6, 8, 12, 14, 16, 24):
    print("excess <=", v, round(float((ex<=v).mean()),3))
print("L<40", float((r[:,0]<40).mean()), "cells share L<40", r[r[:,0]<40,2].sum()/r[:,2].sum())
cf = (r[:,5]==0) & (r[:,0]<=20)
q = r[~cf]
print("after closed form:", len(q), "of", len(r))
print("L>20 & excess 0:", float(((q[:,0]>20)&(q[:,5]==0)).mean()))
for v in (6, 8, 12, 16):
    print("excess ==", v, round(float((q[:,5]==v).mean()),3))
print("L dist", np.percentile(q[:,0],[10,50,90]), "pen", np.percentile(q[:,1],[10,50,90]), "cells mean", q[:,2].mean())
import collections
print(sorted(collections.Counter(q[:,5].tolist()).items()))
print("L>20 share", float((q[:,0]>20).mean()))
