"""CPU (oracle) study: final penalty and wavefront width of the distinct non-trivial pairs with |n-m| < 16 (what sizes the thread tier)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import importlib, numpy as np
pkg = importlib.import_module("trgt-denovo_b200")
from importlib import import_module
synth = import_module("trgt-denovo_b200.synth")
import tdo
p = synth.config("trio_genome")
b = synth.generate(p, 0, 4000)
print(b.n_pairs, b.n_seqs)
seen = {}
al = tdo.Aligner()
hist = collections.Counter()
tot = 0
for i in range(b.n_pairs):
    a, c = int(b.pair_pattern[i]), int(b.pair_text[i])
    s1, s2 = b.seq(a), b.seq(c)
    if s1 == s2: continue
    k = (s1, s2)
    if k in seen: continue
    seen[k] = 1
    L = abs(len(s1) - len(s2))
    if L >= 16: 
        hist[('warp', 0)] += 1; continue
    al.align_end_to_end(s1, s2)
    pen = -al.score()
    st = al.stats()
    if (L == 0 and pen == 8) or (L > 0 and pen == 4 + 2 * L and L <= 20):
        hist[('cf', 0)] += 1; continue
    hist[('q', min(pen, 60) // 4 * 4)] += 1
    hist[('qcells', min(pen, 60) // 4 * 4)] += st.cells
    hist[('qw', min(st.max_width, 40) // 4 * 4)] += 1
    tot += 1
print("unique", len(seen), "quad-nontrivial", tot, "cf", hist[('cf',0)], "warp", hist[('warp',0)])
cum = 0; cc = 0; totc = sum(v for (k, _), v in hist.items() if k == 'qcells')
for pen in range(0, 64, 4):
    cum += hist[('q', pen)]; cc += hist[('qcells', pen)]
    print(f"pen {pen:2d}-{pen+3:2d}: {hist[('q',pen)]:6d} cum {100*cum/tot:5.1f}%  cells cum {100*cc/totc:5.1f}%")
for w in range(0, 44, 4):
    print("maxwidth", w, hist[('qw', w)])
