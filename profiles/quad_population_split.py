"""CPU (oracle) study: which pairs the quad tier is left with once the thread tier exists (thread / deferred / |n-m| 8-15 / long)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import importlib, numpy as np
synth = importlib.import_module("trgt-denovo_b200.synth")
import tdo
b = synth.generate(synth.config("trio_genome"), 0, 4000)
seen = {}
al = tdo.Aligner()
cat = collections.Counter(); cells = collections.Counter(); pens = collections.defaultdict(list)
for i in range(b.n_pairs):
    a, c = int(b.pair_pattern[i]), int(b.pair_text[i])
    s1, s2 = b.seq(a), b.seq(c)
    if s1 == s2: continue
    k = (s1, s2)
    if k in seen: continue
    seen[k] = 1
    n, m = len(s1), len(s2); L = abs(n - m)
    if L >= 16: continue
    al.align_end_to_end(s1, s2)
    pen = -al.score(); st = al.stats()
    if (L == 0 and pen == 8) or (L > 0 and pen == 4 + 2 * L): continue   # closed form
    elig = L < 8 and 16 <= n <= 304 and 16 <= m <= 304
    if elig and pen <= 24: c_ = 'thread'
    elif elig: c_ = 'deferred'
    elif L >= 8: c_ = 'L8-15'
    else: c_ = 'long/short'
    cat[c_] += 1; cells[c_] += st.cells; pens[c_].append(pen)
tot = sum(cat.values()); tc = sum(cells.values())
for k in cat: print(k, cat[k], round(100*cat[k]/tot,1), '% pairs', round(100*cells[k]/tc,1), '% cells', 'pen pct', np.percentile(pens[k],[10,50,90]))
