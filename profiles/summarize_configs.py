#!/usr/bin/env python3
"""profiles/bench_configs.py output (jsonl) -> the markdown tables committed next to it."""
import json, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [json.loads(l) for l in open(src) if l.startswith("{")]
o = open(dst, "w")
o.write("# BASELINE configs 1, 3, 4, 5 on one B200 (profiles/bench_configs.py, e2e through the host-buffer C ABI, dense layout)\n\n")
o.write("| config | loci | pairs | e2e loci/s | e2e alignments/s | CPU port (%d thr, align only) loci/s | ratio | thread tier / quad / tier 1 / long / global | heuristic-sensitive pairs | bit-exact |\n" % rows[0]["cpu_port"]["threads"])
o.write("|---|---|---|---|---|---|---|---|---|---|\n")
for r in rows:
    if "cpu_port" not in r:
        continue
    hc = r.get("heuristic_census")
    o.write("| %s | %d | %d | %.3g | %.3g | %.3g | %.1fx | %d / %s | %s | %s |\n" % (
        r["config"], r["loci"], r["pairs"], r["e2e_loci_per_s"], r["e2e_alignments_per_s"], r["cpu_port"]["loci_per_s"],
        r["e2e_loci_per_s"] / r["cpu_port"]["loci_per_s"], r.get("thread_tier_pairs", 0),
        " / ".join(str(x) for x in (r["tier_pairs"][0], r["tier_pairs"][1], r.get("long_tier_pairs", 0), r["tier_pairs"][2] + r["tier_pairs"][3])),
        ("%d of %d (clipped score %d, penalty %d)" % (hc["heuristic_sensitive_pairs"], hc["pairs"], hc["clipped_score_differs"], hc["penalty_differs"])) if hc else "-",
        r["scores_bit_exact_on_cpu_sample"]))
micro = [r for r in rows if r["config"].startswith("5")]
for h in sorted({r["heuristic"] for r in micro}, reverse=True):
    sub = [r for r in micro if r["heuristic"] == h]
    divs = sorted({r["divergence"] for r in sub})
    o.write("\n## config 5 microbenchmark, heuristic %s: e2e GCUPS on B200 (CPU port GCUPS, %d threads)\n\n" % (h, sub[0]["cpu_threads"]))
    o.write("| length | " + " | ".join("%g%%" % (100 * d) for d in divs) + " |\n|---|" + "---|" * len(divs) + "\n")
    for L in sorted({r["length"] for r in sub}):
        cells = []
        for d in divs:
            r = next(x for x in sub if x["length"] == L and x["divergence"] == d)
            cells.append("%.0f (%.0f)" % (r["e2e_gcups"], r["cpu_port_gcups"]))
        o.write("| %d | " % L + " | ".join(cells) + " |\n")
o.close()
assert all(r["scores_bit_exact_on_cpu_sample"] for r in rows)
print(open(dst).read())
