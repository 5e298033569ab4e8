#!/usr/bin/env python3
"""Turns the files profiles/run_gpu_final.sh brought back (gpurun_out/) into the committed summaries
under profiles/r01/ and the per-pair DRAM traffic bench.py quotes in roofline.traffic."""
import csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01g"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles", "r01")
for f in (f"bench_full_{tag}.json", f"bench_ref_{tag}.json", f"launches_{tag}.csv", f"launches_full_{tag}.csv"):
    if os.path.exists(os.path.join(G, f)):
        subprocess.check_call(["cp", os.path.join(G, f), os.path.join(P, f)])
for name, what in ((f"launches_{tag}", "first 120 launches of: python bench.py --loci 100000 --steps 1 --warmup 3 --no-cpu-baseline"),
                   (f"launches_full_{tag}", "first 150 launches of: python bench.py --steps 3 --warmup 3 --no-cpu-baseline (the judged 937k-locus configuration)")):
    if not os.path.exists(os.path.join(G, name + ".csv")):
        continue
    rows = list(csv.reader(open(os.path.join(G, name + ".csv"))))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
    t, c = {}, {}
    for r in rows[h + 1:]:
        if len(r) > vi:
            k = r[ki][:70]
            t[k] = t.get(k, 0.0) + float(r[vi].replace(",", "")); c[k] = c.get(k, 0) + 1
    tot = sum(t.values())
    with open(os.path.join(P, name + "_summary.txt"), "w") as o:
        o.write(f"# ncu --metrics gpu__time_duration.sum, {what}\n# (cold-cache, serialised: compare shares)\n")
        for k, v in sorted(t.items(), key=lambda x: -x[1]):
            o.write(f"{v/1e6:10.3f} ms {100*v/tot:5.1f}% n={c[k]:3d} {k}\n")
plain = json.loads([l for l in open(os.path.join(G, f"plain30k_{tag}.log")) if l.startswith("{")][-1])
traffic = {}
names = {"quad": "k_align_lock<QuadCfg>", "duo": "k_align_lock<DuoCfg>", "thread": "k_align_thread"}
for kern, ti in (("quad", 0), ("duo", 1), ("thread", -1)):
    rep = os.path.join(G, f"prof_{kern}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[-1]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    keys = open(os.path.join(R, "profiles", "ncu_key_metrics.py")).read()
    with open(os.path.join(P, f"prof_{kern}_{tag}_key_metrics.txt"), "w") as o:
        subprocess.run([sys.executable, os.path.join(R, "profiles", "ncu_key_metrics.py")], input=raw, text=True, stdout=o)
    def to_bytes(name):
        x = float(d[name].replace(",", "")); un = u[name].lower()
        return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(un, 1)
    b = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    pairs = plain["tier_pairs"][ti] if ti >= 0 else plain["thread_tier_pairs"]
    traffic[names[kern]] = {"dram_bytes_per_launch": b, "pairs_in_launch": pairs, "dram_bytes_per_pair": b / max(1, pairs),
                                  "source": f"ncu --set full, prof_{kern}_{tag}.ncu-rep (30 000-locus step)"}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
