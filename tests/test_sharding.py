"""CPU, world_size 2 over gloo: the multi-GPU path is "shard the loci, no data-path collective".

Each rank generates its own locus range (the generator is a pure function of (seed, locus index)),
scores it with the CPU oracle (this is a test: the oracle is the checker), and the ranks only exchange
a checksum and their row counts -- exactly what bench.py does with its timings.  The union of the
shards must equal the single-process result for the whole range.
"""
import os
import subprocess
import sys

import numpy as np

import __graft_entry__ as entry

WORKER = r'''
import os, sys, zlib
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["TDB_ROOT"])
import __graft_entry__ as entry
entry.build()
from importlib import import_module
import trgt_denovo_b200
from oracle import tdo
S = import_module("trgt_denovo_b200.synth"); SH = import_module("trgt_denovo_b200.shard")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = int(os.environ["TDB_N_LOCI"])
b = SH.even_bounds(N, world)
batch = S.generate(S.config("trio_1k"), b[rank], b[rank + 1], 2)
scores, status, _ = tdo.align_batch(batch.blob, batch.seq_off, batch.seq_len, batch.pair_pattern, batch.pair_text, 50, n_threads=2)
out, mask = tdo.assess_batch(batch.loci_ptr, batch.n_loci, scores, 1.0, duo=False, n_threads=2)
cov = np.array([out[2 * i + c].denovo_coverage for i in range(batch.n_loci) for c in range(batch.loci[i].n_child)], np.int64)
mine = torch.tensor([batch.n_loci, batch.n_pairs, int(scores.astype(np.int64).sum()), int(cov.sum()),
                     zlib.crc32(scores.tobytes())], dtype=torch.int64)
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)          # bookkeeping only: no sequence or score data crosses ranks
if rank == 0:
    print("RESULT", [v.tolist() for v in allv], flush=True)
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_shards_cover_the_whole_range(tmp_path):
    entry.build()
    n_loci = 120
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TDB_ROOT=entry.ROOT, TDB_N_LOCI=str(n_loci), MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    parts = eval(line[len("RESULT"):])
    # single-process result on the whole range
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    from oracle import tdo
    S = import_module("trgt_denovo_b200.synth")
    b = S.generate(S.config("trio_1k"), 0, n_loci, 2)
    scores, _, _ = tdo.align_batch(b.blob, b.seq_off, b.seq_len, b.pair_pattern, b.pair_text, 50, n_threads=2)
    out2, _ = tdo.assess_batch(b.loci_ptr, b.n_loci, scores, 1.0, duo=False, n_threads=2)
    cov = sum(out2[2 * i + c].denovo_coverage for i in range(b.n_loci) for c in range(b.loci[i].n_child))
    assert sum(p[0] for p in parts) == n_loci
    assert sum(p[1] for p in parts) == b.n_pairs
    assert sum(p[2] for p in parts) == int(scores.astype(np.int64).sum())
    assert sum(p[3] for p in parts) == cov


def test_balanced_bounds():
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    SH = import_module("trgt_denovo_b200.shard")
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1000, size=5000)
    for world in (1, 2, 4, 8):
        b = SH.balanced_bounds(w, world)
        assert b[0] == 0 and b[-1] == len(w) and all(x <= y for x, y in zip(b, b[1:]))
        sums = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(sums) - min(sums) <= 2 * w.max()
        assert SH.even_bounds(len(w), world)[-1] == len(w)
    assert SH.balanced_bounds([], 4) == [0, 0, 0, 0, 0]
    parts = [(50, ["c"]), (0, ["a", "b"])]
    assert SH.merge_in_locus_order(parts) == ["a", "b", "c"]
