"""Per-locus pipeline above the hot path: TSV rows, read-id file, quick mode and
--partition-by-aln, GPU path (trgt_denovo_b200.rows, every alignment + reduction on the device) against the
loop-for-loop CPU restatement (oracle/tdo_rows.py).  The bar is the north-star one: identical TSV text."""
import math
import random

import numpy as np
import pytest

from oracle import tdo, tdo_rows
from tests.util import make_family_loci


def _rows_mod():
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    return import_module("trgt_denovo_b200.rows")


def test_serialize_with_precision_matches_exact_decimal():
    R = _rows_mod()
    rng = random.Random(1)
    vals = [0.0, 1.0, 0.5, 0.25, 0.125, 0.05, 0.15, 0.25, 0.35, 0.0005, 0.0015, 1.0005, 2.5, 1e-9, 123456.789,
            0.9995, 0.99949, 0.04, 0.05, 0.95, 0.949999, math.nan, math.inf, 0.5 ** 30, 1 / 3, 2 / 3, 19 / 37]
    vals += [rng.random() * 10 ** rng.randint(-4, 3) for _ in range(3000)]
    vals += [i / 1000 for i in range(0, 2000, 7)] + [i / 16 for i in range(64)]
    for v in vals:
        assert R.serialize_with_precision(v) == tdo_rows.fmt_prec(v), v
        f = np.float32(v)
        assert R.serialize_with_precision(f) == tdo_rows.fmt_prec(f), v
    # documented examples: one decimal when it is exact to .0, else up to three without trailing zeros
    assert R.serialize_with_precision(2.0) == "2.0" and R.serialize_with_precision(0.5135) == "0.513"
    assert R.serialize_with_precision(0.25) == "0.25" and R.serialize_with_precision(0.93) == "0.93"
    # quirk kept from src/allele.rs:40-46: a value whose ONE-decimal rounding ends in ".0" is printed that way
    assert R.serialize_with_precision(0.96) == "1.0" and R.serialize_with_precision(0.0004) == "0.0"


def test_quick_mode_checks_match():
    R = _rows_mod()
    rng = random.Random(2)
    aligner = tdo.Aligner()
    for _ in range(300):
        loci, f, m, c = make_family_loci(rng, 1, missing_rate=0.0, reads_per_allele=1)
        p = R.Params()
        fs, ms, cs = (tdo_rows.load_alleles(x[0], p, aligner) for x in (f, m, c))
        ps = R.load_allele_sets(None, [f[0], m[0], c[0]], p)
        for qm in (("AL", None), ("MC", None), ("AL", 0.1), ("MC", 0.25), ("AL", 0.0)):
            assert R.trio_check_field_equivalence(ps[0], ps[1], ps[2], qm) == tdo_rows._trio_skip(fs, ms, cs, (qm[0], None))
            if qm[1] is not None:
                assert R.trio_check_field_similarity(ps[0], ps[1], ps[2], qm) == tdo_rows._trio_skip(fs, ms, cs, qm)
                assert R.duo_check_field_similarity(ps[2], ps[1], qm) == tdo_rows._duo_skip(cs, ms, qm)
            assert R.duo_check_field_equivalence(ps[2], ps[1], qm) == tdo_rows._duo_skip(cs, ms, (qm[0], None))


def test_dropout_and_ploidy():
    R = _rows_mod()
    loc = R.Locus("chrX", 1, 2, ["A"], "t")
    mk = lambda hp, k: R.AlleleSet(alleles=[], hp_counts=hp, karyotype=k)
    assert R.get_naive_dropout(mk((0, 1, 0), "XX"), loc) == "FD"
    assert R.get_naive_dropout(mk((5, 0, 1), "XX"), loc) == "HD"
    assert R.get_naive_dropout(mk((5, 0, 1), "XY"), loc) == "N"  # ploidy one on chrX for XY
    assert R.get_naive_dropout(mk((5, 0, 2), "XX"), loc) == "N"
    assert R.get_ploidy("XX", "chrY") == 0 and R.get_ploidy("weird", "chr2") == 2


@pytest.fixture(scope="module")
def gpu_ctx():
    import __graft_entry__ as entry
    entry.build()
    import trgt_denovo_b200 as m
    assert m.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    c = m.Context(device_id=0)
    yield c
    c.close()


CASES = [
    dict(),
    dict(clip_len=0, parent_quantile=0.9),
    dict(partition_by_alignment=True),
    dict(quick_mode=("AL", None)),
    dict(quick_mode=("MC", 0.1), partition_by_alignment=True, parent_quantile=0.5),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_trio_tsv_identical(gpu_ctx, case):
    R = _rows_mod()
    params = R.Params(**CASES[case])
    rng = random.Random(100 + case)
    loci, f, m, c = make_family_loci(rng, 120)
    got = R.process_trio_batch(gpu_ctx, loci, f, m, c, params)
    aligner = tdo.Aligner()
    want = [tdo_rows.process_alleles_trio(loci[i], f[i], m[i], c[i], params, aligner) for i in range(len(loci))]
    wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
    assert R.format_tsv(got, R.TRIO_COLUMNS) == wtsv
    assert R.format_read_ids(got) == wids
    assert wtsv.count("\n") > len(loci) and ("Y:" in wtsv) and ("\t.\t" in wtsv)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 2, 3])
def test_duo_tsv_identical(gpu_ctx, case):
    R = _rows_mod()
    params = R.Params(**CASES[case])
    rng = random.Random(200 + case)
    loci, _, m, c = make_family_loci(rng, 120, duo=True)
    got = R.process_duo_batch(gpu_ctx, loci, c, m, params)
    aligner = tdo.Aligner()
    want = [tdo_rows.process_alleles_duo(loci[i], c[i], m[i], params, aligner) for i in range(len(loci))]
    wtsv, wids = tdo_rows.write_outputs(want, R.DUO_COLUMNS)
    assert R.format_tsv(got, R.DUO_COLUMNS) == wtsv
    assert R.format_read_ids(got) == wids


@pytest.mark.gpu
def test_read_ids_through_the_sparse_mask_path(monkeypatch):
    """Large host calls fetch the de novo read mask as a list of set positions instead of one byte per pair
    (forced here for a small batch): the read-id file must not change."""
    import __graft_entry__ as entry
    entry.build()
    import trgt_denovo_b200 as tdb
    R = _rows_mod()
    monkeypatch.setenv("TDB_SPARSE_MASK_MIN_PAIRS", "0")
    ctx = tdb.Context(device_id=0)
    monkeypatch.delenv("TDB_SPARSE_MASK_MIN_PAIRS")
    params = R.Params(**CASES[0])
    rng = random.Random(300)
    loci, f, m, c = make_family_loci(rng, 150, denovo_rate=0.5)
    got = R.process_trio_batch(ctx, loci, f, m, c, params)
    aligner = tdo.Aligner()
    want = [tdo_rows.process_alleles_trio(loci[i], f[i], m[i], c[i], params, aligner) for i in range(len(loci))]
    wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
    assert R.format_tsv(got, R.TRIO_COLUMNS) == wtsv
    assert R.format_read_ids(got) == wids and len(wids) > 0
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_gpus", [2, 3])
def test_multi_gpu_rows_equal_one_gpu(n_gpus):
    """shard.MultiGpu (one ctx + host thread per GPU, balanced contiguous locus ranges, rows merged in locus order):
    TSV and read-id text identical to one context and to the CPU restatement.  With fewer visible devices than
    contexts they share devices, so the sharding / merging is checked on any box; with gpurun --gpus 2 the two
    contexts sit on two GPUs."""
    import __graft_entry__ as entry
    entry.build()
    from importlib import import_module
    import trgt_denovo_b200 as tdb
    R = _rows_mod()
    SH = import_module("trgt_denovo_b200.shard")
    params = R.Params()
    rng = random.Random(400 + n_gpus)
    loci, f, m, c = make_family_loci(rng, 200, denovo_rate=0.3)
    mg = SH.MultiGpu(n_gpus)
    got = mg.process_trio_batch(loci, f, m, c, params)
    assert len(mg.bounds) == n_gpus + 1 and all(b > a for a, b in zip(mg.bounds, mg.bounds[1:]))
    if tdb.device_count() >= n_gpus:
        assert sorted({x.device_id for x in mg.ctxs}) == list(range(n_gpus))
    one = tdb.Context(device_id=0)
    single = R.process_trio_batch(one, loci, f, m, c, params)
    aligner = tdo.Aligner()
    want = [tdo_rows.process_alleles_trio(loci[i], f[i], m[i], c[i], params, aligner) for i in range(len(loci))]
    wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
    assert R.format_tsv(got, R.TRIO_COLUMNS) == R.format_tsv(single, R.TRIO_COLUMNS) == wtsv
    assert R.format_read_ids(got) == R.format_read_ids(single) == wids and len(wids) > 0
    got = mg.process_duo_batch(loci, c, m, params)
    single = R.process_duo_batch(one, loci, c, m, params)
    assert R.format_tsv(got, R.DUO_COLUMNS) == R.format_tsv(single, R.DUO_COLUMNS)
    mg.close()
    one.close()
