"""CPU: the host half of the packing role (tdb_host_pack) against a plain numpy statement of the layout."""
import random

import numpy as np
import pytest

import __graft_entry__ as entry
from tests.util import rand_seq, to_arrays


@pytest.fixture(scope="module")
def L():
    entry.build()
    import trgt_denovo_b200 as tdb
    return tdb.lib()


def _reference_pack(seqs, off):
    blob = b"".join(seqs)
    out = np.zeros(len(blob) // 16 + 2, dtype=np.uint32)
    flags = np.zeros(len(seqs), dtype=np.uint8)
    for x, c in enumerate(blob):
        out[x // 16] |= np.uint32(((c >> 1) & 3) << (2 * (x % 16)))
    for i, s in enumerate(seqs):
        if any(c not in b"ACGT" for c in s):
            flags[i] = 1
    return out, flags


def _pack(L, seqs, threads=3):
    blob, off, ln = to_arrays(seqs)
    n = L.tdb_packed_words(blob.size)
    packed = np.full(n, 0xdeadbeef, dtype=np.uint32)
    flag = np.full(len(seqs), 7, dtype=np.uint8)
    r = L.tdb_host_pack(blob.ctypes.data, blob.size, off.ctypes.data, ln.ctypes.data, len(seqs), packed.ctypes.data,
                        flag.ctypes.data, threads)
    return r, packed, flag, off, blob.size


def test_host_pack_matches_layout(L):
    rng = random.Random(3)
    lens = [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 1000, 0, 0, 5] + \
        [rng.randint(1, 400) for _ in range(300)]
    seqs = [rand_seq(rng, n).encode() for n in lens]
    seqs[5] = seqs[5][:10] + b"N" + seqs[5][11:]
    seqs[20] = seqs[20].lower()
    seqs[18] = b"ACGNT"  # right after two empty sequences sharing its offset
    for cut in (None, 1, 37, 700):  # different blob tails (partial words / partial vectors)
        ss = seqs if cut is None else seqs[:-1] + [seqs[-1][:cut]]
        r, packed, flag, off, nbytes = _pack(L, ss)
        assert r == 0
        want, wflag = _reference_pack(ss, off)
        assert (flag == wflag).all()
        nw = (nbytes + 15) // 16
        assert (packed[:nw] == want[:nw]).all()
    assert L.tdb_host_pack_isa() in (b"avx512bw", b"avx2", b"scalar")


def test_host_pack_rejects_bad_layout(L):
    blob = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
    ln = np.array([10, 10], np.uint32)
    packed = np.zeros(L.tdb_packed_words(blob.size), np.uint32)
    flag = np.zeros(2, np.uint8)
    for offs in ([0, 5], [0, 35]):
        off = np.array(offs, np.uint64)
        assert L.tdb_host_pack(blob.ctypes.data, blob.size, off.ctypes.data, ln.ctypes.data, 2, packed.ctypes.data,
                               flag.ctypes.data, 1) == -1


@pytest.mark.parametrize("isa", ["avx2", "scalar"])
def test_host_pack_other_isas(L, isa):
    """The dispatch is fixed per process: run the layout test in a child with the ISA capped."""
    import os
    import subprocess
    import sys
    code = ("import os,sys; sys.path.insert(0, %r); import tests.test_host_pack as t; import trgt_denovo_b200 as tdb; "
            "L = tdb.lib(); t.test_host_pack_matches_layout(L); assert L.tdb_host_pack_isa() in (%r.encode(), b'scalar')"
            % (entry.ROOT, isa))
    env = dict(os.environ, TDB_HOSTPACK_ISA=isa)
    subprocess.check_call([sys.executable, "-c", code], env=env, cwd=entry.ROOT)
