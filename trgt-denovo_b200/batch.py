"""Batch assembly: the packing role north_star assigns to locus.rs / read.rs.

A SeqBatch concatenates the raw ASCII sequences (`read.bases`, src/read.rs:57; `allele.seq` =
left_flank + allele + right_flank, src/allele.rs:406-426) of many loci into one blob plus offset /
length tables, and records (pattern, text) pairs in the order the reference's loops visit them
(src/denovo.rs:29-37,48-52).  2-bit packing itself happens on the GPU (k_pack).
"""
import numpy as np


class SeqBatch:
    def __init__(self):
        self._chunks = []
        self._lens = []
        self._pp = []
        self._pt = []

    def add_seq(self, seq: bytes) -> int:
        self._chunks.append(seq)
        self._lens.append(len(seq))
        return len(self._lens) - 1

    def add_pair(self, pattern_idx: int, text_idx: int) -> int:
        self._pp.append(pattern_idx)
        self._pt.append(text_idx)
        return len(self._pp) - 1

    @property
    def n_pairs(self):
        return len(self._pp)

    def arrays(self):
        """-> blob uint8, seq_off (None: dense layout), seq_len uint32, pair_pattern uint32, pair_text uint32"""
        blob = np.frombuffer(b"".join(self._chunks), dtype=np.uint8)
        lens = np.array(self._lens, dtype=np.uint32)
        offs = None  # dense layout: sequences back to back, the library derives the offsets
        return (blob, offs, lens, np.array(self._pp, dtype=np.uint32), np.array(self._pt, dtype=np.uint32))
