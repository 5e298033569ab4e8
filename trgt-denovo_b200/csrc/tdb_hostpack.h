// tdb_hostpack.h -- host side of the packing role north_star gives locus.rs / read.rs: raw ASCII
// sequences -> the 2-bit word layout the kernels read, produced by the caller's CPU threads so that
// a host batch crosses PCIe at a quarter of its ASCII size.  Pure data re-encoding: no alignment
// arithmetic happens on the host.  Bit layout identical to k_pack (tdb_kernels.cuh): base k of a
// sequence sits in word k/16 at bits 2*(k%16), code = (byte >> 1) & 3, bytes past the end are 0, one
// zero word follows the last base; the word offset of sequence i is (seq_off[i] >> 4) + 2 i.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace tdb {

// Packs sequences [s0, s1).  `packed` is the full-batch word array (indexed by absolute word offset).
// flag[i] = 1 when sequence i holds a byte outside {A,C,G,T} (its pairs use the raw bytes on the
// device), 2 when it violates the layout contract (also reported through the return value).
// Returns the number of layout violations (sequence outside the blob or overlapping its successor).
size_t hostpack_range(const uint8_t *blob, size_t blob_bytes, const uint64_t *seq_off, const uint32_t *seq_len,
                      uint32_t s0, uint32_t s1, uint32_t n_seqs, uint32_t *packed, uint8_t *flag);

// "avx512bw", "avx2" or "scalar": the code path hostpack_range dispatches to on this CPU.
const char *hostpack_isa();

} // namespace tdb
