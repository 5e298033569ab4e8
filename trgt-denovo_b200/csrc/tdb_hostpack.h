// tdb_hostpack.h -- host side of the packing role north_star gives locus.rs / read.rs: the raw ASCII
// blob -> the 2-bit stream the kernels read, produced by the caller's CPU threads so that a host
// batch crosses PCIe at a quarter of its ASCII size.  Pure data re-encoding: no alignment
// arithmetic happens on the host.  Bit layout identical to k_pack (tdb_kernels.cuh): the byte at blob
// position x sits in word x/16 at bits 2*(x%16), code = (byte >> 1) & 3, independent of sequence
// boundaries (a sequence is the window [seq_off, seq_off+len) of the stream).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace tdb {

// Packs blob bytes [x0, x1) into words [x0/16, ceil(x1/16)) of `packed`; x0 must be a multiple of 64.
// A byte outside {A,C,G,T} sets flag[i] = 1 on the sequence i that owns it (binary search in seq_off,
// which must be ascending); flag must be zeroed by the caller.  Thread-safe across disjoint ranges.
// seq_off may be null for a dense layout (sequences back to back from byte 0): make_off(make_arg) is then
// called -- only if such a byte turns up -- and must return the offsets (thread-safe, same pointer every time).
void hostpack_stream(const uint8_t *blob, uint64_t x0, uint64_t x1, uint32_t *packed, const uint64_t *seq_off,
                     const uint32_t *seq_len, uint32_t n_seqs, uint8_t *flag,
                     const uint64_t *(*make_off)(void *) = nullptr, void *make_arg = nullptr);

// Number of sequences of [s0, s1) violating the layout contract (inside the blob, index order,
// non-overlapping).
size_t hostpack_check_layout(const uint64_t *seq_off, const uint32_t *seq_len, uint32_t s0, uint32_t s1, uint32_t n_seqs,
                             size_t blob_bytes);

// "avx512bw", "avx2" or "scalar": the code path hostpack_stream dispatches to on this CPU.
const char *hostpack_isa();

} // namespace tdb
