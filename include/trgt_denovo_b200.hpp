// trgt_denovo_b200.hpp -- C++17 host side over the C ABI (include/trgt_denovo_b200.h).
//
// The reference's host is Rust, which this environment cannot build; this header is the compiled-language
// mirror of the reference interface for the hot path and the callers either side of it, with the
// reference's names, argument meaning and error behaviour:
//
//   tdbh::AlignmentStatus, tdbh::WFAligner                src/aligner.rs:64-86, :174-605
//       align_end_to_end :248, score :290, cigar :404, cigar_wfa :438, cigar_score_clipped :359,
//       cigar_score :378, set_heuristic :552
//   tdbh::create_aligner_with_scoring                     src/commands/shared.rs:39-51
//   tdbh::TrgtRead, Allele, AlleleSet, Locus, Params      src/read.rs:12-29, src/allele.rs:54-67,429-432,
//                                                         src/locus.rs:18-29, src/model.rs:37-42
//   tdbh::load_alleles (incl. assign_reads_by_alignment)  src/allele.rs:146-259
//   tdbh::trio::process_alleles / duo::process_alleles    src/trio/allele.rs:185-349, src/duo/allele.rs:124-246
//       for MANY loci per call: every alignment and reduction of those loci runs in one batched GPU call
//       (tdb_trio_batch_packed / tdb_duo_batch_packed / tdb_align_batch_packed: sequences are handed over as the
//       2-bit stream they were packed into when they were loaded) instead of one FFI call per read x allele
//   tdbh::serialize_with_precision, write_tsv, write_read_ids   src/allele.rs:32-50, src/commands/shared.rs:161-195
//
// Header-only; link with -ltrgt_denovo_b200.  Errors of the C ABI surface as tdbh::Error (the reference
// panics or returns Err in the same places).  There is no CPU path behind any of this.
#ifndef TRGT_DENOVO_B200_HPP
#define TRGT_DENOVO_B200_HPP

#include "trgt_denovo_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <limits>
#include <mutex>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace tdbh {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("tdb error " + std::to_string(c) + ": " + m), code(c) {}
};

// src/aligner.rs:64-86
enum class AlignmentStatus : int {
    StatusAlgCompleted = TDB_STATUS_ALG_COMPLETED,
    StatusAlgPartial = TDB_STATUS_ALG_PARTIAL,
    StatusMaxStepsReached = TDB_STATUS_MAX_STEPS_REACHED,
    StatusOOM = TDB_STATUS_OOM,
    StatusUnattainable = TDB_STATUS_UNATTAINABLE,
};

// src/aligner.rs:32-41 (the variants reachable on this path)
struct Heuristic {
    enum Kind { None = TDB_HEURISTIC_NONE, WFadaptive = TDB_HEURISTIC_WFADAPTIVE } kind = WFadaptive;
    int min_wavefront_length = 10, max_distance_threshold = 50, steps_between_cutoffs = 1;
};

// One tdb_ctx: bound to one GPU, single owner -- the role of the thread-local WFAligner of
// src/commands/trio.rs:113-116.  Move-only, frees in the destructor like Drop (src/aligner.rs:237-245).
class WFAligner {
  public:
    explicit WFAligner(int device_id = 0, int clip_len = 50) {
        tdb_config cfg;
        tdb_default_config(&cfg);
        cfg.device_id = device_id, cfg.clip_len = clip_len;
        const int r = tdb_create(&cfg, &ctx_);
        if (r != TDB_OK) throw Error(r, tdb_last_error(nullptr));
    }
    WFAligner(const WFAligner &) = delete;
    WFAligner &operator=(const WFAligner &) = delete;
    WFAligner(WFAligner &&o) noexcept : ctx_(o.ctx_) { o.ctx_ = nullptr; }
    ~WFAligner() { tdb_destroy(ctx_); }

    tdb_ctx *ctx() { return ctx_; }
    void check(int r) const {
        if (r != TDB_OK) throw Error(r, tdb_last_error(ctx_));
    }

    AlignmentStatus align_end_to_end(const std::string &pattern, const std::string &text) {
        const int r = tdb_align_end_to_end(ctx_, (const uint8_t *)pattern.data(), (int32_t)pattern.size(),
                                           (const uint8_t *)text.data(), (int32_t)text.size());
        switch (r) {
        case TDB_STATUS_ALG_COMPLETED: case TDB_STATUS_ALG_PARTIAL: case TDB_STATUS_MAX_STEPS_REACHED:
        case TDB_STATUS_OOM: case TDB_STATUS_UNATTAINABLE: return (AlignmentStatus)r;
        default: throw Error(r, tdb_last_error(ctx_)); // src/aligner.rs:83: unknown status -> panic
        }
    }
    int score() const { return tdb_score(ctx_); }
    std::string cigar_wfa() const {
        int32_t n = 0;
        const uint8_t *p = tdb_cigar_ops(ctx_, &n);
        return std::string((const char *)p, (size_t)n);
    }
    std::string cigar(std::optional<size_t> flank_len = std::nullopt) const { // src/aligner.rs:404-436
        const std::string ops = cigar_wfa();
        const long off = (long)flank_len.value_or(0), begin = off, end = (long)ops.size() - off;
        std::string out;
        if (begin >= end) return out;
        char last = ops[(size_t)begin];
        size_t run = 1;
        for (long i = begin + 1; i < end; ++i) {
            if (ops[(size_t)i] == last) ++run;
            else out += std::to_string(run) + last, last = ops[(size_t)i], run = 1;
        }
        return out + std::to_string(run) + last;
    }
    int cigar_score_clipped(size_t flank_len) const { return tdb_cigar_score_clipped(ctx_, (int32_t)flank_len); }
    int cigar_score() const { return cigar_score_clipped(0); }
    void set_heuristic(const Heuristic &h) {
        const tdb_heuristic hh = {(int)h.kind, h.min_wavefront_length, h.max_distance_threshold, h.steps_between_cutoffs};
        check(tdb_set_heuristic(ctx_, &hh));
    }
    void set_clip_len(size_t clip_len) { check(tdb_set_clip_len(ctx_, (int32_t)clip_len)); }

  private:
    tdb_ctx *ctx_ = nullptr;
};

// src/commands/shared.rs:39-51
inline WFAligner create_aligner_with_scoring(int device_id = 0) { return WFAligner(device_id); }

// ---- data carriers ------------------------------------------------------------------------------------
// A sequence packed AT THE SOURCE (north_star: "locus.rs and read.rs pack the reads and alleles ... into
// 2-bit/byte-packed device batches"): the reference turns BAM's 4-bit bases into ASCII at src/read.rs:57
// (`record.seq().as_bytes()`); a loader that builds this instead never materialises ASCII, holds a quarter of the
// bytes per read, and the batch call does no per-base CPU work (tdb_*_batch_packed).  Codes are (ASCII >> 1) & 3
// (A 0, C 1, T 2, G 3), base i at word i / 16, bits 2 (i % 16), unused high bits zero.  WFA2 compares raw bytes
// ('N' == 'N', case matters), so a sequence with any byte outside {A,C,G,T} also keeps its raw ASCII.
struct PackedBases {
    std::vector<uint32_t> w;
    uint32_t len = 0;
    std::string raw; // non-empty only when a byte outside ACGT occurs (then raw.size() == len)

    bool empty() const { return len == 0; }
    static PackedBases from_ascii(const char *s, size_t n) {
        PackedBases p;
        p.len = (uint32_t)n;
        p.w.assign((n + 15) / 16, 0u);
        bool odd = false;
        for (size_t i = 0; i < n; ++i) {
            const unsigned char c = (unsigned char)s[i];
            odd |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            p.w[i >> 4] |= (uint32_t)((c >> 1) & 3u) << (2 * (i & 15));
        }
        if (odd) p.raw.assign(s, n);
        return p;
    }
    static PackedBases from_ascii(const std::string &s) { return from_ascii(s.data(), s.size()); }
    // BAM SEQ field: two bases per byte, high nibble first, codes "=ACMGRSVTWYHKDBN" (SAM spec 4.2.3).  A C G T are
    // nibbles 1 2 4 8 -> 2-bit codes 0 1 3 2; eight bytes make one word, ASCII is never written.
    static PackedBases from_bam(const uint8_t *seq4, size_t l_seq) {
        struct Tab {
            uint8_t two[256]; // 4 bits: code(high nibble) | code(low nibble) << 2
            uint8_t bad[256]; // bit 0: high nibble outside ACGT, bit 1: low nibble
            Tab() {
                auto code = [](int nib, bool &ok) { ok = true; switch (nib) { case 1: return 0; case 2: return 1; case 4: return 3; case 8: return 2; default: ok = false; return 0; } };
                for (int b = 0; b < 256; ++b) {
                    bool okh, okl;
                    const int h = code(b >> 4, okh), l = code(b & 15, okl);
                    two[b] = (uint8_t)(h | l << 2), bad[b] = (uint8_t)((okh ? 0 : 1) | (okl ? 0 : 2));
                }
            }
        };
        static const Tab T;
        PackedBases p;
        p.len = (uint32_t)l_seq;
        p.w.assign((l_seq + 15) / 16, 0u);
        const size_t nb = l_seq / 2; // whole bytes
        unsigned odd = 0;
        size_t i = 0;
        for (; i + 8 <= nb; i += 8) { // 16 bases -> one word
            uint32_t x = 0;
            for (int k = 0; k < 8; ++k) x |= (uint32_t)T.two[seq4[i + k]] << (4 * k), odd |= T.bad[seq4[i + k]];
            p.w[i >> 3] = x;
        }
        for (; i < nb; ++i) p.w[i >> 3] |= (uint32_t)T.two[seq4[i]] << (4 * (i & 7)), odd |= T.bad[seq4[i]];
        if (l_seq & 1) p.w[nb >> 3] |= (uint32_t)(T.two[seq4[nb]] & 3u) << (4 * (nb & 7)), odd |= T.bad[seq4[nb]] & 1u;
        if (odd) { // rare: decode to ASCII as the reference does and keep both forms
            static const char code[] = "=ACMGRSVTWYHKDBN";
            std::string a(l_seq, '\0');
            for (size_t k = 0; k < l_seq; ++k) a[k] = code[(seq4[k >> 1] >> ((~k & 1) << 2)) & 15];
            return from_ascii(a);
        }
        return p;
    }
    std::string ascii() const { // `record.seq().as_bytes()` of src/read.rs:57
        if (!raw.empty()) return raw;
        std::string a(len, '\0');
        for (uint32_t i = 0; i < len; ++i) a[i] = "ACTG"[(w[i >> 4] >> (2 * (i & 15))) & 3u];
        return a;
    }
};

struct TrgtRead { // src/read.rs:12-29 (the fields this path uses)
    std::string name, bases;  // bases: ASCII as in the reference; left empty by loaders that fill `packed` instead
    PackedBases packed;       // the read packed at the source (authoritative when non-empty)
    std::optional<int> classification; // AL tag
    std::optional<int> haplotype;      // HP tag
    size_t n_bases() const { return packed.empty() ? bases.size() : packed.len; }
    std::string bases_ascii() const { return packed.empty() ? bases : packed.ascii(); }
};
struct Locus { // src/locus.rs:18-29
    std::string chrom;
    long start = 0, end = 0;
    std::vector<std::string> motifs;
    std::string id;
};
struct Allele { // src/allele.rs:54-67
    std::vector<std::pair<TrgtRead, int>> read_aligns;
    std::string seq, motif_count, allele_length;
    size_t genotype = 0, index = 0;
};
struct AlleleSet { // src/allele.rs:429-432 (+ the sample's karyotype, src/handles.rs:22-62)
    std::vector<Allele> alleles;
    size_t hp_counts[3] = {0, 0, 0};
    std::string karyotype = "XX";
};
struct SampleInput { // what get_vcf_data + get_reads return for one sample at one locus (src/allele.rs:274-426)
    std::vector<std::string> allele_seqs, motif_counts, allele_lengths;
    std::vector<size_t> genotype_indices;
    std::vector<TrgtRead> reads;
    std::string karyotype = "XX";
};
struct QuickMode { // src/model.rs:26-35
    enum Kind { AL, MC } kind = AL;
    std::optional<double> tolerance;
    bool is_zero() const { return !tolerance.has_value(); }
};
struct Params { // src/model.rs:37-42
    size_t clip_len = 50;
    double parent_quantile = 1.0;
    bool partition_by_alignment = false;
    std::optional<QuickMode> quick_mode;
};

inline const char *allele_origin_str(int code) { // src/model.rs:128-138
    static const char *s[] = {"M:1", "M:2", "F:1", "F:2", "?", "."};
    return s[code];
}
inline const char *denovo_status_str(int code) { // src/model.rs:57-83
    static const char *s[] = {"X", "Y:+", "Y:-", "Y:=", "Y:?", "."};
    return s[code];
}

// src/allele.rs:32-50
inline std::string serialize_with_precision(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[512];
    snprintf(buf, sizeof(buf), "%.1f", v);
    std::string s = buf;
    if (s.size() < 2 || s.compare(s.size() - 2, 2, ".0") != 0) {
        snprintf(buf, sizeof(buf), "%.3f", v);
        s = buf;
        while (!s.empty() && s.back() == '0') s.pop_back();
        if (!s.empty() && s.back() == '.') s.push_back('0');
    }
    return s;
}

namespace detail {

// -DTDBH_TIMING: per-phase wall times of process_alleles on stderr
struct PhaseTimer {
#ifdef TDBH_TIMING
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "  phase %-22s %9.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
#else
    void lap(const char *) {}
#endif
};

// Sequences + pair list in the reference's visiting order (src/denovo.rs:29-37,48-52), in the PACKED form of
// tdb_packed_seqs: one 2-bit stream, sequences back to back (dense layout), exception list for sequences with
// bytes outside ACGT.  Reads that were packed at the source are appended with shifts only; ASCII inputs (allele
// targets = flank + allele + flank, src/allele.rs:406-426, or reads of ASCII loaders) are packed as they are added.
struct Batch {
    std::vector<uint32_t> words; // (n_bases + 15) / 16 words while building; finish() adds the ABI's two padding words
    uint64_t n_bases = 0;
    std::vector<uint32_t> len, pp, pt, exc_seq;
    std::string exc_bytes;
    bool finished = false;

    void append_bits(const uint32_t *src, size_t n) { // n bases, unused high bits of the last source word zero
        if (!n) return;
        const unsigned sh = (unsigned)(n_bases & 15) * 2;
        const size_t nw = (n + 15) / 16;
        if (sh == 0) {
            words.insert(words.end(), src, src + nw);
        } else {
            for (size_t i = 0; i < nw; ++i) {
                words.back() |= src[i] << sh;
                words.push_back(src[i] >> (32 - sh));
            }
        }
        n_bases += n;
        words.resize((size_t)((n_bases + 15) / 16));
    }
    uint32_t seq(const PackedBases &p) {
        append_bits(p.w.data(), p.len);
        if (!p.raw.empty()) exc_seq.push_back((uint32_t)len.size()), exc_bytes += p.raw;
        len.push_back(p.len);
        return (uint32_t)len.size() - 1;
    }
    uint32_t seq(const std::string &s) { // ASCII in: packed straight onto the stream
        bool odd = false;
        words.resize((size_t)((n_bases + s.size() + 15) / 16), 0u);
        uint64_t x = n_bases;
        for (const char ch : s) {
            const unsigned char c = (unsigned char)ch;
            odd |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            words[(size_t)(x >> 4)] |= (uint32_t)((c >> 1) & 3u) << (2 * (x & 15));
            ++x;
        }
        n_bases = x;
        if (odd) exc_seq.push_back((uint32_t)len.size()), exc_bytes += s;
        len.push_back((uint32_t)s.size());
        return (uint32_t)len.size() - 1;
    }
    uint32_t seq(const TrgtRead &r) { return r.packed.empty() ? seq(r.bases) : seq(r.packed); }
    uint32_t pair(uint32_t p, uint32_t t) {
        pp.push_back(p), pt.push_back(t);
        return (uint32_t)pp.size() - 1;
    }
    void append(const Batch &o) { // o's sequences after this batch's (pairs are rebased by the caller)
        const uint32_t base = (uint32_t)len.size();
        append_bits(o.words.data(), (size_t)o.n_bases);
        len.insert(len.end(), o.len.begin(), o.len.end());
        for (uint32_t e : o.exc_seq) exc_seq.push_back(e + base);
        exc_bytes += o.exc_bytes;
    }
    tdb_packed_seqs view() { // the ABI struct over this batch (valid until the batch changes)
        if (!finished) words.resize((size_t)tdb_packed_words((size_t)n_bases), 0u), finished = true;
        tdb_packed_seqs v;
        v.packed = words.data(), v.n_bases = (size_t)n_bases, v.seq_off = nullptr, v.seq_len = len.data(), v.n_seqs = len.size();
        v.exc_seq = exc_seq.empty() ? nullptr : exc_seq.data();
        v.exc_bytes = exc_seq.empty() ? nullptr : (const uint8_t *)exc_bytes.data(), v.n_exc = exc_seq.size();
        v.seq_len16 = nullptr;
        return v;
    }
};

// Host threads of the per-locus phases (the reference runs loci on a rayon pool, src/commands/shared.rs:136-146):
// TDBH_THREADS, default the hardware concurrency (at most 64).
inline unsigned host_threads() {
    static const unsigned n = [] {
        const char *e = getenv("TDBH_THREADS");
        const unsigned v = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency();
        return std::max(1u, std::min(v ? v : 1u, 64u));
    }();
    return n;
}
inline size_t shard_count(size_t n, size_t min_per_shard) {
    return std::max<size_t>(1, std::min<size_t>(host_threads(), n / std::max<size_t>(1, min_per_shard)));
}
// fn(shard, lo, hi) over [0, n) cut into shard_count(n, min_per_shard) contiguous ranges, one thread each; the first
// exception is rethrown on the calling thread.
template <class F>
inline void parallel_ranges(size_t n, size_t min_per_shard, F fn) {
    const size_t T = shard_count(n, min_per_shard);
    if (T <= 1) {
        fn((size_t)0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    for (size_t t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            try {
                fn(t, n * t / T, n * (t + 1) / T);
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!err) err = std::current_exception();
            }
        });
    for (auto &x : th) x.join();
    if (err) std::rethrow_exception(err);
}
// Batches assembled shard by shard (same cut as parallel_ranges) -> one batch: sequence and pair indices of shard t are
// rebased by what the shards before it hold.
inline Batch merge_shards(std::vector<Batch> &sb, std::vector<tdb_trio_locus> &tl, std::vector<std::pair<uint32_t, uint32_t>> &own) {
    if (sb.size() == 1) return std::move(sb[0]);
    Batch b;
    size_t bases = 0, seqs = 0, pairs = 0;
    for (const Batch &x : sb) bases += (size_t)x.n_bases, seqs += x.len.size(), pairs += x.pp.size();
    b.words.reserve(bases / 16 + 4), b.len.reserve(seqs), b.pp.resize(pairs), b.pt.resize(pairs);
    std::vector<uint32_t> seq_base(sb.size()), pair_base(sb.size());
    size_t at_seq = 0, at_pair = 0;
    for (size_t t = 0; t < sb.size(); ++t) {
        seq_base[t] = (uint32_t)at_seq, pair_base[t] = (uint32_t)at_pair;
        at_seq += sb[t].len.size(), at_pair += sb[t].pp.size();
        b.append(sb[t]);
    }
    const size_t n = tl.size(), T = sb.size();
    parallel_ranges(T, 1, [&](size_t, size_t t0, size_t t1) {
        for (size_t t = t0; t < t1; ++t) {
            for (size_t k = 0; k < sb[t].pp.size(); ++k)
                b.pp[pair_base[t] + k] = sb[t].pp[k] + seq_base[t], b.pt[pair_base[t] + k] = sb[t].pt[k] + seq_base[t];
            for (size_t li = n * t / T; li < n * (t + 1) / T; ++li)
                for (int c = 0; c < tl[li].n_child && c < 2; ++c) {
                    for (int j = 0; j < 6; ++j) tl[li].list_off[c][j] += pair_base[t];
                    own[2 * li + c].first += pair_base[t], own[2 * li + c].second += pair_base[t];
                }
            sb[t] = Batch();
        }
    });
    return b;
}

inline int ploidy(const std::string &kary, const std::string &chrom) { // src/handles.rs:40-62
    const bool xy = chrom == "X" || chrom == "chrX" || chrom == "Y" || chrom == "chrY";
    if (kary == "XX") return (chrom == "Y" || chrom == "chrY") ? 0 : 2;
    return xy ? 1 : 2;
}
inline std::string naive_dropout(const AlleleSet &a, const Locus &l) { // src/allele.rs:447-469
    const size_t hp1 = a.hp_counts[0], hp2 = a.hp_counts[1], un = a.hp_counts[2];
    if (hp1 + hp2 + un < 2) return "FD";
    if (un < 2 && ploidy(a.karyotype, l.chrom) != 1 && (hp1 < 2 || hp2 < 2)) return "HD";
    return "N";
}
inline const std::string &field(const Allele &a, const QuickMode &q) {
    return q.kind == QuickMode::AL ? a.allele_length : a.motif_count;
}
inline double fdiv(double a, double b) { return a / b; } // IEEE: 0/0 = NaN, x/0 = inf, like Rust f64
inline double rmin(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : std::min(a, b)); } // f64::min
inline double rmax(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : std::max(a, b)); }
template <class T>
inline std::string join(const std::vector<T> &v) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + std::to_string(v[i]);
    return s;
}
inline std::string join_str(const AlleleSet &a, std::string Allele::*m) {
    std::string s;
    for (size_t i = 0; i < a.alleles.size(); ++i) s += (i ? "," : "") + a.alleles[i].*m;
    return s;
}
inline std::vector<size_t> per_allele_reads(const AlleleSet &a) { // src/math.rs:18-20
    std::vector<size_t> v;
    for (const Allele &x : a.alleles) v.push_back(x.read_aligns.size());
    return v;
}
inline size_t total_reads(const AlleleSet &a) {
    size_t t = 0;
    for (const Allele &x : a.alleles) t += x.read_aligns.size();
    return t;
}

} // namespace detail

// load_alleles (src/allele.rs:146-191) for many (locus, sample) inputs at once.  With partition_by_alignment
// every read is aligned against every allele of its own sample in ONE batched GPU call; only the
// order-dependent tie flip (src/allele.rs:215,238-243) runs on the host.  nullopt in -> nullopt out (Err).
// The inputs are taken by value: pass std::move(...) for genome-sized batches and the reads are moved, not copied,
// into the allele sets.
inline std::vector<std::optional<AlleleSet>> load_alleles(WFAligner &aligner, std::vector<std::optional<SampleInput>> samples,
                                                          const Params &params) {
    std::vector<std::optional<AlleleSet>> out;
    std::vector<int32_t> scores, status;
    std::vector<size_t> pair_base;
    if (params.partition_by_alignment) {
        detail::Batch b;
        for (const auto &s : samples) {
            pair_base.push_back(b.pp.size());
            if (!s) continue;
            std::vector<uint32_t> ids;
            for (const std::string &a : s->allele_seqs) ids.push_back(b.seq(a));
            for (const TrgtRead &r : s->reads) {
                const uint32_t ri = b.seq(r);
                for (uint32_t ai : ids) b.pair(ri, ai);
            }
        }
        if (!b.pp.empty()) {
            scores.resize(b.pp.size()), status.resize(b.pp.size());
            aligner.set_clip_len(params.clip_len);
            const tdb_packed_seqs ps = b.view();
            aligner.check(tdb_align_batch_packed(aligner.ctx(), &ps, b.pp.data(), b.pt.data(), b.pp.size(), scores.data(), status.data()));
        }
    }
    out.resize(samples.size());
    detail::parallel_ranges(samples.size(), 128, [&](size_t, size_t s_lo, size_t s_hi) {
    for (size_t si = s_lo; si < s_hi; ++si) {
        if (!samples[si]) continue;
        SampleInput &s = *samples[si];
        const size_t n_all = s.allele_seqs.size();
        AlleleSet set;
        set.karyotype = s.karyotype;
        set.alleles.resize(n_all);
        for (const TrgtRead &r : s.reads) { // calculate_hp_counts, src/allele.rs:117-128
            if (r.haplotype && *r.haplotype == 1) ++set.hp_counts[0];
            else if (r.haplotype && *r.haplotype == 2) ++set.hp_counts[1];
            else if (!r.haplotype) ++set.hp_counts[2];
        }
        if (params.partition_by_alignment) {
            size_t index_flip = 0, j = pair_base[si];
            for (TrgtRead &r : s.reads) {
                std::vector<std::pair<size_t, int>> max_aligns;
                bool have = false;
                int max_score = 0;
                for (size_t i = 0; i < n_all; ++i, ++j) {
                    if (status[j] != TDB_STATUS_ALG_COMPLETED) continue;
                    const int sc = scores[j];
                    if (!have || sc > max_score) have = true, max_score = sc, max_aligns.assign(1, {i, sc});
                    else if (sc == max_score) max_aligns.push_back({i, sc});
                }
                if (!max_aligns.empty()) {
                    if (max_aligns.size() > 1) index_flip = (index_flip + 1) % max_aligns.size();
                    const auto [ai, al] = max_aligns[index_flip % max_aligns.size()];
                    set.alleles[ai].read_aligns.emplace_back(std::move(r), al);
                }
            }
        } else {
            for (TrgtRead &r : s.reads) {
                if (!r.classification) throw Error(TDB_E_INVALID, "read without AL tag (src/allele.rs:255 panics)");
                set.alleles.at((size_t)*r.classification).read_aligns.emplace_back(std::move(r), 0);
            }
        }
        for (size_t i = 0; i < n_all; ++i) {
            Allele &a = set.alleles[i];
            a.seq = std::move(s.allele_seqs[i]), a.motif_count = std::move(s.motif_counts[i]), a.allele_length = std::move(s.allele_lengths[i]);
            a.genotype = s.genotype_indices[i], a.index = i;
        }
        out[si] = std::move(set);
    }
    });
    return out;
}

// One TSV row; the union of the trio (src/trio/allele.rs:26-69) and duo (src/duo/allele.rs:89-122) schemas.
struct AlleleResult {
    std::string chrom, motifs, trid;
    long start = 0, end = 0;
    size_t genotype = 0, denovo_coverage = 0, allele_coverage = 0, child_coverage = 0, index = 0;
    double allele_ratio = 0.0, child_ratio = 0.0, father_dropout_prob = 0.0, mother_dropout_prob = 0.0;
    float mean_diff_father = 0.f, mean_diff_mother = 0.f;
    std::string allele_origin = ".", denovo_status = ".";
    std::string per_allele_reads_father = ".", per_allele_reads_mother = ".", per_allele_reads_child = ".";
    std::string father_dropout = ".", mother_dropout = ".", child_dropout = ".";
    std::string father_MC = ".", mother_MC = ".", child_MC = ".", father_AL = ".", mother_AL = ".", child_AL = ".";
    std::string father_overlap_coverage = ".", mother_overlap_coverage = ".";
    std::optional<std::vector<std::string>> read_ids;
};

namespace detail {

inline void sample_columns(const std::optional<AlleleSet> &a, const Locus &l, std::string &dropout, std::string &reads,
                           std::string &mc, std::string &al, double *prob) {
    if (!a) return;
    dropout = naive_dropout(*a, l);
    reads = join(per_allele_reads(*a));
    mc = join_str(*a, &Allele::motif_count), al = join_str(*a, &Allele::allele_length);
    if (prob) *prob = std::pow(0.5, (double)total_reads(*a)); // math::get_dropout_prob, src/math.rs:44-49
}

// adds the score lists of one child / sample-A allele against the other sample(s); returns the own-read range
inline std::pair<uint32_t, uint32_t> add_lists(Batch &b, tdb_trio_locus &loc, int c, const Allele &ca,
                                               const std::vector<std::vector<uint32_t>> &m_reads,
                                               const std::vector<std::vector<uint32_t>> &f_reads) {
    const uint32_t tgt = b.seq(ca.seq);
    const std::vector<uint32_t> empty;
    const std::vector<uint32_t> *lists[4] = {&m_reads[0], m_reads.size() > 1 ? &m_reads[1] : &empty,
                                             !f_reads.empty() ? &f_reads[0] : &empty, f_reads.size() > 1 ? &f_reads[1] : &empty};
    for (int j = 0; j < 4; ++j) {
        loc.list_off[c][j] = (uint32_t)b.pp.size();
        for (uint32_t ri : *lists[j]) b.pair(ri, tgt);
    }
    loc.list_off[c][4] = (uint32_t)b.pp.size();
    for (const auto &ra : ca.read_aligns) b.pair(b.seq(ra.first), tgt);
    loc.list_off[c][5] = (uint32_t)b.pp.size();
    return {loc.list_off[c][4], loc.list_off[c][5]};
}

inline AlleleResult template_row(const Locus &l) {
    AlleleResult r;
    r.chrom = l.chrom, r.start = l.start, r.end = l.end, r.trid = l.id;
    for (size_t i = 0; i < l.motifs.size(); ++i) r.motifs += (i ? "," : "") + l.motifs[i];
    return r;
}

} // namespace detail

namespace trio {

inline bool check_field_equivalence(const AlleleSet &f, const AlleleSet &m, const AlleleSet &c, const QuickMode &q) { // src/trio/allele.rs:71-99
    std::vector<std::string> fv, mv, cv;
    for (const Allele &a : f.alleles) fv.push_back(detail::field(a, q));
    for (const Allele &a : m.alleles) mv.push_back(detail::field(a, q));
    for (const Allele &a : c.alleles) cv.push_back(detail::field(a, q));
    std::sort(cv.begin(), cv.end());
    auto has = [](const std::vector<std::string> &v, const std::string &x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    if (cv.size() == 1) return has(fv, cv[0]) || has(mv, cv[0]);
    for (const auto &x : fv)
        for (const auto &y : mv) {
            std::vector<std::string> pr = {x, y};
            std::sort(pr.begin(), pr.end());
            if (pr == cv) return true;
        }
    return false;
}

inline bool check_field_similarity(const AlleleSet &f, const AlleleSet &m, const AlleleSet &c, const QuickMode &q) { // src/trio/allele.rs:101-183
    using detail::fdiv;
    const double tol = *q.tolerance, inf = std::numeric_limits<double>::infinity();
    std::vector<int> fv, mv, cv;
    for (const Allele &a : f.alleles) fv.push_back(std::stoi(detail::field(a, q)));
    for (const Allele &a : m.alleles) mv.push_back(std::stoi(detail::field(a, q)));
    for (const Allele &a : c.alleles) cv.push_back(std::stoi(detail::field(a, q)));
    std::sort(cv.begin(), cv.end());
    double rel;
    if (cv.size() == 1) {
        double bf = inf, bm = inf;
        for (int x : fv) bf = detail::rmin(bf, (double)std::abs(x - cv[0]));
        for (int x : mv) bm = detail::rmin(bm, (double)std::abs(x - cv[0]));
        rel = detail::rmin(fdiv(bf, (double)*std::max_element(fv.begin(), fv.end())),
                           fdiv(bm, (double)*std::max_element(mv.begin(), mv.end())));
    } else {
        bool have = false;
        std::pair<int, int> best;
        double best_sum = inf;
        for (int x : fv)
            for (int y : mv) {
                const std::pair<int, int> pr = x < y ? std::make_pair(x, y) : std::make_pair(y, x);
                const double d = (double)std::abs(pr.first - cv[0]) + (double)std::abs(pr.second - cv[1]);
                if (d < best_sum) best_sum = d, best = pr, have = true;
            }
        if (!have) rel = inf;
        else {
            const double d0 = (best.first == 0 && cv[0] == 0) ? 0.0 : fdiv((double)std::abs(best.first - cv[0]), (double)std::max(best.first, cv[0]));
            const double d1 = (best.second == 0 && cv[1] == 0) ? 0.0 : fdiv((double)std::abs(best.second - cv[1]), (double)std::max(best.second, cv[1]));
            rel = detail::rmax(d0, d1);
        }
    }
    return rel <= tol;
}

// process_alleles (src/trio/allele.rs:185-349) for many loci: -> rows per locus, in locus order.
inline std::vector<std::vector<AlleleResult>> process_alleles(WFAligner &aligner, const std::vector<Locus> &loci,
                                                              std::vector<std::optional<SampleInput>> father,
                                                              std::vector<std::optional<SampleInput>> mother,
                                                              std::vector<std::optional<SampleInput>> child,
                                                              const Params &params) {
    const size_t n = loci.size();
    std::vector<std::optional<SampleInput>> all(std::move(father));
    all.insert(all.end(), std::make_move_iterator(mother.begin()), std::make_move_iterator(mother.end()));
    all.insert(all.end(), std::make_move_iterator(child.begin()), std::make_move_iterator(child.end()));
    detail::PhaseTimer pt;
    auto sets = load_alleles(aligner, std::move(all), params);
    pt.lap("load_alleles");
    std::vector<AlleleResult> templ(n);
    std::vector<uint8_t> reaches(n, 0);
    detail::parallel_ranges(n, 256, [&](size_t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            AlleleResult r = detail::template_row(loci[i]);
            const auto &f = sets[i], &m = sets[n + i], &c = sets[2 * n + i];
            detail::sample_columns(f, loci[i], r.father_dropout, r.per_allele_reads_father, r.father_MC, r.father_AL, &r.father_dropout_prob);
            detail::sample_columns(m, loci[i], r.mother_dropout, r.per_allele_reads_mother, r.mother_MC, r.mother_AL, &r.mother_dropout_prob);
            detail::sample_columns(c, loci[i], r.child_dropout, r.per_allele_reads_child, r.child_MC, r.child_AL, nullptr);
            templ[i] = std::move(r);
            if (!f || !m || !c) continue; // missing genotyping: one template row (src/trio/allele.rs:270-286)
            if (params.quick_mode) {
                const bool skip = params.quick_mode->is_zero() ? check_field_equivalence(*f, *m, *c, *params.quick_mode)
                                                               : check_field_similarity(*f, *m, *c, *params.quick_mode);
                if (skip) continue;
            }
            reaches[i] = 1;
        }
    });
    std::vector<size_t> todo;
    for (size_t i = 0; i < n; ++i)
        if (reaches[i]) todo.push_back(i);
    pt.lap("template rows");
    // one batch for every locus that reaches assess_denovo (src/trio/denovo.rs:83-186), assembled shard by shard
    std::vector<tdb_trio_locus> tl(todo.size());
    std::vector<std::pair<uint32_t, uint32_t>> own(2 * todo.size());
    std::vector<detail::Batch> shards(detail::shard_count(todo.size(), 64));
    detail::parallel_ranges(todo.size(), 64, [&](size_t t, size_t lo, size_t hi) {
        detail::Batch &b = shards[t];
        size_t bytes = 0, seqs = 0;
        for (size_t li = lo; li < hi; ++li)
            for (const auto *set : {&*sets[todo[li]], &*sets[n + todo[li]], &*sets[2 * n + todo[li]]})
                for (const Allele &a : set->alleles) {
                    bytes += a.seq.size(), seqs += 1 + a.read_aligns.size();
                    for (const auto &ra : a.read_aligns) bytes += ra.first.n_bases();
                }
        b.words.reserve(bytes / 16 + 4), b.len.reserve(seqs), b.pp.reserve(2 * seqs), b.pt.reserve(2 * seqs);
        for (size_t li = lo; li < hi; ++li) {
            const size_t i = todo[li];
            const AlleleSet &f = *sets[i], &m = *sets[n + i], &c = *sets[2 * n + i];
            tdb_trio_locus &loc = tl[li];
            memset(&loc, 0, sizeof(loc));
            loc.n_child = (int)c.alleles.size(), loc.n_mother = (int)m.alleles.size(), loc.n_father = (int)f.alleles.size();
            std::vector<std::vector<uint32_t>> mr(m.alleles.size()), fr(f.alleles.size());
            for (size_t a = 0; a < m.alleles.size(); ++a)
                for (const auto &ra : m.alleles[a].read_aligns) mr[a].push_back(b.seq(ra.first));
            for (size_t a = 0; a < f.alleles.size(); ++a)
                for (const auto &ra : f.alleles[a].read_aligns) fr[a].push_back(b.seq(ra.first));
            for (int a = 0; a < 2; ++a) {
                loc.child_len[a] = a < loc.n_child ? (int)c.alleles[a].seq.size() : 0;
                loc.mother_len[a] = a < loc.n_mother ? (int)m.alleles[a].seq.size() : 0;
                loc.father_len[a] = a < loc.n_father ? (int)f.alleles[a].seq.size() : 0;
            }
            for (int ci = 0; ci < loc.n_child && ci < 2; ++ci) own[2 * li + ci] = detail::add_lists(b, loc, ci, c.alleles[ci], mr, fr);
        }
    });
    detail::Batch b = detail::merge_shards(shards, tl, own);
    pt.lap("batch assembly");
    std::vector<tdb_allele_out> out(2 * todo.size());
    std::vector<uint8_t> mask(b.pp.size() + 1);
    if (!todo.empty()) {
        aligner.set_clip_len(params.clip_len);
        const tdb_packed_seqs ps = b.view();
        aligner.check(tdb_trio_batch_packed(aligner.ctx(), &ps, b.pp.data(), b.pt.data(), b.pp.size(), tl.data(), tl.size(),
                                            params.parent_quantile, out.data(), nullptr, mask.data()));
    }
    pt.lap("tdb_trio_batch_packed");
    std::vector<std::vector<AlleleResult>> result(n);
    detail::parallel_ranges(n, 256, [&](size_t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) result[i] = {templ[i]};
    });
    detail::parallel_ranges(todo.size(), 64, [&](size_t, size_t lo, size_t hi) {
        for (size_t li = lo; li < hi; ++li) {
            const size_t i = todo[li];
            const AlleleSet &f = *sets[i], &m = *sets[n + i], &c = *sets[2 * n + i];
            const size_t child_cov = detail::total_reads(c);
            std::vector<AlleleResult> rows;
            for (size_t ci = 0; ci < c.alleles.size() && ci < 2; ++ci) {
                const tdb_allele_out &o = out[2 * li + ci];
                if (o.error) throw Error(TDB_E_INVALID, "TRID=" + loci[i].id + ": no parental alignment scores (the reference panics here, src/trio/denovo.rs:333-334)");
                const Allele &ca = c.alleles[ci];
                AlleleResult r = templ[i];
                r.genotype = ca.genotype, r.denovo_coverage = (size_t)o.denovo_coverage, r.allele_coverage = ca.read_aligns.size();
                r.allele_ratio = r.allele_coverage == 0 ? 0.0 : (double)r.denovo_coverage / (double)r.allele_coverage;
                r.child_coverage = child_cov, r.child_ratio = detail::fdiv((double)r.denovo_coverage, (double)child_cov);
                r.mean_diff_father = o.mean_diff_father, r.mean_diff_mother = o.mean_diff_mother;
                r.allele_origin = allele_origin_str(o.allele_origin), r.denovo_status = denovo_status_str(o.denovo_status);
                r.index = ca.index;
                r.father_overlap_coverage = detail::join(std::vector<int>(o.father_overlap, o.father_overlap + f.alleles.size()));
                r.mother_overlap_coverage = detail::join(std::vector<int>(o.mother_overlap, o.mother_overlap + m.alleles.size()));
                std::vector<std::string> ids;
                for (uint32_t k = own[2 * li + ci].first; k < own[2 * li + ci].second; ++k)
                    if (mask[k]) ids.push_back(ca.read_aligns[k - own[2 * li + ci].first].first.name);
                if (!ids.empty()) r.read_ids = std::move(ids);
                rows.push_back(std::move(r));
            }
            if (!rows.empty()) result[i] = std::move(rows);
        }
    });
    pt.lap("rows");
    detail::parallel_ranges(sets.size(), 256, [&](size_t, size_t lo, size_t hi) { // millions of small strings to free
        for (size_t i = lo; i < hi; ++i) sets[i].reset();
    });
    pt.lap("teardown");
    return result;
}

inline const std::vector<std::string> &columns() {
    static const std::vector<std::string> c = {
        "chrom", "start", "end", "motifs", "trid", "genotype", "denovo_coverage", "allele_coverage", "allele_ratio",
        "child_coverage", "child_ratio", "mean_diff_father", "mean_diff_mother", "father_dropout_prob", "mother_dropout_prob",
        "allele_origin", "denovo_status", "per_allele_reads_father", "per_allele_reads_mother", "per_allele_reads_child",
        "father_dropout", "mother_dropout", "child_dropout", "index", "father_MC", "mother_MC", "child_MC", "father_AL",
        "mother_AL", "child_AL", "father_overlap_coverage", "mother_overlap_coverage"};
    return c;
}
inline std::string row_tsv(const AlleleResult &r) {
    const std::string t = "\t";
    return r.chrom + t + std::to_string(r.start) + t + std::to_string(r.end) + t + r.motifs + t + r.trid + t +
           std::to_string(r.genotype) + t + std::to_string(r.denovo_coverage) + t + std::to_string(r.allele_coverage) + t +
           serialize_with_precision(r.allele_ratio) + t + std::to_string(r.child_coverage) + t +
           serialize_with_precision(r.child_ratio) + t + serialize_with_precision((double)r.mean_diff_father) + t +
           serialize_with_precision((double)r.mean_diff_mother) + t + serialize_with_precision(r.father_dropout_prob) + t +
           serialize_with_precision(r.mother_dropout_prob) + t + r.allele_origin + t + r.denovo_status + t +
           r.per_allele_reads_father + t + r.per_allele_reads_mother + t + r.per_allele_reads_child + t + r.father_dropout + t +
           r.mother_dropout + t + r.child_dropout + t + std::to_string(r.index) + t + r.father_MC + t + r.mother_MC + t +
           r.child_MC + t + r.father_AL + t + r.mother_AL + t + r.child_AL + t + r.father_overlap_coverage + t +
           r.mother_overlap_coverage;
}

} // namespace trio

namespace duo {

inline bool check_allele_field_equivalence(const AlleleSet &a, const AlleleSet &b, const QuickMode &q) { // src/duo/allele.rs:33-49
    std::set<std::string> sa, sb;
    for (const Allele &x : a.alleles) sa.insert(detail::field(x, q));
    for (const Allele &x : b.alleles) sb.insert(detail::field(x, q));
    return sa == sb;
}
inline bool check_field_similarity(const AlleleSet &a, const AlleleSet &b, const QuickMode &q) { // src/duo/allele.rs:51-87
    std::vector<int> av, bv;
    for (const Allele &x : a.alleles) av.push_back(std::stoi(detail::field(x, q)));
    for (const Allele &x : b.alleles) bv.push_back(std::stoi(detail::field(x, q)));
    if (av.size() != bv.size()) return false;
    std::sort(av.begin(), av.end()), std::sort(bv.begin(), bv.end());
    for (size_t i = 0; i < av.size(); ++i)
        if (!(detail::fdiv((double)std::abs(av[i] - bv[i]), (double)std::max(av[i], bv[i])) <= *q.tolerance)) return false;
    return true;
}

// duo process_alleles (src/duo/allele.rs:124-246) for many loci.  Row fields reuse AlleleResult: sample A in the
// child_* / per_allele_reads_child slots, sample B in the mother_* slots, mean_diff_b in mean_diff_mother.
inline std::vector<std::vector<AlleleResult>> process_alleles(WFAligner &aligner, const std::vector<Locus> &loci,
                                                              std::vector<std::optional<SampleInput>> sample_a,
                                                              std::vector<std::optional<SampleInput>> sample_b,
                                                              const Params &params) {
    const size_t n = loci.size();
    std::vector<std::optional<SampleInput>> all(std::move(sample_a));
    all.insert(all.end(), std::make_move_iterator(sample_b.begin()), std::make_move_iterator(sample_b.end()));
    auto sets = load_alleles(aligner, std::move(all), params);
    std::vector<AlleleResult> templ(n);
    std::vector<uint8_t> reaches(n, 0);
    detail::parallel_ranges(n, 256, [&](size_t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            AlleleResult r = detail::template_row(loci[i]);
            const auto &a = sets[i], &bb = sets[n + i];
            detail::sample_columns(a, loci[i], r.child_dropout, r.per_allele_reads_child, r.child_MC, r.child_AL, nullptr);
            detail::sample_columns(bb, loci[i], r.mother_dropout, r.per_allele_reads_mother, r.mother_MC, r.mother_AL, nullptr);
            templ[i] = std::move(r);
            if (!a || !bb) continue;
            if (params.quick_mode) {
                const bool skip = params.quick_mode->is_zero() ? check_allele_field_equivalence(*a, *bb, *params.quick_mode)
                                                               : check_field_similarity(*a, *bb, *params.quick_mode);
                if (skip) continue;
            }
            reaches[i] = 1;
        }
    });
    std::vector<size_t> todo;
    for (size_t i = 0; i < n; ++i)
        if (reaches[i]) todo.push_back(i);
    std::vector<tdb_trio_locus> tl(todo.size());
    std::vector<std::pair<uint32_t, uint32_t>> own(2 * todo.size());
    std::vector<detail::Batch> shards(detail::shard_count(todo.size(), 64));
    detail::parallel_ranges(todo.size(), 64, [&](size_t t, size_t lo, size_t hi) {
        detail::Batch &b = shards[t];
        for (size_t li = lo; li < hi; ++li) {
            const size_t i = todo[li];
            const AlleleSet &a = *sets[i], &sb = *sets[n + i];
            tdb_trio_locus &loc = tl[li];
            memset(&loc, 0, sizeof(loc));
            loc.n_child = (int)a.alleles.size(), loc.n_mother = (int)sb.alleles.size(), loc.n_father = 0;
            std::vector<std::vector<uint32_t>> br(sb.alleles.size()), none;
            for (size_t x = 0; x < sb.alleles.size(); ++x)
                for (const auto &ra : sb.alleles[x].read_aligns) br[x].push_back(b.seq(ra.first));
            for (int x = 0; x < 2; ++x) {
                loc.child_len[x] = x < loc.n_child ? (int)a.alleles[x].seq.size() : 0;
                loc.mother_len[x] = x < loc.n_mother ? (int)sb.alleles[x].seq.size() : 0;
            }
            for (int ci = 0; ci < loc.n_child && ci < 2; ++ci) own[2 * li + ci] = detail::add_lists(b, loc, ci, a.alleles[ci], br, none);
        }
    });
    detail::Batch b = detail::merge_shards(shards, tl, own);
    std::vector<tdb_allele_out> out(2 * todo.size());
    std::vector<uint8_t> mask(b.pp.size() + 1);
    if (!todo.empty()) {
        aligner.set_clip_len(params.clip_len);
        const tdb_packed_seqs ps = b.view();
        aligner.check(tdb_duo_batch_packed(aligner.ctx(), &ps, b.pp.data(), b.pt.data(), b.pp.size(), tl.data(), tl.size(),
                                           params.parent_quantile, out.data(), nullptr, mask.data()));
    }
    std::vector<std::vector<AlleleResult>> result(n);
    detail::parallel_ranges(n, 256, [&](size_t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) result[i] = {templ[i]};
    });
    detail::parallel_ranges(todo.size(), 64, [&](size_t, size_t lo, size_t hi) {
        for (size_t li = lo; li < hi; ++li) {
            const size_t i = todo[li];
            const AlleleSet &a = *sets[i], &sb = *sets[n + i];
            const size_t a_cov = detail::total_reads(a);
            std::vector<AlleleResult> rows;
            for (size_t ci = 0; ci < a.alleles.size() && ci < 2; ++ci) {
                const tdb_allele_out &o = out[2 * li + ci];
                if (o.error) throw Error(TDB_E_INVALID, "TRID=" + loci[i].id + ": no sample-B alignment scores (the reference panics here, src/duo/denovo.rs:107)");
                const Allele &ca = a.alleles[ci];
                AlleleResult r = templ[i];
                r.genotype = ca.genotype, r.denovo_coverage = (size_t)o.denovo_coverage, r.allele_coverage = ca.read_aligns.size();
                r.allele_ratio = r.allele_coverage == 0 ? 0.0 : (double)r.denovo_coverage / (double)r.allele_coverage;
                r.child_coverage = a_cov, r.child_ratio = detail::fdiv((double)r.denovo_coverage, (double)a_cov);
                r.mean_diff_mother = o.mean_diff_mother;
                r.denovo_status = denovo_status_str(o.denovo_status), r.index = ca.index;
                r.mother_overlap_coverage = detail::join(std::vector<int>(o.mother_overlap, o.mother_overlap + sb.alleles.size()));
                std::vector<std::string> ids;
                for (uint32_t k = own[2 * li + ci].first; k < own[2 * li + ci].second; ++k)
                    if (mask[k]) ids.push_back(ca.read_aligns[k - own[2 * li + ci].first].first.name);
                if (!ids.empty()) r.read_ids = std::move(ids);
                rows.push_back(std::move(r));
            }
            if (!rows.empty()) result[i] = std::move(rows);
        }
    });
    detail::parallel_ranges(sets.size(), 256, [&](size_t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) sets[i].reset();
    });
    return result;
}

inline const std::vector<std::string> &columns() {
    static const std::vector<std::string> c = {"chrom", "start", "end", "motifs", "trid", "genotype", "denovo_coverage",
                                               "allele_coverage", "allele_ratio", "a_coverage", "a_ratio", "mean_diff_b",
                                               "denovo_status", "per_allele_reads_a", "per_allele_reads_b", "a_dropout",
                                               "b_dropout", "index", "a_MC", "b_MC", "a_AL", "b_AL", "b_overlap_coverage"};
    return c;
}
inline std::string row_tsv(const AlleleResult &r) {
    const std::string t = "\t";
    return r.chrom + t + std::to_string(r.start) + t + std::to_string(r.end) + t + r.motifs + t + r.trid + t +
           std::to_string(r.genotype) + t + std::to_string(r.denovo_coverage) + t + std::to_string(r.allele_coverage) + t +
           serialize_with_precision(r.allele_ratio) + t + std::to_string(r.child_coverage) + t +
           serialize_with_precision(r.child_ratio) + t + serialize_with_precision((double)r.mean_diff_mother) + t +
           r.denovo_status + t + r.per_allele_reads_child + t + r.per_allele_reads_mother + t + r.child_dropout + t +
           r.mother_dropout + t + std::to_string(r.index) + t + r.child_MC + t + r.mother_MC + t + r.child_AL + t +
           r.mother_AL + t + r.mother_overlap_coverage;
}

} // namespace duo

// ---- several GPUs of one box ------------------------------------------------------------------------------
// The reference treats the one locus stream as the unit of parallelism (src/commands/shared.rs:107-146: rayon
// par_bridge at :138, one thread-local aligner per worker, :136-145).  Here: one WFAligner (= one tdb_ctx, one GPU)
// and one host thread per GPU; the locus list is cut into contiguous ranges balanced on the alignment work proxy
// sum over pairs of (n + m), every range runs the whole per-locus pipeline on its own GPU, and the
// rows are concatenated in locus (BED) order -- the order the reference writes with -@1.  No collective: loci are
// independent.  n_gpus may exceed the visible devices (contexts then share devices round-robin), which keeps the
// sharding and merging testable on a one-GPU box.
class MultiGpu {
  public:
    explicit MultiGpu(int n_gpus = 0) {
        const int visible = tdb_device_count();
        if (visible <= 0) throw Error(TDB_E_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
        const int n = n_gpus > 0 ? n_gpus : visible;
        for (int i = 0; i < n; ++i) aligners_.emplace_back(i % visible);
        const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
        for (auto &a : aligners_) a.check(tdb_set_host_threads(a.ctx(), (int32_t)std::max(1u, hc / (unsigned)n)));
    }
    size_t size() const { return aligners_.size(); }

    // contiguous ranges whose weight sums are as equal as the prefix sums allow: parts + 1 boundaries
    static std::vector<size_t> balanced_bounds(const std::vector<uint64_t> &w, size_t parts) {
        const size_t n = w.size();
        std::vector<uint64_t> pre(n + 1, 0);
        for (size_t i = 0; i < n; ++i) pre[i + 1] = pre[i] + w[i];
        std::vector<size_t> b(1, 0);
        for (size_t r = 1; r < parts; ++r) {
            const long double target = (long double)pre[n] * r / parts;
            size_t i = (size_t)(std::lower_bound(pre.begin(), pre.end(), (uint64_t)target) - pre.begin());
            if (i > 0 && std::fabs((long double)pre[i - 1] - target) <= std::fabs((long double)pre[std::min(i, n)] - target)) --i;
            b.push_back(std::min(std::max(i, b.back()), n));
        }
        b.push_back(n);
        return b;
    }
    // sum over the locus' pairs of (read length + target length), from the inputs (before any alignment)
    static uint64_t locus_weight(const std::optional<SampleInput> &parent1, const std::optional<SampleInput> &parent2,
                                 const std::optional<SampleInput> &child) {
        if (!child || !parent1) return 1;
        uint64_t reads = 0, bases = 0, own_reads = child->reads.size(), own_bases = 0, tgt = 0;
        for (const auto *p : {&parent1, &parent2})
            if (*p)
                for (const TrgtRead &r : (**p).reads) ++reads, bases += r.n_bases();
        for (const TrgtRead &r : child->reads) own_bases += r.n_bases();
        for (const std::string &a : child->allele_seqs) tgt += a.size();
        const uint64_t n_t = std::max<uint64_t>(1, child->allele_seqs.size());
        return 1 + n_t * bases + reads * tgt + own_bases + own_reads * (tgt / n_t);
    }

    std::vector<std::vector<AlleleResult>> trio_process_alleles(const std::vector<Locus> &loci,
                                                                std::vector<std::optional<SampleInput>> father,
                                                                std::vector<std::optional<SampleInput>> mother,
                                                                std::vector<std::optional<SampleInput>> child, const Params &params) {
        std::vector<uint64_t> w(loci.size());
        for (size_t i = 0; i < loci.size(); ++i) w[i] = locus_weight(mother[i], father[i], child[i]);
        return run(loci, w, [&](WFAligner &a, size_t lo, size_t hi) {
            return trio::process_alleles(a, slice(loci, lo, hi), take(father, lo, hi), take(mother, lo, hi), take(child, lo, hi), params);
        });
    }
    std::vector<std::vector<AlleleResult>> duo_process_alleles(const std::vector<Locus> &loci,
                                                               std::vector<std::optional<SampleInput>> sample_a,
                                                               std::vector<std::optional<SampleInput>> sample_b, const Params &params) {
        std::vector<uint64_t> w(loci.size());
        for (size_t i = 0; i < loci.size(); ++i) w[i] = locus_weight(sample_b[i], std::nullopt, sample_a[i]);
        return run(loci, w, [&](WFAligner &a, size_t lo, size_t hi) {
            return duo::process_alleles(a, slice(loci, lo, hi), take(sample_a, lo, hi), take(sample_b, lo, hi), params);
        });
    }
    const std::vector<size_t> &last_bounds() const { return bounds_; }

  private:
    template <class T>
    static std::vector<T> slice(const std::vector<T> &v, size_t lo, size_t hi) { return std::vector<T>(v.begin() + (long)lo, v.begin() + (long)hi); }
    template <class T>
    static std::vector<T> take(std::vector<T> &v, size_t lo, size_t hi) { // moves the range out (the reads are large)
        return std::vector<T>(std::make_move_iterator(v.begin() + (long)lo), std::make_move_iterator(v.begin() + (long)hi));
    }
    template <class F>
    std::vector<std::vector<AlleleResult>> run(const std::vector<Locus> &loci, const std::vector<uint64_t> &w, F fn) {
        const size_t G = aligners_.size();
        bounds_ = balanced_bounds(w, G);
        std::vector<std::vector<std::vector<AlleleResult>>> part(G);
        std::vector<std::thread> th;
        std::exception_ptr err;
        std::mutex mu;
        for (size_t g = 0; g < G; ++g)
            th.emplace_back([&, g]() {
                try {
                    if (bounds_[g + 1] > bounds_[g]) part[g] = fn(aligners_[g], bounds_[g], bounds_[g + 1]);
                } catch (...) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!err) err = std::current_exception();
                }
            });
        for (auto &t : th) t.join();
        if (err) std::rethrow_exception(err);
        std::vector<std::vector<AlleleResult>> rows; // BED order: the ranges are contiguous and ascending
        rows.reserve(loci.size());
        for (auto &p : part)
            for (auto &r : p) rows.push_back(std::move(r));
        return rows;
    }
    std::vector<WFAligner> aligners_;
    std::vector<size_t> bounds_;
};

// The writer thread of src/commands/shared.rs:161-195 with -@1 ordering: TSV text and the read-id side file.
template <class RowFn>
inline std::string write_tsv(const std::vector<std::vector<AlleleResult>> &rows, const std::vector<std::string> &cols, RowFn row_tsv) {
    std::string s;
    for (size_t i = 0; i < cols.size(); ++i) s += (i ? "\t" : "") + cols[i];
    s += "\n";
    for (const auto &locus_rows : rows)
        for (const AlleleResult &r : locus_rows) s += row_tsv(r) + "\n";
    return s;
}
inline std::string write_read_ids(const std::vector<std::vector<AlleleResult>> &rows) {
    std::string s;
    for (const auto &locus_rows : rows)
        for (size_t i = 0; i < locus_rows.size(); ++i) {
            const AlleleResult &r = locus_rows[i];
            if (!r.read_ids || r.read_ids->empty()) continue;
            s += ">TRID=" + r.trid + "\tALLELE=" + std::to_string(i) + "\tN=" + std::to_string(r.read_ids->size()) + "\n";
            for (const std::string &id : *r.read_ids) s += id + "\n";
        }
    return s;
}

} // namespace tdbh
#endif // TRGT_DENOVO_B200_HPP
