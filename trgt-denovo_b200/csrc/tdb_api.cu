// tdb_api.cu -- C ABI (include/trgt_denovo_b200.h) over the sm_100a kernels.  No CPU compute path:
// every entry point that does arithmetic launches CUDA kernels or fails.
#include "tdb_wfa_quad.cuh"
#include "tdb_wfa_frame.cuh"
#include "tdb_wfa_thread.cuh"
#include "tdb_hostpack.h"

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace tdb;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void release() {
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
    }
};

struct TierCfg {
    int wlog2, s_cap, ops_cap;
    unsigned prov_cap;
    int workers;
};

// Device control block (one per ctx).  "call" fields live for a whole batch call, "chunk" fields are
// zeroed before every chunk.
struct Ctrl {
    unsigned err;      // ERR_* bits
    unsigned gcount;   // pairs queued for the global-scratch tier (listB)
    unsigned ccount;   // pairs the first global tier could not finish (listC)
    unsigned wc_g[2];  // work counters of the two global tiers
    unsigned n_closed; // pairs resolved in closed form (k_dedup_window)
    unsigned n_failed; // pairs the last tier could not finish
    unsigned dcount;   // pairs the long tier could not finish (listD)
    unsigned wc_long, pad0[3]; // work counter of the long tier
    unsigned long long exec[6][4]; // executed cells, ext16, columns, pairs finished; per tier ([4]: thread-per-pair tier, [5]: long tier)
    unsigned long long algo[5];    // algorithmic cells, ext16, columns (all pairs); pairs with rep == self; failed pairs
    // ---- per chunk (one block per pipeline slot, zeroed before every chunk) ----
    struct Chunk {
        unsigned wc_quad, wc_warp, acount, wc_mini;
        unsigned mcount, pad1[3]; // mcount: pairs the thread-per-pair tier handed on
        unsigned hist[N_HIST];
        unsigned base[N_HIST], cursor[N_HIST]; // work-list buckets: start and fill level
        unsigned seg[8];
        unsigned hist2[EST_BUCKETS], base2[EST_BUCKETS], cursor2[EST_BUCKETS]; // thread tier re-sorted by the penalty estimate
        unsigned seg2[4];
    } ck[2];
};
constexpr int N_SLOT = 2;

} // namespace

struct tdb_ctx {
    tdb_config cfg;
    int dev = 0;
    int sm_count = 0;
    std::string err;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[16] = {};
    std::vector<cudaEvent_t> up_ev, cp_ev; // per chunk: data prepared (preparation stream) / copies done (copy stream)
    cudaStream_t prep_stream = nullptr; // kernels on freshly uploaded data (see run_align)
    // grow-only device buffers
    DevBuf blob, seq_off, seq_len, packed, seq_flag, seq_hash, pair_p, pair_t, score, status, penalty;
    DevBuf rep, work, list_b, list_c, list_d, scan_tmp, mask_hits, len16;
    std::vector<unsigned> h_hits; // host copy of the set mask positions
    // Chunks alternate between two pipeline slots (own stream + own scratch) so that the tail of one chunk's
    // alignment kernels overlaps the next chunk's duplicate elimination and alignment.  (3 and 4 slots, and a
    // warp tier deferred to a call-wide list, were measured and made no difference: profiles/r01/NOTES.md.)
    struct Slot {
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf skey_in, sval_out, list_a, list_m, list_t;
    } slot[2];
    DevBuf ctrl, loci, aout, mask, scores_in, cigar, cigar_off, cigar_len;
    DevBuf scratch[2];
    DevBuf lockq_scratch[2]; // the same for tier 0
    DevBuf lock_scratch[2]; // provenance / step records of the four-pairs-per-warp tier 1 (per pipeline slot)
    int duo4_ctas_per_sm = 0, long_ctas_per_sm = 0, quad8_ctas_per_sm = 0;
    bool quad8 = false; // TDB_QUAD8=1: tier 0 with eight pairs per warp (Quad8Cfg)
    int long_sort_per_warp = 8; // the long tier's list is sorted heaviest first beyond this many pairs per warp (TDB_LONG_SORT_PER_WARP)
    bool disable_long = false; // TDB_DISABLE_LONG: no long tier (k_align_lock<LongCfg>) in front of the global-scratch tiers
    bool duo4 = true; // TDB_DUO_G16=1: tier 1 with two pairs per warp (16 lanes each), everything in shared memory
    int t0_ctas_per_sm = 0, quad_ctas_per_sm = 0, duo_ctas_per_sm = 0, mini_ctas_per_sm = 0;
    bool disable_mini = false;
    int global_warps_per_sm = 16; // workers of the first global-scratch tier (TDB_GLOBAL_WARPS_PER_SM)
    size_t sparse_mask_min_pairs = (size_t)8 << 20; // host calls at least this large fetch the mask as a list of set positions
    bool disable_est = false; // TDB_DISABLE_EST_SORT: thread tier in |n-m| order, no penalty estimate
    bool old_lock = true; // TDB_FRAME_LOCK=1: k_align_frame instead of k_align_lock (measured 2-10 % slower, profiles/r02/NOTES.md)
    int fquad_ctas_per_sm = 0, fduo_ctas_per_sm = 0;
    bool disable_quad = false, disable_duo = false, disable_dedup = false, disable_closed_form = false, disable_rle = false;
    size_t chunk_pairs = 2u << 20;
    size_t hostpack_min_bytes = 8u << 20; // smaller host batches are packed on the device
    int host_threads = 0;                 // 0: all hardware threads but one
    void *h_packed = nullptr, *h_flag = nullptr, *h_runs = nullptr; // pinned staging of the host packer / planner
    size_t h_packed_cap = 0, h_flag_cap = 0, h_runs_cap = 0;
    DevBuf runs;
    // host-pack mode: blocks the packer threads have not reached can ship as ASCII and be packed on the GPU.  Automatic:
    // on when the call has at most 8 packer threads -- several ranks sharing one host -- where it measured +10-15 % e2e
    // (look-ahead rule driven by the copy stream, see run_align); off with more threads, where it measured nothing
    // (profiles/r01/NOTES.md).  TDB_RAW_EVERY=n forces every n-th chunk instead (n < 2: off).
    bool steal_forced = false, allow_steal = false, steal_auto_always = false; // TDB_STEAL_AUTO=1: the automatic rule whatever the thread count
    int steal_div = 2;
    bool count_work = false;
    int test_max_score = 0; // TDB_TEST_MAX_SCORE: caps the score range of the global tiers (tests force a last-tier failure)
    tdb_counters counters;
    // single-pair shim state
    std::vector<uint8_t> last_cigar;
    int last_penalty = INT32_MIN, last_status = TDB_STATUS_UNATTAINABLE;
};

namespace {

int fail(tdb_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? TDB_E_OOM : TDB_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
    } while (0)

int ensure(tdb_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return TDB_OK;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    want = (want + 255) & ~(size_t)255;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream)); // the old allocation may still be in use by queued work
        CK(cudaFree(b.p));
    }
    b.p = nullptr, b.cap = 0;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return TDB_OK;
}
#define ENSURE(buf, bytes)                          \
    do {                                            \
        int r_ = ensure(ctx, (buf), (bytes));       \
        if (r_ != TDB_OK) return r_;                \
    } while (0)

// Where the batch lives.  Device pointers always; host pointers only for the host-buffer entry points
// (then the device arrays are ctx buffers that run_align fills chunk by chunk).
struct DevBatch {
    const uint8_t *blob;
    size_t blob_bytes;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    size_t n_seqs;
    const uint32_t *pp, *pt;
    size_t n_pairs;
    int32_t *score, *status, *penalty;
    uint8_t *cigar_buf;
    const uint64_t *cigar_off;
    uint32_t *cigar_len;
    // host sources (nullptr: inputs already resident)
    const uint8_t *h_blob;
    const uint64_t *h_seq_off;
    const uint32_t *h_seq_len;
    const uint16_t *h_seq_len16; // packed input: lengths as uint16 instead (h_seq_len == nullptr)
    const uint32_t *h_pp, *h_pt;
    uint8_t *h_zero; // optional host array of n_pairs bytes the calling thread zeroes chunk by chunk (result mask)
    // packed host input (tdb_*_batch_packed): the 2-bit stream as tdb_host_pack lays it out, plus the sequences that
    // hold a byte outside ACGT (ascending indices, raw bytes back to back)
    // an upload the alignment stage does not need (the locus descriptors of the fused calls): queued on the copy stream
    // behind the last chunk, so that it does not delay the first one
    void *late_dst;
    const void *late_src;
    size_t late_bytes;
    const uint32_t *h_packed;
    const uint32_t *h_exc_seq;
    const uint8_t *h_exc_bytes;
    size_t n_exc;
};

int check_sizes(tdb_ctx *ctx, size_t blob_bytes, size_t n_seqs, size_t n_pairs) {
    if (n_seqs >= 0x7fffffffull || n_pairs >= 0x7fffffffull)
        return fail(ctx, TDB_E_UNSUPPORTED, "batch too large: n_seqs=%zu n_pairs=%zu (max 2^31-2 per call)", n_seqs,
                    n_pairs);
    if (blob_bytes / 16 + 2 * n_seqs + 2 >= 0xffffffffull)
        return fail(ctx, TDB_E_UNSUPPORTED, "sequence blob too large for 32-bit word offsets: %zu bytes", blob_bytes);
    return TDB_OK;
}

constexpr int SCAN_PARTS = 4; // a chunk's pair list is scanned by this many independent tasks
struct Chunk {
    size_t p0, p1;       // pair range
    uint32_t s_lo, s_hi; // sequences referenced: [s_lo, s_hi)
    size_t run0 = 0;     // this chunk's slice of the run buffer (SCAN_PARTS sub-slices of run_cap runs)
    uint32_t run_cap = 0, n_runs = 0;
    uint32_t part_runs[SCAN_PARTS] = {}, part_lo[SCAN_PARTS] = {}, part_hi[SCAN_PARTS] = {};
    bool part_ok[SCAN_PARTS] = {};
    bool rle = false;    // pair list shipped as runs
};

// Worker threads that 2-bit pack a host blob block by block, in byte order, while the calling thread
// queues uploads and kernels for the chunks whose bytes are already packed.
class HostPacker {
  public:
    static constexpr uint64_t BLOCK = 256u << 10; // bytes of ASCII per work item (a multiple of 64)
    HostPacker() = default;
    HostPacker(const HostPacker &) = delete;
    ~HostPacker() { join(); }
    // Work order on the shared threads: `n_pre` barrier tasks first (they zero the flag array the packer writes
    // into), then, chunk by chunk, the scan parts of chunk k (pair list -> sequence range + run-length encoding)
    // followed by that chunk's share of the pack blocks -- so the first chunk is ready after a K-th of the work
    // and planning never holds up packing.
    void start(const uint8_t *blob, uint64_t blob_bytes, const uint64_t *off, const uint32_t *len, uint32_t n_seqs,
               uint32_t *packed, uint8_t *flag, unsigned n_threads, size_t n_pre, std::function<void(size_t)> pre_fn,
               size_t n_scan_chunks, int parts, std::function<void(size_t, int)> scan_fn, int raw_every_block,
               const uint64_t *(*make_off)(void *) = nullptr, void *make_arg = nullptr) {
        n_blocks_ = packed ? (size_t)((blob_bytes + BLOCK - 1) / BLOCK) : 0;
        done_ = std::vector<std::atomic<uint8_t>>(n_blocks_);
        // raw_every_block (tests only): every n-th block is left to the GPU from the start, whatever the timing
        for (size_t i = 0; i < n_blocks_; ++i)
            done_[i].store(raw_every_block > 1 && i % (size_t)raw_every_block == (size_t)raw_every_block - 1 ? RAW : PENDING,
                           std::memory_order_relaxed);
        scan_left_ = std::vector<std::atomic<int>>(n_scan_chunks);
        for (auto &x : scan_left_) x.store(parts, std::memory_order_relaxed);
        tasks_.clear();
        const size_t groups = std::max<size_t>(1, n_scan_chunks);
        for (size_t k = 0; k < groups; ++k) {
            if (k < n_scan_chunks)
                for (int pt = 0; pt < parts; ++pt) tasks_.push_back({1, (uint32_t)k, (uint32_t)pt});
            for (size_t i = n_blocks_ * k / groups; i < n_blocks_ * (k + 1) / groups; ++i)
                if (done_[i].load(std::memory_order_relaxed) == PENDING) tasks_.push_back({2, (uint32_t)i, 0});
        }
        next_.store(0), next_pre_.store(0), pre_left_.store((long)n_pre);
        for (unsigned t = 0; t < n_threads; ++t)
            threads_.emplace_back([=]() {
                for (;;) {
                    const size_t i = next_pre_.fetch_add(1, std::memory_order_relaxed);
                    if (i >= n_pre) break;
                    pre_fn(i);
                    pre_left_.fetch_sub(1, std::memory_order_release);
                }
                while (pre_left_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
                for (;;) {
                    const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
                    if (i >= tasks_.size()) return;
                    const Task tk = tasks_[i];
                    if (tk.kind == 1) {
                        scan_fn(tk.a, (int)tk.b);
                        scan_left_[tk.a].fetch_sub(1, std::memory_order_release);
                    } else {
                        uint8_t expect = PENDING;
                        if (!done_[tk.a].compare_exchange_strong(expect, BUSY, std::memory_order_acq_rel)) continue; // stolen
                        const uint64_t x0 = (uint64_t)tk.a * BLOCK, x1 = std::min(blob_bytes, x0 + BLOCK);
                        hostpack_stream(blob, x0, x1, packed, off, len, n_seqs, flag, make_off, make_arg);
                        done_[tk.a].store(PACKED, std::memory_order_release);
                        last_pack_ns_.store(std::chrono::steady_clock::now().time_since_epoch().count(), std::memory_order_relaxed);
                    }
                }
            });
    }
    enum : uint8_t { PENDING = 0, PACKED = 1, RAW = 2, BUSY = 3 };
    size_t n_blocks() const { return n_blocks_; }
    uint8_t peek(size_t i) const { return done_[i].load(std::memory_order_acquire); }
    // The calling thread takes a block no worker has started: it ships as ASCII and the GPU packs it.
    bool try_steal(size_t i) {
        uint8_t st = PENDING;
        return done_[i].compare_exchange_strong(st, RAW, std::memory_order_acq_rel);
    }
    // State of block i for the calling thread: RAW (left to the GPU) or PACKED (waits for the worker that is on it).
    uint8_t acquire(size_t i) const {
        for (;;) {
            const uint8_t st = done_[i].load(std::memory_order_acquire);
            if (st == PACKED || st == RAW) return st;
            std::this_thread::yield();
        }
    }
    void wait_scan(size_t k) const {
        if (k >= scan_left_.size()) return;
        while (scan_left_[k].load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }
    void join() {
        next_.store(tasks_.size()); // unclaimed tasks are abandoned (error paths)
        for (auto &t : threads_) t.join();
        threads_.clear();
    }

  private:
    struct Task {
        uint8_t kind; // 1 scan part (a = chunk, b = part), 2 pack block a
        uint32_t a, b;
    };
    size_t n_blocks_ = 0;
    std::vector<Task> tasks_;
    std::vector<std::atomic<uint8_t>> done_;
    std::vector<std::atomic<int>> scan_left_;
    std::atomic<size_t> next_{0}, next_pre_{0};
    std::atomic<long> pre_left_{0};
    std::atomic<long long> last_pack_ns_{0};
    std::vector<std::thread> threads_;

  public:
    long long last_pack_ns() const { return last_pack_ns_.load(); }
};

int ensure_pinned(tdb_ctx *ctx, void *&p, size_t &cap, size_t bytes) {
    if (bytes <= cap) return TDB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr, cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
        return fail(ctx, TDB_E_OOM, "pinned staging allocation of %zu bytes failed", want);
    }
    cap = want;
    return TDB_OK;
}

// Dense layout (seq_off == NULL at the ABI): sequence i starts where sequence i-1 ends, sequence 0 at byte 0.
// The device derives the offsets with a prefix sum of the lengths; the host needs them only to find the owner
// of a byte outside ACGT, and builds the table the first time such a byte turns up.
struct DenseOffsets {
    const uint32_t *len = nullptr;
    size_t n = 0;
    std::once_flag once;
    std::vector<uint64_t> off;
    const uint64_t *get() {
        std::call_once(once, [this]() {
            off.resize(n + 1);
            uint64_t s = 0;
            for (size_t i = 0; i < n; ++i) off[i] = s, s += len[i];
            off[n] = s;
        });
        return off.data();
    }
    static const uint64_t *thunk(void *self) { return static_cast<DenseOffsets *>(self)->get(); }
};
struct LenTo64 {
    __host__ __device__ __forceinline__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};
// d_off[i] = base + sum of d_len[0 .. i) for i in [0, n)
int scan_offsets(tdb_ctx *ctx, const uint32_t *d_len, uint64_t *d_off, size_t n, uint64_t base, cudaStream_t st) {
    size_t tmp = ctx->scan_tmp.cap;
    CK(cub::DeviceScan::ExclusiveScan(ctx->scan_tmp.p, tmp, thrust::make_transform_iterator(d_len, LenTo64()), d_off,
                                      cub::Sum(), base, (int)n, st));
    return TDB_OK;
}
int scan_offsets_reserve(tdb_ctx *ctx, size_t n_max) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, thrust::make_transform_iterator((const uint32_t *)nullptr, LenTo64()),
                                      (uint64_t *)nullptr, cub::Sum(), (uint64_t)0, (int)n_max, ctx->stream));
    ENSURE(ctx->scan_tmp, tmp + 16);
    return TDB_OK;
}

// Splits a host batch into pair chunks whose sequence ranges advance with the pair index (true for
// batches assembled locus by locus), so that a chunk can be aligned while the next one is still in
// flight over PCIe.  Falls back to a single chunk when the pair list jumps around the sequence table.
// plan_prepare sizes the chunk table; scan_chunk(k) (any thread) fills chunk k; plan_finish validates.
void plan_prepare(const tdb_ctx *ctx, const DevBatch &b, std::vector<Chunk> &out) {
    out.clear();
    const size_t n_pairs = b.n_pairs;
    if (!b.h_pp || n_pairs <= ctx->chunk_pairs + ctx->chunk_pairs / 2) {
        out.push_back({0, n_pairs, 0, (uint32_t)b.n_seqs});
        return;
    }
    const size_t K = (n_pairs + ctx->chunk_pairs - 1) / ctx->chunk_pairs;
    const size_t per = (n_pairs + K - 1) / K;
    out.resize(K);
    size_t run0 = 0;
    for (size_t k = 0; k < K; ++k) {
        out[k] = {k * per, std::min(n_pairs, (k + 1) * per), 0, 0};
        out[k].run0 = run0, out[k].run_cap = (uint32_t)((out[k].p1 - out[k].p0) / (4 * SCAN_PARTS) + 16);
        run0 += (size_t)out[k].run_cap * SCAN_PARTS;
    }
}
// Part `part` of a chunk's pair list: the sequence range it touches and, when `runs` is given, its run-length
// encoding (consecutive patterns against one text).  A part whose runs do not fit a quarter of its pairs makes
// the whole chunk travel as plain arrays.
void scan_chunk_part(const DevBatch &b, Chunk &c, int part, PairRun *runs) {
    const size_t len = c.p1 - c.p0, q0 = c.p0 + len * part / SCAN_PARTS, q1 = c.p0 + len * (part + 1) / SCAN_PARTS;
    uint32_t lo = 0xffffffffu, hi = 0;
    PairRun *out = runs ? runs + c.run0 + (size_t)c.run_cap * part : nullptr;
    uint32_t n_runs = 0, next_p = 0, cur_t = 0;
    bool ok = runs != nullptr;
    for (size_t j = q0; j < q1; ++j) {
        const uint32_t a = b.h_pp[j], t = b.h_pt[j];
        lo = std::min(lo, std::min(a, t));
        hi = std::max(hi, std::max(a, t));
        if (ok && (n_runs == 0 || a != next_p || t != cur_t)) {
            if (n_runs == c.run_cap) ok = false;
            else out[n_runs++] = {(uint32_t)(j - c.p0), a, t};
        }
        next_p = a + 1, cur_t = t;
    }
    c.part_lo[part] = lo, c.part_hi[part] = hi, c.part_runs[part] = n_runs, c.part_ok[part] = ok;
}
void finish_chunk_scan(const DevBatch &b, Chunk &c) {
    uint32_t lo = 0xffffffffu, hi = 0;
    c.rle = true, c.n_runs = 0;
    for (int p = 0; p < SCAN_PARTS; ++p) {
        lo = std::min(lo, c.part_lo[p]), hi = std::max(hi, c.part_hi[p]);
        c.rle = c.rle && c.part_ok[p];
        c.n_runs += c.part_runs[p];
    }
    c.rle = c.rle && c.n_runs > 0;
    if (lo > hi) lo = hi = 0;
    hi = (uint32_t)std::min<size_t>((size_t)hi + 1, b.n_seqs); // out-of-range indices fail on the device
    c.s_lo = std::min(lo, hi), c.s_hi = hi;
}
float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

template <class T>
uint64_t sum_lengths(const T *__restrict__ p, size_t n) {
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    size_t i = 0;
    for (; i + 4 <= n; i += 4) s0 += p[i], s1 += p[i + 1], s2 += p[i + 2], s3 += p[i + 3];
    for (; i < n; ++i) s0 += p[i];
    return s0 + s1 + s2 + s3;
}

// lengths of sequences [a, a + n) to the device: uint32 as they are, or uint16 (widened next to the data by k_widen_len,
// which the caller launches on the preparation stream)
int upload_lengths(tdb_ctx *ctx, const DevBatch &b, uint32_t a, size_t n, cudaStream_t cs, tdb_counters &cn) {
    if (b.h_seq_len) {
        CK(cudaMemcpyAsync((uint32_t *)b.seq_len + a, b.h_seq_len + a, n * 4, cudaMemcpyHostToDevice, cs));
        cn.bytes_h2d += n * 4;
    } else {
        CK(cudaMemcpyAsync((uint16_t *)ctx->len16.p + a, b.h_seq_len16 + a, n * 2, cudaMemcpyHostToDevice, cs));
        cn.bytes_h2d += n * 2;
    }
    return TDB_OK;
}

// Pack + duplicate elimination + alignment tiers + scatter.  Leaves results in b.score/status/penalty.
int run_align(tdb_ctx *ctx, const DevBatch &b_in, bool cigar_mode) {
    DevBatch b = b_in;
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    const uint32_t n_seqs = (uint32_t)b.n_seqs, n_pairs = (uint32_t)b.n_pairs;
    tdb_counters &cn = ctx->counters;
    memset(&cn, 0, sizeof(cn));
    cn.pairs = n_pairs;
    CK(cudaEventRecord(ctx->ev[0], s));
    if (n_pairs == 0 || n_seqs == 0) {
        if (b.late_bytes) CK(cudaMemcpyAsync(b.late_dst, b.late_src, b.late_bytes, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(ctx->ev[1], s));
        CK(cudaStreamSynchronize(s));
        return TDB_OK;
    }
    if (!b.h_pp && !b.seq_off) { // device batch in the dense layout: offsets = prefix sum of the lengths
        ENSURE(ctx->seq_off, (size_t)n_seqs * 8 + 8);
        int r = scan_offsets_reserve(ctx, n_seqs);
        if (r) return r;
        r = scan_offsets(ctx, b.seq_len, (uint64_t *)ctx->seq_off.p, n_seqs, 0, s);
        if (r) return r;
        b.seq_off = (const uint64_t *)ctx->seq_off.p;
        cn.kernel_launches += 1;
    }
    const tdb_penalties &cp = ctx->cfg.penalties;
    const bool default_pen = cp.mismatch == 8 && cp.gap_opening1 == 4 && cp.gap_extension1 == 2 &&
                             cp.gap_opening2 == 24 && cp.gap_extension2 == 1;
    const bool quad_ok = default_pen && !ctx->disable_quad;
    const bool mini_ok = quad_ok && !ctx->disable_mini; // its leftovers go to the lockstep tier
    const bool dedup = !ctx->disable_dedup && !cigar_mode;

    const auto t_call = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    // ---- host batches: the caller's CPU threads 2-bit pack the blob (a quarter of the ASCII bytes cross
    // PCIe) while this thread plans and queues chunks; small batches and CIGAR calls ship the ASCII and
    // pack on the device -------------------------------------------------------------------------------
    const bool prepacked = b.h_pp && b.h_packed != nullptr; // the caller packed at load time: no per-base host work
    const bool host_pack = b.h_pp && !prepacked && !cigar_mode && b.blob_bytes >= ctx->hostpack_min_bytes;
    const bool dense = b.h_pp && !b.h_seq_off; // host batch without offsets: sequences back to back
    DenseOffsets dense_off; // declared before the packer: its threads may call into it
    dense_off.len = b.h_seq_len, dense_off.n = n_seqs; // (only the ASCII host packer asks for it)
    auto hlen = [&b](uint32_t i) -> uint32_t { return b.h_seq_len ? b.h_seq_len[i] : (uint32_t)b.h_seq_len16[i]; };
    HostPacker packer;
    std::vector<Chunk> chunks;
    plan_prepare(ctx, b, chunks);
    if (dense) {
        int r = scan_offsets_reserve(ctx, (size_t)n_seqs + 1);
        if (r) return r;
    }
    if (host_pack) {
        int r = ensure_pinned(ctx, ctx->h_packed, ctx->h_packed_cap, (b.blob_bytes / 16 + 2) * 4);
        if (r) return r;
        r = ensure_pinned(ctx, ctx->h_flag, ctx->h_flag_cap, (size_t)n_seqs);
        if (r) return r;
        cn.host_packed = 1;
    }
    if (prepacked) cn.host_packed = 2;
    const bool lazy_scan = b.h_pp && chunks.size() > 1;
    std::vector<Chunk> scan_chunks = chunks; // what the worker threads fill in; `chunks` is the calling thread's copy
    if (b.h_pp && (host_pack || lazy_scan)) {
        // one core stays free for the calling thread, which must react quickly to every finished chunk
        const unsigned hc = std::thread::hardware_concurrency();
        const unsigned hw = ctx->host_threads > 0 ? (unsigned)ctx->host_threads : std::max(1u, std::min(32u, hc > 2 ? hc - 1 : 1u));
        if (!ctx->steal_forced) ctx->allow_steal = hw <= 8 || ctx->steal_auto_always;
        const size_t n_zero = host_pack ? 16 : 0;
        PairRun *hruns = nullptr;
        if (lazy_scan && !ctx->disable_rle) { // pair lists travel run-length encoded
            const size_t total = chunks.back().run0 + (size_t)chunks.back().run_cap * SCAN_PARTS;
            int r = ensure_pinned(ctx, ctx->h_runs, ctx->h_runs_cap, total * sizeof(PairRun));
            if (r) return r;
            ENSURE(ctx->runs, total * sizeof(PairRun));
            hruns = (PairRun *)ctx->h_runs;
        }
        uint8_t *hflag = (uint8_t *)ctx->h_flag;
        packer.start(b.h_blob, b.blob_bytes, b.h_seq_off, b.h_seq_len, n_seqs, host_pack ? (uint32_t *)ctx->h_packed : nullptr,
                     hflag, hw, n_zero,
                     [&b, n_zero, hflag](size_t k) { // zero a slice of the flag array
                         const size_t per = (b.n_seqs + n_zero - 1) / n_zero, z0 = std::min(b.n_seqs, k * per);
                         memset(hflag + z0, 0, std::min(b.n_seqs, z0 + per) - z0);
                     },
                     lazy_scan ? chunks.size() : 0, SCAN_PARTS,
                     [&scan_chunks, &b, hruns](size_t k, int part) { scan_chunk_part(b, scan_chunks[k], part, hruns); },
                     getenv("TDB_TEST_RAW_BLOCKS") ? atoi(getenv("TDB_TEST_RAW_BLOCKS")) : 0,
                     dense ? &DenseOffsets::thunk : nullptr, &dense_off);
    }
    if (lazy_scan) {
        // The first chunk tells whether the pair list walks the sequence table in order; if it does not (a chunk
        // would touch a large share of all sequences) the batch runs as one chunk.
        packer.wait_scan(0);
        chunks[0] = scan_chunks[0];
        finish_chunk_scan(b, chunks[0]);
        if ((size_t)(chunks[0].s_hi - chunks[0].s_lo) > 4 * (b.n_seqs / chunks.size()) + 64) {
            chunks.clear();
            chunks.push_back({0, b.n_pairs, 0, (uint32_t)b.n_seqs});
        }
    }
    const bool scanned_lazily = lazy_scan && chunks.size() > 1;
    cn.ms_host_plan = ms_since(t_call);
    const size_t K = chunks.size();
    cn.chunks = K;
    const bool staged = K == 1; // per-stage CUDA-event timing only when the call is one chunk
    size_t max_cp = 0;
    for (const Chunk &c : chunks) max_cp = std::max(max_cp, c.p1 - c.p0);

    // ---- buffers ------------------------------------------------------------------------------------
    const size_t max_words = b.blob_bytes / 16 + 2;
    ENSURE(ctx->packed, (max_words + PACK_SLACK) * 4); // the comparison loops read a few words past a sequence
    ENSURE(ctx->seq_flag, (size_t)n_seqs);
    ENSURE(ctx->seq_hash, (size_t)n_seqs * 8);
    ENSURE(ctx->ctrl, sizeof(Ctrl));
    ENSURE(ctx->list_b, (size_t)n_pairs * 4);
    const bool want_work = ctx->count_work && !cigar_mode;
    // closed form for single-substitution / single-gap (<= 20 bases) contents: production penalties, wf-adaptive
    // or no heuristic, and no exact work counting requested (derivation at closed_form, tdb_kernels.cuh)
    const tdb_heuristic &hh = ctx->cfg.heuristic;
    ClosedFormCfg cf;
    cf.enabled = default_pen && !ctx->count_work && !ctx->disable_closed_form &&
                 (hh.kind == TDB_HEURISTIC_NONE || hh.min_wavefront_length >= 10);
    cf.clip = ctx->cfg.clip_len;
    const int n_slot = K > 1 ? N_SLOT : 1;
    if (!cigar_mode) {
        ENSURE(ctx->rep, (size_t)n_pairs * 4);
        if (want_work) ENSURE(ctx->work, (size_t)n_pairs * 12);
        for (int i = 0; i < n_slot; ++i) {
            tdb_ctx::Slot &sl = ctx->slot[i];
            ENSURE(sl.skey_in, max_cp);
            ENSURE(sl.sval_out, max_cp * 4);
            ENSURE(sl.list_a, max_cp * 4);
            if (mini_ok) ENSURE(sl.list_m, max_cp * 4);
            if (mini_ok && !ctx->disable_est) ENSURE(sl.list_t, max_cp * 4);
        }
    }
    while (b.h_pp && ctx->up_ev.size() < K) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->up_ev.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->cp_ev.push_back(e);
    }
    Ctrl *ctrl = (Ctrl *)ctx->ctrl.p;
    CK(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), s));
    CK(cudaEventRecord(ctx->ev[10], s)); // the other slot's stream and the copy stream start after the reset
    for (int i = 1; i < n_slot; ++i) CK(cudaStreamWaitEvent(ctx->slot[i].st, ctx->ev[10], 0));
    if (b.h_pp) CK(cudaStreamWaitEvent(cs, ctx->ev[10], 0));
    if (b.h_pp) CK(cudaStreamWaitEvent(ctx->prep_stream, ctx->ev[10], 0));

    AlignParams p;
    memset(&p, 0, sizeof(p));
    p.pn = {cp.mismatch, cp.gap_opening1, cp.gap_extension1, cp.gap_opening2, cp.gap_extension2};
    p.heur = {ctx->cfg.heuristic.kind, ctx->cfg.heuristic.min_wavefront_length,
              ctx->cfg.heuristic.max_distance_threshold, ctx->cfg.heuristic.steps_between_cutoffs};
    p.clip = ctx->cfg.clip_len;
    p.packed = (const uint32_t *)ctx->packed.p, p.seq_flag = (const uint8_t *)ctx->seq_flag.p;
    p.blob = b.blob, p.seq_off = b.seq_off, p.seq_len = b.seq_len, p.pair_p = b.pp, p.pair_t = b.pt;
    p.out_score = b.score, p.out_status = b.status, p.out_penalty = b.penalty;
    p.work = want_work ? (uint32_t *)ctx->work.p : nullptr;
    p.cigar_buf = b.cigar_buf, p.cigar_off = b.cigar_off, p.cigar_len = b.cigar_len;
    p.fail_count = &ctrl->n_failed;

    uint32_t seq_up = 0; // sequences [0, seq_up) are uploaded / packed
    uint64_t dense_pos = 0; // dense layout: byte offset of sequence seq_up
    size_t exc_i = 0;       // packed input: next entry of the exception list, and where its raw bytes start
    uint64_t exc_pos = 0;
    std::vector<std::pair<uint64_t, uint32_t>> exc_at; // (blob offset, length) of this chunk's exceptions
    std::vector<std::pair<uint64_t, uint64_t>> raw_ranges; // this chunk's byte ranges shipped as ASCII in host-pack mode

    // ---- chunks ---------------------------------------------------------------------------------------
    for (size_t k = 0; k < K; ++k) {
        if (scanned_lazily && k > 0) {
            const auto t_wait = std::chrono::steady_clock::now();
            packer.wait_scan(k);
            cn.ms_host_pack_wait += ms_since(t_wait);
            chunks[k] = scan_chunks[k];
            finish_chunk_scan(b, chunks[k]);
        }
        const Chunk &c = chunks[k];
        const uint32_t cpairs = (uint32_t)(c.p1 - c.p0);
        if (b.h_zero) memset(b.h_zero + c.p0, 0, c.p1 - c.p0);
        const uint32_t a = seq_up, e = std::max(seq_up, c.s_hi); // sequences first needed by this chunk: [a, e)
        seq_up = e;
        // blob bytes of the new sequences (device batches: offsets unknown on the host -> everything at once)
        uint64_t b0 = 0, b1 = b.blob_bytes;
        size_t n_len_up = 0; // lengths uploaded for this chunk (the kernels on them run on the preparation stream)
        const uint64_t off_a = dense_pos;
        if (b.h_pp && e > a) {
            uint64_t first = 0, end = 0;
            exc_at.clear();
            if (dense) {
                // bytes of the new sequences = sum of their lengths: a tight loop over plain local pointers (this runs on the
                // calling thread once per chunk, over ~a million lengths: it must not be slower than the chunk's upload),
                // cut at the listed exceptions, whose offsets it yields on the way
                const uint32_t *l32 = b.h_seq_len;
                const uint16_t *l16 = b.h_seq_len16;
                auto span_sum = [l32, l16](uint32_t i0, uint32_t i1) {
                    return l32 ? sum_lengths(l32 + i0, i1 - i0) : sum_lengths(l16 + i0, i1 - i0);
                };
                uint64_t sum = 0;
                uint32_t at = a;
                for (size_t xi = exc_i; xi < b.n_exc && b.h_exc_seq[xi] < e; ++xi) {
                    const uint32_t x = b.h_exc_seq[xi];
                    if (x < at) continue;
                    sum += span_sum(at, x);
                    exc_at.push_back({off_a + sum, hlen(x)});
                    at = x;
                }
                sum += span_sum(at, e);
                first = off_a, end = off_a + sum;
                dense_pos = end;
            } else {
                for (size_t xi = exc_i; xi < b.n_exc && b.h_exc_seq[xi] < e; ++xi)
                    if (b.h_exc_seq[xi] >= a) exc_at.push_back({b.h_seq_off[b.h_exc_seq[xi]], hlen(b.h_exc_seq[xi])});
                first = b.h_seq_off[a], end = b.h_seq_off[e - 1] + hlen(e - 1);
            }
            b0 = std::min<uint64_t>(first, b.blob_bytes) & ~(uint64_t)15;
            b1 = e == n_seqs ? b.blob_bytes : std::min<uint64_t>(end, b.blob_bytes);
            b1 = std::max(b1, b0);
        }
        if (b.h_pp) { // queue this chunk's upload on the copy stream
            if (e > a) {
                // one offset beyond the chunk: k_hash checks every sequence against its successor's offset
                const size_t n_off = (size_t)(e - a) + (e < n_seqs ? 1 : 0);
                if (dense) { // lengths only (one beyond the chunk feeds the scan's last output); offsets by prefix sum
                    { int r = upload_lengths(ctx, b, a, n_off, cs, cn); if (r) return r; }
                    n_len_up = n_off;
                } else {
                    CK(cudaMemcpyAsync((uint64_t *)b.seq_off + a, b.h_seq_off + a, n_off * 8, cudaMemcpyHostToDevice, cs));
                    { int r = upload_lengths(ctx, b, a, (size_t)(e - a), cs, cn); if (r) return r; }
                    n_len_up = (size_t)(e - a);
                    cn.bytes_h2d += n_off * 8;
                }
                if (prepacked) {
                    // the words of the new sequences, as they are; flags and raw bytes only for the listed exceptions
                    if (b1 > b0) {
                        const size_t w0 = (size_t)(b0 >> 4), w1 = (size_t)((b1 + 15) >> 4);
                        CK(cudaMemcpyAsync((uint32_t *)ctx->packed.p + w0, b.h_packed + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, cs));
                        cn.bytes_h2d += (uint64_t)(w1 - w0) * 4;
                    }
                    CK(cudaMemsetAsync((uint8_t *)ctx->seq_flag.p + a, 0, (size_t)(e - a), cs));
                    static const uint8_t one = 1;
                    for (size_t q = 0; q < exc_at.size(); ++q, ++exc_i) {
                        const uint32_t si = b.h_exc_seq[exc_i];
                        const uint64_t o = exc_at[q].first;
                        const uint32_t ln = exc_at[q].second;
                        CK(cudaMemcpyAsync((uint8_t *)ctx->seq_flag.p + si, &one, 1, cudaMemcpyHostToDevice, cs));
                        if (ln && o <= b.blob_bytes && ln <= b.blob_bytes - o) {
                            CK(cudaMemcpyAsync((uint8_t *)b.blob + o, b.h_exc_bytes + exc_pos, ln, cudaMemcpyHostToDevice, cs));
                            cn.bytes_h2d += ln;
                        }
                        exc_pos += ln;
                        cn.bytes_h2d += 1;
                    }
                } else if (!host_pack) {
                    if (b1 > b0) CK(cudaMemcpyAsync((uint8_t *)b.blob + b0, b.h_blob + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
                    cn.bytes_h2d += b1 - b0;
                } else {
                    const uint8_t *hf = (const uint8_t *)ctx->h_flag;
                    // Walk this chunk's blocks: packed ones ship as 2-bit words, the share left to the GPU ships as
                    // ASCII for k_pack.  Runs of equal state become one copy.
                    raw_ranges.clear();
                    const auto t_wait = std::chrono::steady_clock::now();
                    if (b1 > b0) {
                        const uint64_t BLK = HostPacker::BLOCK;
                        size_t i = (size_t)(b0 / BLK);
                        const size_t i_end = (size_t)((b1 - 1) / BLK);
                        // Blocks the workers have not reached yet can be left to the GPU: they cross PCIe as ASCII and
                        // k_pack packs them.  Pays when the host threads are the bottleneck and the link has room (several
                        // ranks sharing one host's cores).  Automatic mode: whenever the copy stream has drained -- the link
                        // would idle -- this thread takes about one chunk's worth of blocks AHEAD of the chunk the workers
                        // are on, so the split between host-packed and GPU-packed bytes follows the two rates.
                        // TDB_RAW_EVERY=n: every n-th chunk instead (deterministic, used by the tests).
                        if (ctx->allow_steal && K > 1) {
                            if (ctx->steal_forced) {
                                if (k % (size_t)ctx->steal_div == (size_t)ctx->steal_div - 1)
                                    for (size_t q = i; q <= i_end; ++q) packer.try_steal(q);
                            } else if (k >= 1 && cudaEventQuery(ctx->up_ev[k - 1]) == cudaSuccess) {
                                const size_t per_chunk = std::max<size_t>(1, packer.n_blocks() / K);
                                const size_t from = i_end + 1 + per_chunk;
                                for (size_t q = from; q < std::min(packer.n_blocks(), from + per_chunk); ++q) packer.try_steal(q);
                            }
                        }
                        while (i <= i_end) {
                            const uint8_t st0 = packer.acquire(i);
                            size_t j = i + 1; // extend the run of equal state
                            while (j <= i_end && packer.peek(j) == st0) ++j;
                            const uint64_t x0 = std::max<uint64_t>(b0, (uint64_t)i * BLK), x1 = std::min<uint64_t>(b1, (uint64_t)j * BLK);
                            if (st0 == HostPacker::PACKED) {
                                const size_t w0 = (size_t)(x0 >> 4), w1 = (size_t)((x1 + 15) >> 4);
                                CK(cudaMemcpyAsync((uint32_t *)ctx->packed.p + w0, (const uint32_t *)ctx->h_packed + w0, (w1 - w0) * 4,
                                                   cudaMemcpyHostToDevice, cs));
                                cn.bytes_h2d += (uint64_t)(w1 - w0) * 4;
                            } else {
                                cn.bytes_raw += x1 - x0;
                                raw_ranges.push_back({x0, x1});
                            }
                            i = j;
                        }
                    }
                    cn.ms_host_pack_wait += ms_since(t_wait);
                    // The share left to the GPU ships as ASCII, widened to whole sequences: a sequence k_pack flags there
                    // may reach into a neighbouring host-packed block, and flagged sequences are compared as raw bytes.
                    if (!raw_ranges.empty()) {
                        std::vector<std::pair<uint64_t, uint64_t>> ext(raw_ranges);
                        uint64_t pos = off_a;
                        size_t ri = 0;
                        for (uint32_t i = a; i < e && ri < raw_ranges.size(); ++i) {
                            if (!dense) pos = b.h_seq_off[i];
                            const uint64_t nxt = pos + b.h_seq_len[i];
                            while (ri < raw_ranges.size() && raw_ranges[ri].second <= pos) ++ri;
                            for (size_t r = ri; r < raw_ranges.size() && raw_ranges[r].first < nxt; ++r)
                                if (pos < raw_ranges[r].second) ext[r].first = std::min(ext[r].first, pos), ext[r].second = std::max(ext[r].second, nxt);
                            pos = nxt;
                        }
                        for (const auto &x : ext) {
                            const uint64_t y0 = std::max<uint64_t>(x.first, b0), y1 = std::min<uint64_t>(x.second, b.blob_bytes);
                            if (y1 > y0) {
                                CK(cudaMemcpyAsync((uint8_t *)b.blob + y0, b.h_blob + y0, y1 - y0, cudaMemcpyHostToDevice, cs));
                                cn.bytes_h2d += y1 - y0;
                            }
                        }
                    }
                    // flags found by the host packers (almost always none: one vectorised pass instead of a copy)
                    const bool any_flag = memchr(hf + a, 1, (size_t)(e - a)) != nullptr;
                    if (any_flag) {
                        CK(cudaMemcpyAsync((uint8_t *)ctx->seq_flag.p + a, hf + a, (size_t)(e - a), cudaMemcpyHostToDevice, cs));
                        cn.bytes_h2d += (uint64_t)(e - a);
                    } else {
                        CK(cudaMemsetAsync((uint8_t *)ctx->seq_flag.p + a, 0, (size_t)(e - a), cs));
                    }
                    // sequences with bytes outside ACGT are compared as raw bytes on the device: ship those bytes
                    size_t n_flag = 0;
                    if (any_flag)
                        for (uint32_t i = a; i < e; ++i) n_flag += hf[i] != 0;
                    const uint64_t *hoff = !n_flag ? nullptr : dense ? dense_off.get() : b.h_seq_off;
                    if (n_flag > 256) {
                        if (b1 > b0) CK(cudaMemcpyAsync((uint8_t *)b.blob + b0, b.h_blob + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
                        cn.bytes_h2d += b1 - b0;
                    } else if (n_flag) {
                        for (uint32_t i = a; i < e; ++i) {
                            const uint64_t o = hoff[i];
                            if (hf[i] && b.h_seq_len[i] && o <= b.blob_bytes && b.h_seq_len[i] <= b.blob_bytes - o) {
                                CK(cudaMemcpyAsync((uint8_t *)b.blob + o, b.h_blob + o, b.h_seq_len[i], cudaMemcpyHostToDevice, cs));
                                cn.bytes_h2d += b.h_seq_len[i];
                            }
                        }
                    }
                }
            }
            if (c.rle) {
                PairRun *dr = (PairRun *)ctx->runs.p + c.run0; // the parts' sub-slices land back to back
                size_t at = 0;
                for (int pt = 0; pt < SCAN_PARTS; ++pt) {
                    if (!c.part_runs[pt]) continue;
                    CK(cudaMemcpyAsync(dr + at, (const PairRun *)ctx->h_runs + c.run0 + (size_t)c.run_cap * pt,
                                       (size_t)c.part_runs[pt] * sizeof(PairRun), cudaMemcpyHostToDevice, cs));
                    at += c.part_runs[pt];
                }
                cn.bytes_h2d += (uint64_t)c.n_runs * sizeof(PairRun);
            } else {
                CK(cudaMemcpyAsync((uint32_t *)b.pp + c.p0, b.h_pp + c.p0, (c.p1 - c.p0) * 4, cudaMemcpyHostToDevice, cs));
                CK(cudaMemcpyAsync((uint32_t *)b.pt + c.p0, b.h_pt + c.p0, (c.p1 - c.p0) * 4, cudaMemcpyHostToDevice, cs));
                cn.bytes_h2d += (uint64_t)(c.p1 - c.p0) * 8;
            }
        }
        // Everything that is a KERNEL on this chunk's fresh data -- widening of uint16 lengths, the offsets' prefix sum, the
        // expansion of the run-length encoded pair list, 2-bit packing (unless the host did it), content hashes -- runs
        // on a preparation stream behind the chunk's copies, so that the copy stream carries nothing but copies and the
        // link never waits for a kernel (68 chunks x a few kernels had cost ~8 ms of link time per 937 k loci).  The
        // preparation stream is in order: sequences become ready in order whichever slot needs them.
        cudaStream_t ps = b.h_pp ? ctx->prep_stream : s;
        if (b.h_pp) {
            CK(cudaEventRecord(ctx->cp_ev[k], cs));
            CK(cudaStreamWaitEvent(ps, ctx->cp_ev[k], 0));
            if (n_len_up && !b.h_seq_len) {
                k_widen_len<<<(unsigned)((n_len_up + 255) / 256), 256, 0, ps>>>((const uint16_t *)ctx->len16.p + a,
                                                                                (uint32_t *)b.seq_len + a, (uint32_t)n_len_up);
                CK(cudaGetLastError());
                cn.kernel_launches += 1;
            }
            if (dense && n_len_up) {
                int r = scan_offsets(ctx, b.seq_len + a, (uint64_t *)b.seq_off + a, n_len_up, off_a, ps);
                if (r) return r;
                cn.kernel_launches += 1;
            }
            if (c.rle) {
                k_expand_runs<<<(cpairs + 255) / 256, 256, 0, ps>>>((const PairRun *)ctx->runs.p + c.run0, c.n_runs, (uint32_t)c.p0,
                                                                     cpairs, (uint32_t *)b.pp, (uint32_t *)b.pt);
                CK(cudaGetLastError());
                cn.kernel_launches += 1;
            }
        }
        if (e > a) {
            for (const auto &rr : raw_ranges) { // blocks taken over from the host packer: 2-bit pack them here
                const uint64_t x0 = rr.first & ~(uint64_t)15, x1 = rr.second;
                const unsigned blocks = (unsigned)std::min<uint64_t>(((x1 - x0 + 15) / 16 + 511) / 512, (uint64_t)ctx->sm_count * 32);
                k_pack<<<blocks, 256, 0, ps>>>(b.blob, x0, x1, b.seq_off, b.seq_len, a, e, (uint32_t *)ctx->packed.p,
                                               (uint8_t *)ctx->seq_flag.p);
                CK(cudaGetLastError());
                cn.kernel_launches += 1;
            }
            if (!host_pack && !prepacked) {
                CK(cudaMemsetAsync((uint8_t *)ctx->seq_flag.p + a, 0, (size_t)(e - a), ps));
                if (b1 > b0) {
                    const unsigned blocks = (unsigned)std::min<uint64_t>(((b1 - b0 + 15) / 16 + 511) / 512, (uint64_t)ctx->sm_count * 32);
                    k_pack<<<blocks, 256, 0, ps>>>(b.blob, b0, b1, b.seq_off, b.seq_len, a, e, (uint32_t *)ctx->packed.p,
                                                   (uint8_t *)ctx->seq_flag.p);
                    CK(cudaGetLastError());
                    cn.kernel_launches += 1;
                }
            }
            const unsigned blocks = (unsigned)(((size_t)(e - a) + 255) / 256);
            k_hash<<<blocks, 256, 0, ps>>>(b.seq_off, b.seq_len, a, e, n_seqs, b.blob_bytes, (const uint32_t *)ctx->packed.p,
                                           (uint8_t *)ctx->seq_flag.p, (uint64_t *)ctx->seq_hash.p, &ctrl->err);
            CK(cudaGetLastError());
            cn.kernel_launches += 1;
        }
        const int si = (int)(k % (size_t)n_slot);
        tdb_ctx::Slot &sl = ctx->slot[si];
        cudaStream_t st = sl.st;
        if (b.h_pp) {
            CK(cudaEventRecord(ctx->up_ev[k], ps));
            CK(cudaStreamWaitEvent(st, ctx->up_ev[k], 0));
        }
        if (staged) CK(cudaEventRecord(ctx->ev[2], st));
        if (cigar_mode) continue; // CIGAR output lives in the general kernel: no dedup, no shared-memory tiers
        Ctrl::Chunk *ck = &ctrl->ck[si];
        CK(cudaMemsetAsync(ck, 0, sizeof(Ctrl::Chunk), st));
        // duplicate elimination + classification: windows of DW_PAIRS consecutive pairs, tables in shared memory
        ClassifyOut co;
        co.rep = (uint32_t *)ctx->rep.p, co.key = (uint8_t *)sl.skey_in.p;
        co.glist = (uint32_t *)ctx->list_b.p, co.gcount = &ctrl->gcount, co.hist = ck->hist, co.n_closed = &ctrl->n_closed;
        co.out_score = b.score, co.out_status = b.status, co.out_penalty = b.penalty, co.work = p.work;
        k_dedup_window<<<(cpairs + DW_PAIRS - 1) / DW_PAIRS, DW_THREADS, sizeof(DedupSmem), st>>>(
            b.pp, b.pt, (uint32_t)c.p0, (uint32_t)c.p1, n_seqs, (const uint64_t *)ctx->seq_hash.p, b.seq_len,
            (const uint8_t *)ctx->seq_flag.p, b.seq_off, (const uint32_t *)ctx->packed.p, dedup ? 1 : 0, quad_ok ? 1 : 0,
            mini_ok ? 1 : 0, T0_MAXLEN, cf, co, &ctrl->err);
        CK(cudaGetLastError());
        k_plan<<<1, 32, 0, st>>>(ck->hist, ck->seg, ck->base);
        CK(cudaGetLastError());
        k_bucket<<<(cpairs + 256 * BUCKET_PER_THREAD - 1) / (256 * BUCKET_PER_THREAD), 256, 0, st>>>(
            (const uint8_t *)sl.skey_in.p, (uint32_t)c.p0, cpairs, ck->base, ck->cursor, (uint32_t *)sl.sval_out.p);
        CK(cudaGetLastError());
        cn.kernel_launches += 3; // dedup window, plan, bucket
        if (staged) CK(cudaEventRecord(ctx->ev[3], st));
        p.src1 = (const uint32_t *)sl.sval_out.p;
        // thread-per-pair tier: the lightest pairs (|n-m| < 8), up to penalty 24; the rest moves on to tier 0
        if (mini_ok) {
            p.src1_seg = ck->seg + 4, p.src2 = nullptr, p.src2_count = nullptr;
            if (default_pen && !ctx->disable_est) {
                // order by an upper bound of the penalty (k_estimate); pairs bounded beyond the tier's last score go straight on
                const unsigned eg = (unsigned)std::min<size_t>(((size_t)cpairs / 8 + 255) / 256 + 1, (size_t)ctx->sm_count * 8);
                k_estimate<<<eg, 256, 0, st>>>((const uint32_t *)ctx->packed.p, b.seq_off, b.seq_len, b.pp, b.pt,
                                               (const uint32_t *)sl.sval_out.p, ck->seg + 4, TH_SMAX, (uint8_t *)sl.skey_in.p, ck->hist2,
                                               (uint32_t *)sl.list_m.p, &ck->mcount);
                CK(cudaGetLastError());
                k_plan2<<<1, 32, 0, st>>>(ck->hist2, ck->base2, ck->seg2);
                CK(cudaGetLastError());
                k_bucket_list<<<eg, 256, 0, st>>>((const uint8_t *)sl.skey_in.p, (const uint32_t *)sl.sval_out.p, ck->seg + 4, ck->base2,
                                                  ck->cursor2, (uint32_t *)sl.list_t.p);
                CK(cudaGetLastError());
                cn.kernel_launches += 3;
                p.src1 = (const uint32_t *)sl.list_t.p, p.src1_seg = ck->seg2;
            }
            p.work_counter = &ck->wc_mini, p.next_todo = (uint32_t *)sl.list_m.p, p.next_count = &ck->mcount;
            p.counters = ctrl->exec[4];
            const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->mini_ctas_per_sm,
                                                             ((size_t)cpairs + TH_NT - 1) / TH_NT + 1);
            k_align_thread<8, 4, 2, 24, 1><<<grid, TH_NT, th_smem_bytes(), st>>>(p);
            CK(cudaGetLastError());
            cn.kernel_launches += 1;
        }
        if (staged) CK(cudaEventRecord(ctx->ev[11], st));
        // tier 0: four pairs per warp, shared memory (default penalties only)
        if (quad_ok) {
            p.src1 = (const uint32_t *)sl.sval_out.p, p.src1_seg = ck->seg + 2;
            p.src2 = mini_ok ? (const uint32_t *)sl.list_m.p : nullptr, p.src2_count = mini_ok ? &ck->mcount : nullptr;
            p.work_counter = &ck->wc_quad, p.next_todo = (uint32_t *)sl.list_a.p, p.next_count = &ck->acount;
            p.counters = ctrl->exec[0];
            const size_t smem = lock_smem_bytes<QuadCfg>();
            const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->quad_ctas_per_sm,
                                                             ((size_t)cpairs + 31) / 32 + 1);
            if (ctx->old_lock && ctx->quad8) {
                const unsigned g8 = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->quad8_ctas_per_sm,
                                                               ((size_t)cpairs + 8 * Q_WARPS - 1) / (8 * Q_WARPS) + 1);
                ENSURE(ctx->lockq_scratch[si], (size_t)ctx->sm_count * ctx->quad8_ctas_per_sm * Q_WARPS * 8 * lock_gscratch_bytes<Quad8Cfg>());
                p.scratch = (uint8_t *)ctx->lockq_scratch[si].p;
                k_align_lock<Quad8Cfg, 8, 4, 2, 24, 1><<<g8, Q_WARPS * 32, lock_smem_bytes<Quad8Cfg>(), st>>>(p);
            } else if (ctx->old_lock) {
                if (QuadCfg::GP) {
                    ENSURE(ctx->lockq_scratch[si], (size_t)ctx->sm_count * ctx->quad_ctas_per_sm * Q_WARPS * 4 * lock_gscratch_bytes<QuadCfg>());
                    p.scratch = (uint8_t *)ctx->lockq_scratch[si].p;
                }
                k_align_lock<QuadCfg, 8, 4, 2, 24, 1><<<grid, Q_WARPS * 32, smem, st>>>(p);
            }
            else {
                const unsigned fgrid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->fquad_ctas_per_sm,
                                                                  ((size_t)cpairs + 32 * F_WARPS - 1) / (32 * F_WARPS) + 1);
                k_align_frame<QuadF, 8, 4, 2, 24, 1><<<fgrid, F_WARPS * 32, frame_smem_bytes<QuadF>(), st>>>(p);
            }
            CK(cudaGetLastError());
            cn.kernel_launches += 1;
        }
        if (staged) CK(cudaEventRecord(ctx->ev[4], st));
        // tier 1: wavefronts up to 64 diagonals, shared memory, one pair per warp; the quad tier's leftovers (the
        // heaviest pairs) go first.  Production penalties: two pairs per warp in lockstep (TDB_DISABLE_DUO=1: one per warp).
        {
            p.src1 = (const uint32_t *)sl.sval_out.p;
            p.src1_seg = ck->seg + 0, p.src2 = (const uint32_t *)sl.list_a.p, p.src2_count = &ck->acount;
            p.work_counter = &ck->wc_warp, p.next_todo = (uint32_t *)ctx->list_b.p, p.next_count = &ctrl->gcount;
            p.counters = ctrl->exec[1];
            if (default_pen && !ctx->disable_duo) {
                const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->duo_ctas_per_sm,
                                                                 ((size_t)cpairs + 3) / 4 + 1);
                if (ctx->old_lock && ctx->duo4) {
                    const unsigned g4 = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->duo4_ctas_per_sm,
                                                                   ((size_t)cpairs + 8 * Q_WARPS - 1) / (8 * Q_WARPS) + 1);
                    ENSURE(ctx->lock_scratch[si], (size_t)ctx->sm_count * ctx->duo4_ctas_per_sm * Q_WARPS * 4 * lock_gscratch_bytes<Duo4Cfg>());
                    p.scratch = (uint8_t *)ctx->lock_scratch[si].p;
                    k_align_lock<Duo4Cfg, 8, 4, 2, 24, 1><<<g4, Q_WARPS * 32, lock_smem_bytes<Duo4Cfg>(), st>>>(p);
                } else if (ctx->old_lock) k_align_lock<DuoCfg, 8, 4, 2, 24, 1><<<grid, Q_WARPS * 32, lock_smem_bytes<DuoCfg>(), st>>>(p);
                else {
                    const unsigned fgrid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->fduo_ctas_per_sm,
                                                                      ((size_t)cpairs + 4 * F_WARPS - 1) / (4 * F_WARPS) + 1);
                    k_align_frame<DuoF, 8, 4, 2, 24, 1><<<fgrid, F_WARPS * 32, frame_smem_bytes<DuoF>(), st>>>(p);
                }
            } else {
                const size_t smem = sizeof(T0Smem) * T0_WARPS;
                const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * ctx->t0_ctas_per_sm,
                                                                 ((size_t)cpairs + T0_WARPS - 1) / T0_WARPS);
                k_align_warp<<<grid, T0_WARPS * 32, smem, st>>>(p);
            }
            CK(cudaGetLastError());
            cn.kernel_launches += 1;
        }
        if (staged) CK(cudaEventRecord(ctx->ev[5], st));
    }
    if (b.late_bytes) {
        if (b.h_pp) {
            CK(cudaMemcpyAsync(b.late_dst, b.late_src, b.late_bytes, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(ctx->ev[12], cs));
            CK(cudaStreamWaitEvent(s, ctx->ev[12], 0));
        } else {
            CK(cudaMemcpyAsync(b.late_dst, b.late_src, b.late_bytes, cudaMemcpyHostToDevice, s));
        }
    }
    cn.ms_host_enqueued = ms_since(t_call);
    if (packer.last_pack_ns())
        cn.ms_host_packed = std::chrono::duration<float, std::milli>(
                                std::chrono::steady_clock::duration(packer.last_pack_ns()) - t_call.time_since_epoch()).count();
    for (int i = 1; i < n_slot; ++i) { // join the other slot's stream
        CK(cudaEventRecord(ctx->slot[i].done, ctx->slot[i].st));
        CK(cudaStreamWaitEvent(s, ctx->slot[i].done, 0));
    }

    // ---- global-scratch tiers: pairs too long or too wide for shared memory, once per call ------------
    const uint64_t cap = ctx->cfg.scratch_bytes ? ctx->cfg.scratch_bytes : (48ull << 30);
    const TierCfg gtiers[2] = {
        {10, 1 << 16, 1 << 17, 8u << 20, ctx->sm_count * ctx->global_warps_per_sm}, // W=1024, S<65536, 8 MiB provenance
        {15, 1 << 18, 1 << 19, 192u << 20, ctx->sm_count * 1}, // W=32768, S<262144, 192 MiB provenance
    };
    TierCfg gt[2] = {gtiers[0], gtiers[1]};
    if (ctx->test_max_score > 0)
        for (auto &t : gt) t.s_cap = std::min(t.s_cap, ctx->test_max_score);
    Ctrl h;
    CK(cudaMemcpyAsync(&h, ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h.err & ERR_LAYOUT)
        return fail(ctx, TDB_E_INVALID, "sequence blob layout: sequences must be stored in index order, non-overlapping, inside the blob");
    if (h.err & ERR_PAIR_INDEX) return fail(ctx, TDB_E_INVALID, "pair index out of range");
    unsigned pending = cigar_mode ? n_pairs : h.gcount;
    if (staged) CK(cudaEventRecord(ctx->ev[6], s));
    // ---- long tier: one pair per warp, wavefronts (up to 192 diagonals) in shared memory, provenance / step records / op
    // stream in a per-warp slice of global scratch; production penalties, sequences up to 16 000 bases.  What it cannot
    // hold moves on to the global-scratch tiers.
    const uint32_t *g_src = cigar_mode ? nullptr : (const uint32_t *)ctx->list_b.p;
    if (!cigar_mode && default_pen && !ctx->disable_long && pending > 0) {
        const size_t per_pair = lock_gscratch_bytes<LongCfg>();
        uint64_t warps = std::min<uint64_t>((uint64_t)ctx->sm_count * ctx->long_ctas_per_sm * Q_WARPS,
                                            std::max<uint64_t>(Q_WARPS, (cap / 2) / per_pair));
        warps = std::min<uint64_t>(warps, ((uint64_t)pending + Q_WARPS - 1) / Q_WARPS * Q_WARPS);
        const unsigned grid = (unsigned)std::max<uint64_t>(1, warps / Q_WARPS);
        ENSURE(ctx->scratch[0], per_pair * (size_t)grid * Q_WARPS);
        ENSURE(ctx->list_d, (size_t)pending * 4);
        if ((uint64_t)pending > (uint64_t)ctx->long_sort_per_warp * warps) { // many pairs per warp: heaviest first (one CTA sorts: ~80 us per 13 k entries, which a
                                               // short launch does not win back; list_c is free until tier 2 runs)
            ENSURE(ctx->list_c, (size_t)pending * 4);
            k_sort_long<<<1, 1024, 0, s>>>(g_src, pending, b.pp, b.pt, b.seq_len, (uint32_t *)ctx->list_c.p);
            CK(cudaGetLastError());
            cn.kernel_launches += 1;
            g_src = (const uint32_t *)ctx->list_c.p;
        }
        p.scratch = (uint8_t *)ctx->scratch[0].p;
        p.src1 = g_src, p.src1_seg = nullptr, p.src1_count = pending, p.src2 = nullptr, p.src2_count = nullptr;
        p.next_todo = (uint32_t *)ctx->list_d.p, p.next_count = &ctrl->dcount;
        p.work_counter = &ctrl->wc_long, p.counters = ctrl->exec[5];
        k_align_lock<LongCfg, 8, 4, 2, 24, 1><<<grid, Q_WARPS * 32, lock_smem_bytes<LongCfg>(), s>>>(p);
        CK(cudaGetLastError());
        cn.kernel_launches += 1;
        if (staged) CK(cudaEventRecord(ctx->ev[14], s));
        unsigned next = 0;
        CK(cudaMemcpyAsync(&next, &ctrl->dcount, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (staged) cn.ms_long_tier = elapsed(ctx->ev[6], ctx->ev[14]);
        pending = next;
        g_src = (const uint32_t *)ctx->list_d.p;
    }
    for (int g = 0; g < 2 && pending > 0; ++g) {
        const TierCfg tc = gt[g];
        const size_t stride = global_tier_stride(tc.wlog2, tc.s_cap, tc.ops_cap, tc.prov_cap);
        uint64_t workers = std::min<uint64_t>((uint64_t)tc.workers, std::max<uint64_t>(1, (cap / 2) / stride));
        workers = std::min<uint64_t>(workers, pending);
        workers = (workers + 3) & ~(uint64_t)3; // whole CTAs of 4 warps
        ENSURE(ctx->scratch[g], stride * (size_t)workers);
        p.scratch = (uint8_t *)ctx->scratch[g].p, p.scratch_stride = stride;
        p.wlog2 = tc.wlog2, p.s_cap = tc.s_cap, p.ops_cap = tc.ops_cap, p.prov_cap = tc.prov_cap;
        p.src1_seg = nullptr, p.src1_count = pending, p.src2 = nullptr, p.src2_count = nullptr;
        if (g == 0) {
            p.src1 = g_src;
            ENSURE(ctx->list_c, (size_t)pending * 4);
            p.next_todo = (uint32_t *)ctx->list_c.p, p.next_count = &ctrl->ccount;
        } else {
            p.src1 = (const uint32_t *)ctx->list_c.p;
            p.next_todo = nullptr, p.next_count = nullptr;
        }
        p.work_counter = &ctrl->wc_g[g], p.counters = ctrl->exec[2 + g];
        const unsigned grid = (unsigned)(workers / 4);
        if (cigar_mode) k_align_global<true><<<grid, 128, 0, s>>>(p);
        else k_align_global<false><<<grid, 128, 0, s>>>(p);
        CK(cudaGetLastError());
        cn.kernel_launches += 1;
        unsigned next = 0;
        if (g == 0) {
            CK(cudaMemcpyAsync(&next, &ctrl->ccount, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
        pending = next;
    }
    if (staged) CK(cudaEventRecord(ctx->ev[7], s));
    if (!cigar_mode) {
        const unsigned blocks = (unsigned)std::min<size_t>(((size_t)n_pairs + 255) / 256, (size_t)ctx->sm_count * 16);
        k_scatter<<<blocks, 256, 0, s>>>((const uint32_t *)ctx->rep.p, n_pairs, b.score, b.status, b.penalty,
                                         (const uint32_t *)p.work, ctrl->algo, &ctrl->n_failed);
        CK(cudaGetLastError());
        cn.kernel_launches += 1;
    }
    CK(cudaEventRecord(ctx->ev[1], s));
    CK(cudaMemcpyAsync(&h, ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t aligned = 0;
    for (int t = 0; t < 4; ++t) {
        cn.exec_cells += h.exec[t][0], cn.exec_ext16 += h.exec[t][1], cn.exec_columns += h.exec[t][2];
        cn.tier_cells[t] = h.exec[t][0], cn.tier_ext16[t] = h.exec[t][1], cn.tier_columns[t] = h.exec[t][2];
        cn.tier_pairs[t] = h.exec[t][3];
        aligned += h.exec[t][3];
    }
    cn.exec_cells += h.exec[4][0], cn.exec_ext16 += h.exec[4][1], cn.exec_columns += h.exec[4][2];
    cn.thread_cells = h.exec[4][0], cn.thread_ext16 = h.exec[4][1], cn.thread_columns = h.exec[4][2];
    cn.thread_pairs = h.exec[4][3];
    aligned += h.exec[4][3];
    cn.exec_cells += h.exec[5][0], cn.exec_ext16 += h.exec[5][1], cn.exec_columns += h.exec[5][2];
    cn.long_cells = h.exec[5][0], cn.long_ext16 = h.exec[5][1], cn.long_columns = h.exec[5][2];
    cn.long_pairs = h.exec[5][3];
    aligned += h.exec[5][3];
    if (cigar_mode) {
        cn.cells = cn.exec_cells, cn.ext16 = cn.exec_ext16, cn.columns = cn.exec_columns;
    } else {
        cn.cells = h.algo[0], cn.ext16 = h.algo[1], cn.columns = h.algo[2];
        cn.pairs_duplicate = n_pairs - h.algo[3];
        cn.pairs_closed_form = h.n_closed;
        const uint64_t settled = aligned + h.n_closed + h.n_failed; // representatives that are not pattern == text
        cn.pairs_identical = h.algo[3] >= settled ? h.algo[3] - settled : 0;
    }
    cn.pairs_failed = cigar_mode ? h.n_failed : h.algo[4]; // duplicates of a failed content count too
    cn.ms_total_device = elapsed(ctx->ev[0], ctx->ev[1]);
    cn.ms_host_call = ms_since(t_call);
    if (staged) {
        cn.ms_pack = elapsed(ctx->ev[0], ctx->ev[2]);
        if (!cigar_mode) {
            cn.ms_dedup = elapsed(ctx->ev[2], ctx->ev[3]);
            cn.ms_thread_tier = elapsed(ctx->ev[3], ctx->ev[11]);
            cn.ms_tier[0] = elapsed(ctx->ev[11], ctx->ev[4]);
            cn.ms_tier[1] = elapsed(ctx->ev[4], ctx->ev[5]);
            cn.ms_scatter = elapsed(ctx->ev[7], ctx->ev[1]);
        }
        cn.ms_tier[2] = elapsed(ctx->ev[6], ctx->ev[7]);
        cn.ms_align = cn.ms_thread_tier + cn.ms_tier[0] + cn.ms_tier[1] + cn.ms_tier[2];
    }
    return TDB_OK;
}

constexpr unsigned MASK_HITS_CAP = 1u << 20; // set mask bytes a host call fetches as a list (beyond: the whole mask)
// h_out != nullptr (fused host calls): the rows go to the host as they are produced -- the loci are reduced in slices and
// the download of a slice (copy stream) runs under the reduction of the next one.
int run_reduce_device(tdb_ctx *ctx, const tdb_trio_locus *dloci, size_t n_loci, const int32_t *dscores, double q,
                      int is_duo, tdb_allele_out *dout, uint8_t *dmask, unsigned *hits = nullptr, tdb_allele_out *h_out = nullptr) {
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    CK(cudaEventRecord(ctx->ev[8], s));
    const size_t n_slices = h_out && n_loci >= (1u << 16) ? 8 : 1;
    for (size_t k = 0; k < n_slices && n_loci; ++k) {
        const size_t l0 = n_loci * k / n_slices, l1 = n_loci * (k + 1) / n_slices;
        const unsigned threads = 128, blocks = (unsigned)((l1 - l0 + threads - 1) / threads);
        k_reduce<<<blocks, threads, 0, s>>>(dloci + l0, (uint32_t)(l1 - l0), dscores, q, is_duo, dout + 2 * l0, dmask, hits,
                                            hits ? hits + 1 : nullptr, hits ? MASK_HITS_CAP : 0u);
        CK(cudaGetLastError());
        ctx->counters.kernel_launches += 1;
        if (h_out) {
            CK(cudaEventRecord(ctx->ev[13], s));
            CK(cudaStreamWaitEvent(cs, ctx->ev[13], 0));
            CK(cudaMemcpyAsync(h_out + 2 * l0, dout + 2 * l0, (l1 - l0) * 2 * sizeof(tdb_allele_out), cudaMemcpyDeviceToHost, cs));
        }
    }
    CK(cudaEventRecord(ctx->ev[9], s));
    if (h_out) CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(s));
    const float ms = elapsed(ctx->ev[8], ctx->ev[9]);
    ctx->counters.ms_reduce = ms;
    ctx->counters.ms_total_device += ms;
    return TDB_OK;
}

int upload(tdb_ctx *ctx, DevBuf &buf, const void *src, size_t bytes, size_t pad = 0) {
    ENSURE(buf, bytes + pad + 1);
    if (bytes) CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (pad) CK(cudaMemsetAsync((uint8_t *)buf.p + bytes, 0, pad, ctx->stream));
    return TDB_OK;
}
#define UPLOAD(buf, src, bytes, pad)                            \
    do {                                                        \
        int r_ = upload(ctx, (buf), (src), (bytes), (pad));     \
        if (r_ != TDB_OK) return r_;                            \
    } while (0)

int validate_common(tdb_ctx *ctx, const void *blob, const void *off, const void *len, const void *pp, const void *pt,
                    size_t n_seqs, size_t n_pairs, const void *out_score) {
    if (!ctx) return TDB_E_INVALID;
    ctx->err.clear();
    // off == NULL selects the dense layout: sequence i starts where sequence i-1 ends, sequence 0 at byte 0
    if ((n_seqs && !len) || (n_pairs && (!pp || !pt || !out_score)) || (n_pairs && !n_seqs))
        return fail(ctx, TDB_E_INVALID, "null pointer argument");
    (void)blob, (void)off;
    return TDB_OK;
}

int host_align_common(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                      const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern, const uint32_t *pair_text,
                      size_t n_pairs, DevBatch &b, bool need_blob = true) {
    int r = check_sizes(ctx, blob_bytes, n_seqs, n_pairs);
    if (r) return r;
    CK(cudaSetDevice(ctx->dev));
    ENSURE(ctx->blob, (need_blob ? blob_bytes : 0) + 64); // packed input without exceptions never touches raw bytes
    ENSURE(ctx->seq_off, n_seqs * 8 + 8);
    ENSURE(ctx->seq_len, n_seqs * 4 + 4);
    ENSURE(ctx->pair_p, n_pairs * 4 + 4);
    ENSURE(ctx->pair_t, n_pairs * 4 + 4);
    ENSURE(ctx->score, n_pairs * 4 + 4);
    ENSURE(ctx->status, n_pairs * 4 + 4);
    ENSURE(ctx->penalty, n_pairs * 4 + 4);
    memset(&b, 0, sizeof(b));
    b.blob = (const uint8_t *)ctx->blob.p, b.blob_bytes = blob_bytes;
    b.seq_off = (const uint64_t *)ctx->seq_off.p, b.seq_len = (const uint32_t *)ctx->seq_len.p, b.n_seqs = n_seqs;
    b.pp = (const uint32_t *)ctx->pair_p.p, b.pt = (const uint32_t *)ctx->pair_t.p, b.n_pairs = n_pairs;
    b.score = (int32_t *)ctx->score.p, b.status = (int32_t *)ctx->status.p, b.penalty = (int32_t *)ctx->penalty.p;
    // the chunked uploads are queued by run_align on the copy stream
    b.h_blob = seq_blob, b.h_seq_off = seq_off, b.h_seq_len = seq_len, b.h_pp = pair_pattern, b.h_pt = pair_text;
    return TDB_OK;
}

} // namespace

extern "C" {

void tdb_default_config(tdb_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->penalties = {8, 4, 2, 24, 1};
    cfg->heuristic = {TDB_HEURISTIC_WFADAPTIVE, 10, 50, 1};
    cfg->clip_len = 50;
}

int tdb_abi_version(void) { return TDB_ABI_VERSION; }

int tdb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int check_config(tdb_ctx *ctx, const tdb_penalties &pn, const tdb_heuristic &h, int clip) {
    if (pn.mismatch < 1 || pn.gap_opening1 < 0 || pn.gap_extension1 < 1 || pn.gap_opening2 < 0 ||
        pn.gap_extension2 < 1)
        return fail(ctx, TDB_E_INVALID, "penalties must be positive");
    const int scope = std::max(pn.mismatch, std::max(pn.gap_opening1 + pn.gap_extension1,
                                                     pn.gap_opening2 + pn.gap_extension2)) + 1;
    if (scope > MS || pn.gap_extension1 + 1 > GS || pn.gap_extension2 + 1 > GS)
        return fail(ctx, TDB_E_UNSUPPORTED,
                    "penalties outside kernel ring sizes (need max(x,o1+e1,o2+e2) < %d and e1,e2 < %d)", MS, GS);
    if (h.kind != TDB_HEURISTIC_NONE && h.kind != TDB_HEURISTIC_WFADAPTIVE)
        return fail(ctx, TDB_E_UNSUPPORTED, "only Heuristic::None and Heuristic::WFadaptive are implemented");
    if (h.kind == TDB_HEURISTIC_WFADAPTIVE && (h.min_wavefront_length < 1 || h.steps_between_cutoffs < 1))
        return fail(ctx, TDB_E_INVALID, "bad wf-adaptive parameters");
    if (clip < 0) return fail(ctx, TDB_E_INVALID, "clip_len must be >= 0");
    return TDB_OK;
}

int tdb_create(const tdb_config *cfg, tdb_ctx **out) {
    tdb_ctx *ctx = nullptr; // CK() reports into the thread-local slot while ctx is null
    if (!cfg || !out) return fail(nullptr, TDB_E_INVALID, "null argument");
    *out = nullptr;
    int r = check_config(nullptr, cfg->penalties, cfg->heuristic, cfg->clip_len);
    if (r) return r;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, TDB_E_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    }
    if (cfg->device_id < 0 || cfg->device_id >= ndev)
        return fail(nullptr, TDB_E_INVALID, "device_id %d out of range (%d devices)", cfg->device_id, ndev);
    CK(cudaSetDevice(cfg->device_id));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device_id));
    if (prop.major != 10)
        return fail(nullptr, TDB_E_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", cfg->device_id,
                    prop.major, prop.minor);
    tdb_ctx *c = new tdb_ctx();
    c->cfg = *cfg;
    c->dev = cfg->device_id;
    c->sm_count = prop.multiProcessorCount;
    memset(&c->counters, 0, sizeof(c->counters));
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->prep_stream, cudaStreamNonBlocking);
    c->slot[0].st = c->stream;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->slot[1].st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->slot[1].done, cudaEventDisableTiming);
    for (int i = 0; i < 16 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    const size_t smem = sizeof(T0Smem) * T0_WARPS;
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_align_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_align_warp, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int occ = 0;
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_align_warp, T0_WARPS * 32, smem);
    if (e != cudaSuccess || occ < 1) {
        int code = fail(nullptr, TDB_E_CUDA, "context setup failed: %s (occupancy %d)", cudaGetErrorString(e), occ);
        tdb_destroy(c);
        return code;
    }
    c->t0_ctas_per_sm = occ;
    {
        auto setup = [&](auto kern, size_t smem, int &occ_out) {
            cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            int occ2 = 0;
            if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, Q_WARPS * 32, smem);
            occ_out = occ2;
            return e2;
        };
        e = cudaFuncSetAttribute(k_dedup_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DedupSmem));
        if (e == cudaSuccess) e = setup(k_align_lock<QuadCfg, 8, 4, 2, 24, 1>, lock_smem_bytes<QuadCfg>(), c->quad_ctas_per_sm);
        if (e == cudaSuccess) e = setup(k_align_lock<DuoCfg, 8, 4, 2, 24, 1>, lock_smem_bytes<DuoCfg>(), c->duo_ctas_per_sm);
        if (e == cudaSuccess) e = setup(k_align_lock<Duo4Cfg, 8, 4, 2, 24, 1>, lock_smem_bytes<Duo4Cfg>(), c->duo4_ctas_per_sm);
        c->duo4 = getenv("TDB_DUO_G16") == nullptr && c->duo4_ctas_per_sm >= 1;
        if (e == cudaSuccess) e = setup(k_align_lock<Quad8Cfg, 8, 4, 2, 24, 1>, lock_smem_bytes<Quad8Cfg>(), c->quad8_ctas_per_sm);
        c->quad8 = getenv("TDB_QUAD8") != nullptr && c->quad8_ctas_per_sm >= 1;
        if (e == cudaSuccess) e = setup(k_align_lock<LongCfg, 8, 4, 2, 24, 1>, lock_smem_bytes<LongCfg>(), c->long_ctas_per_sm);
        c->disable_long = getenv("TDB_DISABLE_LONG") != nullptr || c->long_ctas_per_sm < 1;
        if (const char *ls = getenv("TDB_LONG_SORT_PER_WARP")) c->long_sort_per_warp = std::max(0, atoi(ls));
        auto fsetup = [&](auto kern, size_t smem, int &occ_out) {
            cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            int occ2 = 0;
            if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, F_WARPS * 32, smem);
            occ_out = occ2;
            return e2;
        };
        if (e == cudaSuccess) e = fsetup(k_align_frame<QuadF, 8, 4, 2, 24, 1>, frame_smem_bytes<QuadF>(), c->fquad_ctas_per_sm);
        if (e == cudaSuccess) e = fsetup(k_align_frame<DuoF, 8, 4, 2, 24, 1>, frame_smem_bytes<DuoF>(), c->fduo_ctas_per_sm);
        if (e == cudaSuccess && (c->fquad_ctas_per_sm < 1 || c->fduo_ctas_per_sm < 1)) c->quad_ctas_per_sm = 0;
        c->old_lock = getenv("TDB_FRAME_LOCK") == nullptr;
        c->disable_est = getenv("TDB_DISABLE_EST_SORT") != nullptr;
        if (e != cudaSuccess || c->quad_ctas_per_sm < 1 || c->duo_ctas_per_sm < 1) {
            int code = fail(nullptr, TDB_E_CUDA, "lockstep kernel setup failed: %s (occupancy %d / %d)", cudaGetErrorString(e),
                            c->quad_ctas_per_sm, c->duo_ctas_per_sm);
            tdb_destroy(c);
            return code;
        }
        {
            cudaError_t e2 = cudaFuncSetAttribute(k_align_thread<8, 4, 2, 24, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)th_smem_bytes());
            if (e2 == cudaSuccess)
                e2 = cudaFuncSetAttribute(k_align_thread<8, 4, 2, 24, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            if (e2 == cudaSuccess)
                e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->mini_ctas_per_sm, k_align_thread<8, 4, 2, 24, 1>, TH_NT,
                                                                   th_smem_bytes());
            if (e2 != cudaSuccess || c->mini_ctas_per_sm < 1) {
                int code = fail(nullptr, TDB_E_CUDA, "thread-per-pair kernel setup failed: %s (occupancy %d)",
                                cudaGetErrorString(e2), c->mini_ctas_per_sm);
                tdb_destroy(c);
                return code;
            }
        }
        c->disable_quad = getenv("TDB_DISABLE_QUAD") != nullptr;
        c->disable_mini = getenv("TDB_DISABLE_THREAD_TIER") != nullptr;
        if (const char *gw = getenv("TDB_GLOBAL_WARPS_PER_SM")) c->global_warps_per_sm = std::max(1, std::min(32, atoi(gw)));
        if (const char *sm = getenv("TDB_SPARSE_MASK_MIN_PAIRS")) c->sparse_mask_min_pairs = (size_t)std::max(0ll, atoll(sm));
        // tier 1 runs two pairs per warp in lockstep (16 lanes each) with the production penalties: 17 % faster than
        // one pair per warp on the 30x trio (profiles/r01/NOTES.md); TDB_DISABLE_DUO=1 falls back to k_align_warp
        c->disable_duo = getenv("TDB_DISABLE_DUO") != nullptr;
        c->disable_dedup = getenv("TDB_DISABLE_DEDUP") != nullptr;
        c->disable_closed_form = getenv("TDB_DISABLE_CLOSED_FORM") != nullptr;
        c->disable_rle = getenv("TDB_DISABLE_RLE") != nullptr;
        c->steal_auto_always = getenv("TDB_STEAL_AUTO") != nullptr;
        if (const char *sd = getenv("TDB_RAW_EVERY")) {
            c->steal_forced = true;
            c->allow_steal = atoi(sd) > 1;
            c->steal_div = std::max(2, atoi(sd));
        }
        if (const char *ms = getenv("TDB_TEST_MAX_SCORE")) c->test_max_score = atoi(ms);
        if (const char *ht = getenv("TDB_HOST_THREADS")) c->host_threads = std::max(0, std::min(256, atoi(ht)));
        if (const char *hp = getenv("TDB_HOSTPACK_MIN_BYTES")) {
            const long long v = atoll(hp); // < 0 disables the host packer
            c->hostpack_min_bytes = v < 0 ? (size_t)-1 : (size_t)v;
        }
        if (const char *cp = getenv("TDB_CHUNK_PAIRS")) {
            const long long v = atoll(cp);
            if (v >= 1024) c->chunk_pairs = (size_t)v;
        }
    }
    *out = c;
    return TDB_OK;
}

void tdb_destroy(tdb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->dev);
    DevBuf *bufs[] = {&c->blob, &c->seq_off, &c->seq_len, &c->packed, &c->seq_flag, &c->seq_hash, &c->pair_p, &c->pair_t,
                      &c->score, &c->status, &c->penalty, &c->rep, &c->work, &c->list_b, &c->list_c, &c->scan_tmp, &c->mask_hits, &c->len16, &c->ctrl, &c->loci, &c->runs,
                      &c->aout, &c->mask, &c->scores_in, &c->cigar, &c->cigar_off, &c->cigar_len, &c->scratch[0],
                      &c->list_d, &c->scratch[1], &c->lock_scratch[0], &c->lock_scratch[1], &c->lockq_scratch[0], &c->lockq_scratch[1]};
    for (DevBuf *b : bufs) b->release();
    for (auto &sl : c->slot) {
        DevBuf *sb[] = {&sl.skey_in, &sl.sval_out, &sl.list_a, &sl.list_m, &sl.list_t};
        for (DevBuf *b : sb) b->release();
    }
    if (c->slot[1].st) cudaStreamDestroy(c->slot[1].st);
    if (c->slot[1].done) cudaEventDestroy(c->slot[1].done);
    for (int i = 0; i < 16; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (cudaEvent_t e : c->up_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->cp_ev) cudaEventDestroy(e);
    if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
    if (c->h_packed) cudaFreeHost(c->h_packed);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_runs) cudaFreeHost(c->h_runs);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

const char *tdb_last_error(const tdb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tdb_set_clip_len(tdb_ctx *ctx, int32_t clip_len) {
    if (!ctx) return TDB_E_INVALID;
    if (clip_len < 0) return fail(ctx, TDB_E_INVALID, "clip_len must be >= 0");
    ctx->cfg.clip_len = clip_len;
    return TDB_OK;
}

int tdb_set_host_threads(tdb_ctx *ctx, int32_t n_threads) {
    if (!ctx) return TDB_E_INVALID;
    if (n_threads < 0 || n_threads > 256) return fail(ctx, TDB_E_INVALID, "host threads must be in 0..256");
    ctx->host_threads = n_threads;
    return TDB_OK;
}

int tdb_set_heuristic(tdb_ctx *ctx, const tdb_heuristic *h) {
    if (!ctx || !h) return TDB_E_INVALID;
    int r = check_config(ctx, ctx->cfg.penalties, *h, ctx->cfg.clip_len);
    if (r) return r;
    ctx->cfg.heuristic = *h;
    return TDB_OK;
}

int tdb_set_count_work(tdb_ctx *ctx, int32_t on) {
    if (!ctx) return TDB_E_INVALID;
    ctx->count_work = on != 0;
    return TDB_OK;
}

int tdb_get_counters(const tdb_ctx *ctx, tdb_counters *out) {
    if (!ctx || !out) return TDB_E_INVALID;
    *out = ctx->counters;
    return TDB_OK;
}

int tdb_align_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                    const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern, const uint32_t *pair_text,
                    size_t n_pairs, int32_t *out_score, int32_t *out_status) {
    int r = validate_common(ctx, seq_blob, seq_off, seq_len, pair_pattern, pair_text, n_seqs, n_pairs, out_score);
    if (r) return r;
    DevBatch b;
    r = host_align_common(ctx, seq_blob, blob_bytes, seq_off, seq_len, n_seqs, pair_pattern, pair_text, n_pairs, b);
    if (r) return r;
    r = run_align(ctx, b, false);
    if (r) return r;
    if (n_pairs) {
        CK(cudaMemcpyAsync(out_score, b.score, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_status) CK(cudaMemcpyAsync(out_status, b.status, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return TDB_OK;
}

int tdb_align_batch_device(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                           const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern,
                           const uint32_t *pair_text, size_t n_pairs, int32_t *out_score, int32_t *out_status) {
    int r = validate_common(ctx, seq_blob, seq_off, seq_len, pair_pattern, pair_text, n_seqs, n_pairs, out_score);
    if (r) return r;
    r = check_sizes(ctx, blob_bytes, n_seqs, n_pairs);
    if (r) return r;
    CK(cudaSetDevice(ctx->dev));
    DevBatch b;
    memset(&b, 0, sizeof(b));
    b.blob = seq_blob, b.blob_bytes = blob_bytes, b.seq_off = seq_off, b.seq_len = seq_len, b.n_seqs = n_seqs;
    b.pp = pair_pattern, b.pt = pair_text, b.n_pairs = n_pairs;
    b.score = out_score, b.status = out_status, b.penalty = nullptr;
    return run_align(ctx, b, false);
}

int tdb_align_batch_cigar(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                          const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern,
                          const uint32_t *pair_text, size_t n_pairs, int32_t *out_score, int32_t *out_status,
                          int32_t *out_penalty, uint8_t *cigar_buf, const uint64_t *cigar_off, uint32_t *cigar_len) {
    int r = validate_common(ctx, seq_blob, seq_off, seq_len, pair_pattern, pair_text, n_seqs, n_pairs, out_score);
    if (r) return r;
    if (n_pairs && (!cigar_buf || !cigar_off || !cigar_len)) return fail(ctx, TDB_E_INVALID, "null cigar argument");
    DevBatch b;
    r = host_align_common(ctx, seq_blob, blob_bytes, seq_off, seq_len, n_seqs, pair_pattern, pair_text, n_pairs, b);
    if (r) return r;
    // size of the cigar area = max(cigar_off[j] + plen + tlen)
    size_t total = 0;
    for (size_t j = 0; j < n_pairs; ++j) {
        if (pair_pattern[j] >= n_seqs || pair_text[j] >= n_seqs) return fail(ctx, TDB_E_INVALID, "pair index out of range");
        total = std::max<size_t>(total, cigar_off[j] + seq_len[pair_pattern[j]] + seq_len[pair_text[j]]);
    }
    ENSURE(ctx->cigar, total + 16);
    UPLOAD(ctx->cigar_off, cigar_off, n_pairs * 8, 0);
    ENSURE(ctx->cigar_len, n_pairs * 4 + 4);
    b.cigar_buf = (uint8_t *)ctx->cigar.p, b.cigar_off = (const uint64_t *)ctx->cigar_off.p;
    b.cigar_len = (uint32_t *)ctx->cigar_len.p;
    r = run_align(ctx, b, true);
    if (r) return r;
    if (n_pairs) {
        CK(cudaMemcpyAsync(out_score, b.score, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_status) CK(cudaMemcpyAsync(out_status, b.status, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_penalty) CK(cudaMemcpyAsync(out_penalty, b.penalty, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(cigar_buf, b.cigar_buf, total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(cigar_len, b.cigar_len, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return TDB_OK;
}

static int reduce_host(tdb_ctx *ctx, int is_duo, const tdb_trio_locus *loci, size_t n_loci, const int32_t *scores,
                       size_t n_scores, double q, tdb_allele_out *out, uint8_t *denovo_mask) {
    if (!ctx) return TDB_E_INVALID;
    ctx->err.clear();
    if (n_loci && (!loci || !out || (n_scores && !scores))) return fail(ctx, TDB_E_INVALID, "null pointer argument");
    if (n_loci >= 0xffffffffull) return fail(ctx, TDB_E_UNSUPPORTED, "too many loci");
    CK(cudaSetDevice(ctx->dev));
    UPLOAD(ctx->loci, loci, n_loci * sizeof(tdb_trio_locus), 0);
    UPLOAD(ctx->scores_in, scores, n_scores * 4, 0);
    ENSURE(ctx->aout, n_loci * 2 * sizeof(tdb_allele_out) + 16);
    ENSURE(ctx->mask, n_scores + 16);
    CK(cudaMemsetAsync(ctx->aout.p, 0, n_loci * 2 * sizeof(tdb_allele_out), ctx->stream));
    CK(cudaMemsetAsync(ctx->mask.p, 0, n_scores + 16, ctx->stream));
    memset(&ctx->counters, 0, sizeof(ctx->counters));
    int r = run_reduce_device(ctx, (const tdb_trio_locus *)ctx->loci.p, n_loci, (const int32_t *)ctx->scores_in.p, q,
                              is_duo, (tdb_allele_out *)ctx->aout.p, (uint8_t *)ctx->mask.p);
    if (r) return r;
    if (n_loci) CK(cudaMemcpyAsync(out, ctx->aout.p, n_loci * 2 * sizeof(tdb_allele_out), cudaMemcpyDeviceToHost, ctx->stream));
    if (denovo_mask && n_scores) CK(cudaMemcpyAsync(denovo_mask, ctx->mask.p, n_scores, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TDB_OK;
}

int tdb_trio_reduce(tdb_ctx *ctx, const tdb_trio_locus *loci, size_t n_loci, const int32_t *scores, size_t n_scores,
                    double parent_quantile, tdb_allele_out *out, uint8_t *denovo_mask) {
    return reduce_host(ctx, 0, loci, n_loci, scores, n_scores, parent_quantile, out, denovo_mask);
}
int tdb_duo_reduce(tdb_ctx *ctx, const tdb_trio_locus *loci, size_t n_loci, const int32_t *scores, size_t n_scores,
                   double parent_quantile, tdb_allele_out *out, uint8_t *denovo_mask) {
    return reduce_host(ctx, 1, loci, n_loci, scores, n_scores, parent_quantile, out, denovo_mask);
}

static int assess_tail(tdb_ctx *ctx, int is_duo, DevBatch &b, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                       double q, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);

static int assess_host(tdb_ctx *ctx, int is_duo, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                       const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern, const uint32_t *pair_text,
                       size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci, double q, tdb_allele_out *out,
                       int32_t *out_score, uint8_t *denovo_mask) {
    int32_t dummy = 0;
    int r = validate_common(ctx, seq_blob, seq_off, seq_len, pair_pattern, pair_text, n_seqs, n_pairs, &dummy);
    if (r) return r;
    if (n_loci && (!loci || !out)) return fail(ctx, TDB_E_INVALID, "null pointer argument");
    if (n_loci >= 0xffffffffull) return fail(ctx, TDB_E_UNSUPPORTED, "too many loci");
    DevBatch b;
    r = host_align_common(ctx, seq_blob, blob_bytes, seq_off, seq_len, n_seqs, pair_pattern, pair_text, n_pairs, b);
    if (r) return r;
    return assess_tail(ctx, is_duo, b, n_pairs, loci, n_loci, q, out, out_score, denovo_mask);
}

// reduction + result download of the fused host-buffer calls (ASCII and packed input alike)
static int assess_tail(tdb_ctx *ctx, int is_duo, DevBatch &b, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                       double q, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    int r;
    const auto t_fused = std::chrono::steady_clock::now();
    ENSURE(ctx->loci, n_loci * sizeof(tdb_trio_locus) + 16); // uploaded behind the last chunk (run_align)
    b.late_dst = ctx->loci.p, b.late_src = loci, b.late_bytes = n_loci * sizeof(tdb_trio_locus);
    ENSURE(ctx->aout, n_loci * 2 * sizeof(tdb_allele_out) + 16);
    ENSURE(ctx->mask, n_pairs + 16);
    ENSURE(ctx->mask_hits, ((size_t)MASK_HITS_CAP + 1) * 4);
    CK(cudaMemsetAsync(ctx->aout.p, 0, n_loci * 2 * sizeof(tdb_allele_out), ctx->stream));
    CK(cudaMemsetAsync(ctx->mask.p, 0, n_pairs + 16, ctx->stream));
    CK(cudaMemsetAsync(ctx->mask_hits.p, 0, 4, ctx->stream));
    // The mask has a few set bytes in a million: the host copy is zeroed here (chunk by chunk, while the calling
    // thread waits for the packer threads anyway) and only the list of set positions comes back over PCIe.
    const bool sparse_mask = denovo_mask && n_pairs >= ctx->sparse_mask_min_pairs;
    b.h_zero = sparse_mask ? denovo_mask : nullptr;
    r = run_align(ctx, b, false);
    if (r) return r;
    const auto t_aligned = std::chrono::steady_clock::now();
    unsigned *hits = sparse_mask ? (unsigned *)ctx->mask_hits.p : nullptr;
    r = run_reduce_device(ctx, (const tdb_trio_locus *)ctx->loci.p, n_loci, b.score, q, is_duo,
                          (tdb_allele_out *)ctx->aout.p, (uint8_t *)ctx->mask.p, hits, out);
    if (r) return r;
    uint64_t d2h = n_loci * 2 * sizeof(tdb_allele_out); // downloaded slice by slice under the reduction
    if (out_score && n_pairs) {
        CK(cudaMemcpyAsync(out_score, b.score, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        d2h += n_pairs * 4;
    }
    if (denovo_mask && n_pairs && !sparse_mask) {
        CK(cudaMemcpyAsync(denovo_mask, ctx->mask.p, n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        d2h += n_pairs;
    }
    if (sparse_mask) {
        constexpr unsigned FIRST = 1u << 14; // entries fetched together with the count
        ctx->h_hits.resize((size_t)MASK_HITS_CAP + 1);
        unsigned *hh = ctx->h_hits.data();
        CK(cudaMemcpyAsync(hh, hits, ((size_t)FIRST + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        d2h += ((size_t)FIRST + 1) * 4;
        CK(cudaStreamSynchronize(ctx->stream));
        const unsigned cnt = hh[0];
        if (cnt > MASK_HITS_CAP) { // dense after all: the whole mask
            CK(cudaMemcpyAsync(denovo_mask, ctx->mask.p, n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
            d2h += n_pairs;
        } else {
            if (cnt > FIRST) {
                CK(cudaMemcpyAsync(hh + 1 + FIRST, hits + 1 + FIRST, (size_t)(cnt - FIRST) * 4, cudaMemcpyDeviceToHost, ctx->stream));
                d2h += (size_t)(cnt - FIRST) * 4;
            }
            CK(cudaStreamSynchronize(ctx->stream));
            for (unsigned i = 0; i < cnt; ++i) denovo_mask[hh[1 + i]] = 1;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->counters.bytes_d2h = d2h;
    const auto t_end = std::chrono::steady_clock::now();
    ctx->counters.ms_fused_call = std::chrono::duration<float, std::milli>(t_end - t_fused).count();
    ctx->counters.ms_fused_tail = std::chrono::duration<float, std::milli>(t_end - t_aligned).count();
    return TDB_OK;
}

// ---- packed input ------------------------------------------------------------------------------------------------
static int packed_common(tdb_ctx *ctx, const tdb_packed_seqs *sq, const uint32_t *pair_pattern, const uint32_t *pair_text,
                         size_t n_pairs, const void *out_score, DevBatch &b) {
    if (!ctx) return TDB_E_INVALID;
    ctx->err.clear();
    if (!sq) return fail(ctx, TDB_E_INVALID, "null pointer argument");
    if ((sq->n_bases && !sq->packed) || (sq->n_exc && (!sq->exc_seq || !sq->exc_bytes)))
        return fail(ctx, TDB_E_INVALID, "null pointer argument");
    if (sq->seq_len && sq->seq_len16) return fail(ctx, TDB_E_INVALID, "give seq_len or seq_len16, not both");
    const void *any_len = sq->seq_len ? (const void *)sq->seq_len : (const void *)sq->seq_len16;
    int r = validate_common(ctx, sq->packed, sq->seq_off, any_len, pair_pattern, pair_text, sq->n_seqs, n_pairs, out_score);
    if (r) return r;
    for (size_t i = 0; i < sq->n_exc; ++i)
        if (sq->exc_seq[i] >= sq->n_seqs || (i && sq->exc_seq[i] <= sq->exc_seq[i - 1]))
            return fail(ctx, TDB_E_INVALID, "exc_seq must be ascending sequence indices");
    r = host_align_common(ctx, nullptr, sq->n_bases, sq->seq_off, sq->seq_len, sq->n_seqs, pair_pattern, pair_text, n_pairs, b,
                          sq->n_exc != 0);
    if (r) return r;
    b.h_packed = sq->packed, b.h_exc_seq = sq->exc_seq, b.h_exc_bytes = sq->exc_bytes, b.n_exc = sq->n_exc;
    b.h_seq_len16 = sq->seq_len16;
    if (sq->seq_len16) ENSURE(ctx->len16, sq->n_seqs * 2 + 16);
    return TDB_OK;
}

int tdb_align_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                           const uint32_t *pair_text, size_t n_pairs, int32_t *out_score, int32_t *out_status) {
    DevBatch b;
    int r = packed_common(ctx, seqs, pair_pattern, pair_text, n_pairs, out_score, b);
    if (r) return r;
    r = run_align(ctx, b, false);
    if (r) return r;
    if (n_pairs) {
        CK(cudaMemcpyAsync(out_score, b.score, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_status) CK(cudaMemcpyAsync(out_status, b.status, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return TDB_OK;
}

static int assess_packed(tdb_ctx *ctx, int is_duo, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                         const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci, double q,
                         tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    int32_t dummy = 0;
    DevBatch b;
    int r = packed_common(ctx, seqs, pair_pattern, pair_text, n_pairs, &dummy, b);
    if (r) return r;
    if (n_loci && (!loci || !out)) return fail(ctx, TDB_E_INVALID, "null pointer argument");
    if (n_loci >= 0xffffffffull) return fail(ctx, TDB_E_UNSUPPORTED, "too many loci");
    return assess_tail(ctx, is_duo, b, n_pairs, loci, n_loci, q, out, out_score, denovo_mask);
}
int tdb_trio_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                          const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                          double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    return assess_packed(ctx, 0, seqs, pair_pattern, pair_text, n_pairs, loci, n_loci, parent_quantile, out, out_score,
                         denovo_mask);
}
int tdb_duo_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                         const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                         double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    return assess_packed(ctx, 1, seqs, pair_pattern, pair_text, n_pairs, loci, n_loci, parent_quantile, out, out_score,
                         denovo_mask);
}

int tdb_trio_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                   const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern, const uint32_t *pair_text,
                   size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci, double parent_quantile,
                   tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    return assess_host(ctx, 0, seq_blob, blob_bytes, seq_off, seq_len, n_seqs, pair_pattern, pair_text, n_pairs, loci,
                       n_loci, parent_quantile, out, out_score, denovo_mask);
}
int tdb_duo_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                  const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern, const uint32_t *pair_text,
                  size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci, double parent_quantile,
                  tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask) {
    return assess_host(ctx, 1, seq_blob, blob_bytes, seq_off, seq_len, n_seqs, pair_pattern, pair_text, n_pairs, loci,
                       n_loci, parent_quantile, out, out_score, denovo_mask);
}

int tdb_assess_batch_device(tdb_ctx *ctx, int32_t is_duo, const uint8_t *seq_blob, size_t blob_bytes,
                            const uint64_t *seq_off, const uint32_t *seq_len, size_t n_seqs,
                            const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                            const tdb_trio_locus *loci, size_t n_loci, double parent_quantile, tdb_allele_out *out,
                            int32_t *out_score, uint8_t *denovo_mask) {
    int32_t dummy = 0;
    int r = validate_common(ctx, seq_blob, seq_off, seq_len, pair_pattern, pair_text, n_seqs, n_pairs, &dummy);
    if (r) return r;
    if (n_loci && (!loci || !out)) return fail(ctx, TDB_E_INVALID, "null pointer argument");
    r = check_sizes(ctx, blob_bytes, n_seqs, n_pairs);
    if (r) return r;
    CK(cudaSetDevice(ctx->dev));
    int32_t *dscore = out_score;
    if (!dscore) {
        ENSURE(ctx->score, n_pairs * 4 + 4);
        dscore = (int32_t *)ctx->score.p;
    }
    DevBatch b;
    memset(&b, 0, sizeof(b));
    b.blob = seq_blob, b.blob_bytes = blob_bytes, b.seq_off = seq_off, b.seq_len = seq_len, b.n_seqs = n_seqs;
    b.pp = pair_pattern, b.pt = pair_text, b.n_pairs = n_pairs;
    b.score = dscore, b.status = nullptr, b.penalty = nullptr;
    r = run_align(ctx, b, false);
    if (r) return r;
    return run_reduce_device(ctx, loci, n_loci, dscore, parent_quantile, is_duo, out, denovo_mask);
}

int tdb_align_end_to_end(tdb_ctx *ctx, const uint8_t *pattern, int32_t pattern_len, const uint8_t *text,
                         int32_t text_len) {
    if (!ctx) return TDB_E_INVALID;
    ctx->err.clear();
    if (pattern_len < 0 || text_len < 0 || (pattern_len && !pattern) || (text_len && !text))
        return fail(ctx, TDB_E_INVALID, "bad sequence argument");
    std::vector<uint8_t> blob((size_t)pattern_len + (size_t)text_len);
    if (pattern_len) memcpy(blob.data(), pattern, (size_t)pattern_len);
    if (text_len) memcpy(blob.data() + pattern_len, text, (size_t)text_len);
    const uint64_t off[2] = {0, (uint64_t)pattern_len};
    const uint32_t len[2] = {(uint32_t)pattern_len, (uint32_t)text_len};
    const uint32_t pp = 0, pt = 1;
    const uint64_t coff = 0;
    int32_t score = 0, status = TDB_STATUS_UNATTAINABLE, pen = INT32_MIN;
    uint32_t clen = 0;
    ctx->last_cigar.assign(blob.size() + 1, 0);
    const int saved_clip = ctx->cfg.clip_len;
    int r = tdb_align_batch_cigar(ctx, blob.data(), blob.size(), off, len, 2, &pp, &pt, 1, &score, &status, &pen,
                                  ctx->last_cigar.data(), &coff, &clen);
    ctx->cfg.clip_len = saved_clip;
    if (r) {
        ctx->last_status = TDB_STATUS_UNATTAINABLE, ctx->last_penalty = INT32_MIN, ctx->last_cigar.clear();
        return r;
    }
    ctx->last_cigar.resize(status == 0 ? clen : 0);
    ctx->last_status = status;
    ctx->last_penalty = pen;
    return status;
}

int32_t tdb_score(const tdb_ctx *ctx) {
    if (!ctx || ctx->last_status != 0) return INT32_MIN;
    return -ctx->last_penalty;
}

const uint8_t *tdb_cigar_ops(const tdb_ctx *ctx, int32_t *len) {
    if (len) *len = ctx ? (int32_t)ctx->last_cigar.size() : 0;
    return ctx ? ctx->last_cigar.data() : nullptr;
}

// src/aligner.rs:320-357: pure re-scoring of the CIGAR the device produced (host string scan, as the
// Rust side of the reference does over wfa2's cigar_t buffer).
int32_t tdb_cigar_score_clipped(const tdb_ctx *ctx, int32_t flank_len) {
    if (!ctx) return 0;
    const tdb_penalties &pn = ctx->cfg.penalties;
    const long begin = flank_len, end = (long)ctx->last_cigar.size() - flank_len;
    int score = 0, oplen = 0;
    uint8_t last = 0;
    auto cost = [&](uint8_t op, int len) {
        if (op == 'M') return 0;
        if (op == 'X') return len * pn.mismatch;
        return std::min(pn.gap_opening1 + pn.gap_extension1 * len, pn.gap_opening2 + pn.gap_extension2 * len);
    };
    for (long i = begin; i < end; ++i) {
        const uint8_t cur = ctx->last_cigar[(size_t)i];
        if (last && cur != last) score -= cost(last, oplen), oplen = 0;
        last = cur, ++oplen;
    }
    if (last) score -= cost(last, oplen);
    return score;
}

size_t tdb_packed_words(size_t blob_bytes) { return blob_bytes / 16 + 2; }

int tdb_host_pack(const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off, const uint32_t *seq_len,
                  size_t n_seqs, uint32_t *packed, uint8_t *flag, int n_threads) {
    if ((blob_bytes && (!seq_blob || !packed)) || (n_seqs && (!seq_off || !seq_len || !flag))) return TDB_E_INVALID;
    if (n_seqs >= 0x7fffffffull || blob_bytes / 16 + 2 >= 0xffffffffull) return TDB_E_UNSUPPORTED;
    const uint32_t n = (uint32_t)n_seqs;
    const unsigned nt = (unsigned)std::max(1, std::min(n_threads, 256));
    const uint64_t BLK = HostPacker::BLOCK;
    std::atomic<uint64_t> next{0};
    std::atomic<uint32_t> next_seq{0};
    std::atomic<size_t> violations{0};
    auto run = [&](auto fn) {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(fn);
        fn();
        for (auto &x : th) x.join();
    };
    run([&]() { // layout contract first: the packer's owner search needs ascending offsets
        for (;;) {
            const uint32_t a = next_seq.fetch_add(65536);
            if (a >= n) return;
            const uint32_t e = (uint32_t)std::min<uint64_t>(n, (uint64_t)a + 65536);
            violations += hostpack_check_layout(seq_off, seq_len, a, e, n, blob_bytes);
            memset(flag + a, 0, e - a);
        }
    });
    if (violations.load()) return TDB_E_INVALID;
    run([&]() {
        for (;;) {
            const uint64_t x0 = next.fetch_add(BLK);
            if (x0 >= blob_bytes) return;
            hostpack_stream(seq_blob, x0, std::min<uint64_t>(blob_bytes, x0 + BLK), packed, seq_off, seq_len, n, flag);
        }
    });
    return TDB_OK;
}

const char *tdb_host_pack_isa(void) { return hostpack_isa(); }

void *tdb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void tdb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

} // extern "C"
