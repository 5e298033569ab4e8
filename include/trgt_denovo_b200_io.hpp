// trgt_denovo_b200_io.hpp -- real-input front end for the C++ host side: the
// files the reference reads -- repeat catalog (BED, plain or BGZF), indexed reference FASTA, TRGT VCF
// (*.sorted.vcf.gz) and TRGT spanning-read BAM (*.spanning.sorted.bam) -- turned into the `Locus` /
// `SampleInput` carriers that `tdbh::trio::process_alleles` / `tdbh::duo::process_alleles` batch for the GPU.
//
// The reference does this with htslib (rust-htslib: faidx, bgzf, bcf::IndexedReader,
// bam::IndexedReader; src/readers.rs:6-34, src/handles.rs:64-180); htslib is not available here, so
// the formats are decoded from their specifications with zlib only.  What each piece follows:
//   read_catalog      src/locus.rs:88-140,222-259 (+ src/region.rs:12-34): 4 whitespace fields,
//                     ID= / MOTIFS= info keys, flanks = upper-cased reference around the region
//   SampleFiles::open src/handles.rs:64-180: karyotype from the BAM @PG CL field ("--karyotype", default
//                     XY), TRGT version from @PG VN / ##trgtVersion (must agree; major > 0 => alleles
//                     carry a padding base)
//   sample_input      src/allele.rs:274-426 (get_reads, get_vcf_data, process_genotypes,
//                     build_allele_seqs) and src/read.rs:32-98 (TR / AL / HP tags, accepted tag types)
// One deliberate difference: the reference makes one indexed region query per locus and sample from
// its worker threads; a batched GPU pipeline wants every locus of a shard at once, so the VCF and the BAM
// are each streamed ONCE and keyed by TRID (INFO/TRID resp. the TR tag).  Same records per locus as
// long as a TRID names one locus (which TRGT guarantees); a read without a TR tag, on which the
// reference fails the locus, is skipped here and counted in `SampleFiles::reads_without_tr`.
// Parity note: there is no htslib and no TRGT output in this environment to check against; the readers
// are tested against files written by an independent writer of the same specifications
// (tests/io_fixtures.py), and end to end against the in-memory path.
#pragma once
#include "trgt_denovo_b200.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <thread>
#include <tuple>
#include <unordered_map>

namespace tdbh {
namespace io {

struct IoError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Sequential reader over a plain file or a (multi-member) gzip file -- BGZF is a series of gzip members.
class GzStream {
  public:
    explicit GzStream(const std::string &path) : path_(path) {
        f_ = fopen(path.c_str(), "rb");
        if (!f_) throw IoError("cannot open " + path);
        in_.resize(1 << 16), out_.resize(1 << 17);
        n_in_ = fread(in_.data(), 1, in_.size(), f_);
        gz_ = n_in_ >= 2 && (unsigned char)in_[0] == 0x1f && (unsigned char)in_[1] == 0x8b;
        if (gz_) {
            memset(&z_, 0, sizeof(z_));
            if (inflateInit2(&z_, 15 + 32) != Z_OK) throw IoError("zlib init failed");
            z_.next_in = reinterpret_cast<Bytef *>(in_.data()), z_.avail_in = (uInt)n_in_;
        }
    }
    GzStream(const GzStream &) = delete;
    ~GzStream() {
        if (gz_) inflateEnd(&z_);
        if (f_) fclose(f_);
    }
    // up to n bytes; fewer only at end of file
    size_t read(void *dst, size_t n) {
        size_t got = 0;
        while (got < n) {
            if (pos_ == len_ && !fill()) break;
            const size_t take = std::min(n - got, len_ - pos_);
            memcpy(static_cast<char *>(dst) + got, out_.data() + pos_, take);
            pos_ += take, got += take;
        }
        return got;
    }
    void read_exact(void *dst, size_t n, const char *what) {
        if (read(dst, n) != n) throw IoError(path_ + ": truncated " + what);
    }
    bool getline(std::string &line) { // strips "\n" / "\r\n"; false at end of file with nothing read
        line.clear();
        bool any = false;
        for (;;) {
            if (pos_ == len_ && !fill()) break;
            any = true;
            const char *b = out_.data() + pos_;
            const char *nl = static_cast<const char *>(memchr(b, '\n', len_ - pos_));
            if (nl) {
                line.append(b, nl - b);
                pos_ += (size_t)(nl - b) + 1;
                break;
            }
            line.append(b, len_ - pos_);
            pos_ = len_;
        }
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return any;
    }

  private:
    bool fill() {
        pos_ = len_ = 0;
        if (!gz_) {
            if (n_in_) { // the probe block read by the constructor
                out_.assign(in_.begin(), in_.begin() + (long)n_in_);
                len_ = n_in_, n_in_ = 0;
                out_.resize(std::max<size_t>(out_.size(), 1 << 17));
                return len_ > 0;
            }
            len_ = fread(out_.data(), 1, out_.size(), f_);
            return len_ > 0;
        }
        while (len_ == 0) {
            if (z_.avail_in == 0) {
                n_in_ = fread(in_.data(), 1, in_.size(), f_);
                if (n_in_ == 0) return false;
                z_.next_in = reinterpret_cast<Bytef *>(in_.data()), z_.avail_in = (uInt)n_in_;
            }
            z_.next_out = reinterpret_cast<Bytef *>(out_.data()), z_.avail_out = (uInt)out_.size();
            const int r = inflate(&z_, Z_NO_FLUSH);
            len_ = out_.size() - z_.avail_out;
            if (r == Z_STREAM_END) {
                if (inflateReset(&z_) != Z_OK) throw IoError(path_ + ": zlib reset failed");
            } else if (r != Z_OK && r != Z_BUF_ERROR) {
                throw IoError(path_ + ": corrupt gzip data");
            }
        }
        return true;
    }
    std::string path_;
    FILE *f_ = nullptr;
    bool gz_ = false;
    z_stream z_;
    std::vector<char> in_, out_;
    size_t n_in_ = 0, pos_ = 0, len_ = 0;
};

inline std::vector<std::string> split(const std::string &s, char d) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find(d, a);
        out.push_back(s.substr(a, b == std::string::npos ? b : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}
inline std::vector<std::string> split_whitespace(const std::string &s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < s.size() && !isspace((unsigned char)s[j])) ++j;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

// Indexed reference: <fasta> + <fasta>.fai (samtools faidx), plain FASTA.  src/readers.rs:14-25.
class Fasta {
  public:
    explicit Fasta(const std::string &path) : path_(path) {
        std::ifstream fai(path + ".fai");
        if (!fai) throw IoError("Reference index file not found: " + path + ".fai. Create it using 'samtools faidx " + path + "'");
        std::string line;
        while (std::getline(fai, line)) {
            const auto f = split(line, '\t');
            if (f.size() < 5) continue;
            Entry e{std::stoull(f[1]), std::stoull(f[2]), std::stoull(f[3]), std::stoull(f[4])};
            index_[f[0]] = e;
        }
        f_ = fopen(path.c_str(), "rb");
        if (!f_) throw IoError("cannot open " + path);
        unsigned char magic[2] = {0, 0};
        if (fread(magic, 1, 2, f_) == 2 && magic[0] == 0x1f && magic[1] == 0x8b)
            throw IoError(path + ": compressed reference FASTA is not supported, decompress it first");
    }
    Fasta(const Fasta &) = delete;
    ~Fasta() {
        if (f_) fclose(f_);
    }
    bool has(const std::string &contig) const { return index_.count(contig) != 0; }
    // bases [start, end] (0-based, inclusive) like faidx::Reader::fetch_seq, clipped to the contig
    std::string fetch(const std::string &contig, uint64_t start, uint64_t end) const {
        const auto it = index_.find(contig);
        if (it == index_.end()) throw IoError("contig " + contig + " not in reference");
        const Entry &e = it->second;
        if (start >= e.len) return "";
        end = std::min(end, e.len - 1);
        const uint64_t b0 = e.offset + start / e.linebases * e.linewidth + start % e.linebases;
        const uint64_t b1 = e.offset + end / e.linebases * e.linewidth + end % e.linebases;
        std::string raw((size_t)(b1 - b0 + 1), '\0');
        if (fseeko(f_, (off_t)b0, SEEK_SET) != 0 || fread(&raw[0], 1, raw.size(), f_) != raw.size())
            throw IoError(path_ + ": short read");
        std::string seq;
        seq.reserve((size_t)(end - start + 1));
        for (char c : raw)
            if (c != '\n' && c != '\r') seq.push_back(c);
        return seq;
    }

  private:
    struct Entry {
        uint64_t len, offset, linebases, linewidth;
    };
    std::string path_;
    FILE *f_ = nullptr;
    std::unordered_map<std::string, Entry> index_;
};

struct LocusRec { // src/locus.rs:18-29 with the flanks the hot path concatenates around every allele
    Locus locus;
    std::string left_flank, right_flank;
};

// src/locus.rs:88-140.  Throws IoError with the reference's messages.
// src/locus.rs:249-260: "name=value"; anything without '=' is an error (the reference's test at :266-277)
inline std::pair<std::string, std::string> decode_info_field(const std::string &enc) {
    const size_t eq = enc.find('=');
    if (eq == std::string::npos) throw IoError("Invalid entry: " + enc);
    return {enc.substr(0, eq), enc.substr(eq + 1)};
}

// src/allele.rs:406-426: genotype indices sorted (phased 1|0 == 0|1), every allele = left flank + VCF allele + right
// flank; a homozygous genotype gives two identical targets (the reference's tests at :529-580)
inline std::pair<std::vector<std::string>, std::vector<size_t>> build_allele_seqs(const std::vector<std::string> &vcf_allele_seqs,
                                                                                 std::vector<size_t> genotype_indices,
                                                                                 const std::string &left_flank,
                                                                                 const std::string &right_flank) {
    std::sort(genotype_indices.begin(), genotype_indices.end());
    std::vector<std::string> alleles;
    for (size_t g : genotype_indices) alleles.push_back(left_flank + vcf_allele_seqs.at(g) + right_flank);
    return {alleles, genotype_indices};
}

inline LocusRec parse_catalog_line(const Fasta &genome, const std::string &line, size_t flank_len) {
    const auto f = split_whitespace(line);
    if (f.size() != 4)
        throw IoError("Expected 4 fields in the format 'chrom start end info', found " + std::to_string(f.size()) + ": " + line);
    if (!genome.has(f[0])) throw IoError("Chromosome " + f[0] + " not found in reference");
    auto parse_u32 = [](const std::string &s, const char *what) -> uint64_t {
        if (s.empty() || s.size() > 10 || s.find_first_not_of("0123456789") != std::string::npos || std::stoull(s) > 0xffffffffull)
            throw IoError(std::string("Invalid ") + what + " position: " + s);
        return std::stoull(s);
    };
    const uint64_t start = parse_u32(f[1], "start"), end = parse_u32(f[2], "end");
    if (start >= end) throw IoError("Invalid region: start " + f[1] + " >= end " + f[2]);
    LocusRec r;
    r.locus.chrom = f[0], r.locus.start = (long)start, r.locus.end = (long)end;
    bool have_id = false, have_motifs = false;
    std::vector<std::string> seen;
    for (const std::string &enc : split(f[3], ';')) {
        const auto [name, value] = decode_info_field(enc);
        if (std::find(seen.begin(), seen.end(), name) != seen.end()) throw IoError("Duplicate field: " + name);
        seen.push_back(name);
        if (name == "ID") r.locus.id = value, have_id = true;
        if (name == "MOTIFS") r.locus.motifs = split(value, ','), have_motifs = true;
    }
    if (!have_id) throw IoError("ID field missing");
    if (!have_motifs) throw IoError("MOTIFS field missing");
    if (start < flank_len) throw IoError("Error fetching sequence for region " + f[0] + ": flank reaches before the contig start");
    std::string full = genome.fetch(f[0], start - flank_len, end + flank_len - 1); // src/locus.rs:222-247
    if (full.size() < 2 * flank_len) throw IoError("Error fetching sequence for region " + f[0] + ":" + f[1] + "-" + f[2]);
    for (char &c : full) c = (char)toupper((unsigned char)c);
    r.left_flank = full.substr(0, flank_len), r.right_flank = full.substr(full.size() - flank_len);
    return r;
}

// src/locus.rs:185-220: every line of the catalog; a line that fails is reported and skipped.
inline std::vector<LocusRec> read_catalog(const std::string &bed_path, const Fasta &genome, size_t flank_len,
                                          std::vector<std::string> *errors = nullptr) {
    GzStream in(bed_path);
    std::vector<LocusRec> out;
    std::string line;
    size_t n = 0;
    while (in.getline(line)) {
        ++n;
        try {
            out.push_back(parse_catalog_line(genome, line, flank_len));
        } catch (const IoError &e) {
            if (errors) errors->push_back("Error at BED line " + std::to_string(n) + ": " + e.what());
        }
    }
    return out;
}

// One sample: its TRGT VCF and spanning BAM, streamed once and keyed by TRID.
class SampleFiles {
  public:
    std::string karyotype = "XY", trgt_version;
    bool is_trgt_v1 = false;
    size_t reads_without_tr = 0;

    // src/handles.rs:64-79: <prefix>.spanning.sorted.bam + <prefix>.sorted.vcf.gz
    static SampleFiles from_prefix(const std::string &prefix) {
        return SampleFiles(prefix + ".sorted.vcf.gz", prefix + ".spanning.sorted.bam");
    }
    SampleFiles(const std::string &vcf_path, const std::string &bam_path) {
        std::string vcf_version;
        read_vcf(vcf_path, vcf_version);
        read_bam(bam_path);
        if (trgt_version != vcf_version) // src/handles.rs:150-157
            throw IoError("TRGT version mismatch: BAM version is '" + trgt_version + "', VCF version is '" + vcf_version + "'");
        const std::string major = trgt_version.substr(0, trgt_version.find('.'));
        is_trgt_v1 = !major.empty() && major.size() <= 3 && major.find_first_not_of("0123456789") == std::string::npos &&
                     std::stoi(major) > 0 && std::stoi(major) < 256;
    }

    // get_vcf_data + get_reads for one locus (src/allele.rs:274-426); nullopt where the reference returns Err
    // (TRID absent from the VCF, missing genotype, malformed MC / AL, no reads), with the reason in *why.
    // `take` moves the reads out of this object (a genome-sized shard should not hold its reads twice); a second call
    // for the same TRID then finds none.
    std::optional<SampleInput> sample_input(const LocusRec &l, std::string *why = nullptr, bool take = false) {
        auto fail = [&](const std::string &m) {
            if (why) *why = m;
            return std::nullopt;
        };
        const std::string &id = l.locus.id;
        const auto v = vcf_.find(id);
        if (v == vcf_.end()) return fail("TRID=" + id + " missing");
        const VcfRec &rec = v->second;
        if (rec.gt.empty() || rec.gt[0] < 0) return fail("Missing genotype for TRID=" + id);
        if (rec.mc.empty()) return fail("Missing or malformed MC field for TRID=" + id);
        if (rec.al.empty()) return fail("Missing or malformed AL field for TRID=" + id);
        SampleInput s;
        s.karyotype = karyotype;
        for (int g : rec.gt)
            if (g >= 0) s.genotype_indices.push_back((size_t)g);
        const size_t skip = is_trgt_v1 ? 1 : 0; // TRGT >= 1.0 alleles carry a padding base (src/allele.rs:386-390)
        std::vector<std::string> vcf_alleles;
        for (const std::string &a : rec.alleles) vcf_alleles.push_back(a.size() > skip ? a.substr(skip) : std::string());
        for (size_t g : s.genotype_indices)
            if (g >= rec.alleles.size()) return fail("genotype index beyond the alleles of TRID=" + id);
        std::tie(s.allele_seqs, s.genotype_indices) = build_allele_seqs(vcf_alleles, s.genotype_indices, l.left_flank, l.right_flank);
        s.motif_counts = rec.mc, s.allele_lengths = rec.al;
        const auto r = reads_.find(id);
        if (r == reads_.end() || r->second.empty()) return fail("No reads found for locus: " + id);
        if (take) s.reads = std::move(r->second), r->second.clear();
        else s.reads = r->second;
        return s;
    }

  private:
    struct VcfRec {
        std::vector<std::string> alleles; // REF, ALT...
        std::vector<int> gt;              // allele indices of the first sample, -1 = missing
        std::vector<std::string> mc, al;
    };
    std::unordered_map<std::string, VcfRec> vcf_;
    std::unordered_map<std::string, std::vector<TrgtRead>> reads_;

    void read_vcf(const std::string &path, std::string &version) {
        GzStream in(path);
        std::string line;
        while (in.getline(line)) {
            if (line.empty()) continue;
            if (line[0] == '#') {
                if (line.rfind("##trgtVersion=", 0) == 0) version = line.substr(14); // last one wins, src/handles.rs:96-107
                continue;
            }
            const auto f = split(line, '\t');
            if (f.size() < 10) continue;
            std::string trid;
            for (const std::string &kv : split(f[7], ';'))
                if (kv.rfind("TRID=", 0) == 0) trid = kv.substr(5);
            if (trid.empty() || vcf_.count(trid)) continue; // first record of a TRID wins (first match of the region query)
            VcfRec rec;
            rec.alleles.push_back(f[3]);
            if (f[4] != ".")
                for (const std::string &a : split(f[4], ',')) rec.alleles.push_back(a);
            const auto keys = split(f[8], ':'), vals = split(f[9], ':');
            for (size_t k = 0; k < keys.size() && k < vals.size(); ++k) {
                if (keys[k] == "GT") {
                    std::string tok;
                    for (size_t i = 0; i <= vals[k].size(); ++i) {
                        if (i == vals[k].size() || vals[k][i] == '/' || vals[k][i] == '|') {
                            const bool num = !tok.empty() && tok.size() <= 6 && tok.find_first_not_of("0123456789") == std::string::npos;
                            rec.gt.push_back(num ? std::stoi(tok) : -1); // "." or anything unreadable: a missing allele
                            tok.clear();
                        } else {
                            tok.push_back(vals[k][i]);
                        }
                    }
                } else if (keys[k] == "MC") {
                    if (vals[k] != ".") rec.mc = split(vals[k], ',');
                } else if (keys[k] == "AL") {
                    if (vals[k] != ".") rec.al = split(vals[k], ',');
                }
            }
            vcf_.emplace(std::move(trid), std::move(rec));
        }
    }

    static uint32_t le32(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
    static uint16_t le16(const unsigned char *p) { return (uint16_t)(p[0] | p[1] << 8); }

    // @PG line of TRGT: "ID:trgt" or "PN:trgt" among its fields; returns the field starting with `tag` (src/handles.rs:81-94)
    static std::string pg_field(const std::string &header, const char *tag) {
        for (const std::string &line : split(header, '\n')) {
            if (line.rfind("@PG", 0) != 0) continue;
            const auto f = split(line, '\t');
            if (std::find(f.begin(), f.end(), "ID:trgt") == f.end() && std::find(f.begin(), f.end(), "PN:trgt") == f.end()) continue;
            for (const std::string &x : f)
                if (x.rfind(tag, 0) == 0) return x.substr(3);
        }
        return "";
    }

    void read_bam(const std::string &path) {
        GzStream in(path);
        unsigned char b4[4];
        in.read_exact(b4, 4, "BAM magic");
        if (memcmp(b4, "BAM\1", 4) != 0) throw IoError(path + ": not a BAM file");
        in.read_exact(b4, 4, "BAM header");
        if (le32(b4) > (1u << 30)) throw IoError(path + ": implausible BAM header length");
        std::string text(le32(b4), '\0');
        if (!text.empty()) in.read_exact(&text[0], text.size(), "BAM header text");
        trgt_version = pg_field(text, "VN:");
        const std::string cl = pg_field(text, "CL:");
        const char *key = "--karyotype ";
        const size_t at = cl.find(key);
        std::string kary = "XY"; // src/handles.rs:109-121
        if (at != std::string::npos) {
            const auto rest = split_whitespace(cl.substr(at + strlen(key)));
            if (!rest.empty()) kary = rest[0];
        }
        karyotype = kary == "XX" ? "XX" : "XY"; // Karyotype::new: anything else is treated as XY
        in.read_exact(b4, 4, "reference count");
        for (uint32_t i = 0, n = le32(b4); i < n; ++i) {
            in.read_exact(b4, 4, "reference name length");
            if (le32(b4) > (1u << 16)) throw IoError(path + ": implausible reference name length");
            std::string name(le32(b4), '\0');
            if (!name.empty()) in.read_exact(&name[0], name.size(), "reference name");
            in.read_exact(b4, 4, "reference length");
        }
        std::vector<unsigned char> rec;
        for (;;) {
            if (in.read(b4, 4) != 4) break;
            const uint32_t bs = le32(b4);
            if (bs < 32 || bs > (1u << 28)) throw IoError(path + ": corrupt BAM record");
            rec.resize(bs);
            in.read_exact(rec.data(), bs, "BAM record");
            const unsigned char *p = rec.data();
            const uint32_t l_name = p[8], n_cigar = le16(p + 12), l_seq = le32(p + 16);
            size_t off = 32;
            if (off + l_name + 4ull * n_cigar + (l_seq + 1) / 2 + l_seq > bs) throw IoError(path + ": corrupt BAM record");
            TrgtRead r;
            r.name.assign(reinterpret_cast<const char *>(p + off), l_name ? l_name - 1 : 0);
            off += l_name + 4ull * n_cigar;
            // packed at the source: BAM's 4-bit bases go straight to 2-bit codes, ASCII (src/read.rs:57) is never written
            r.packed = PackedBases::from_bam(p + off, l_seq);
            off += (l_seq + 1) / 2 + l_seq;
            std::string tr;
            bool have_tr = false;
            while (off + 3 <= bs) { // auxiliary fields
                const char t0 = (char)p[off], t1 = (char)p[off + 1], ty = (char)p[off + 2];
                off += 3;
                auto is = [&](const char *tag) { return t0 == tag[0] && t1 == tag[1]; };
                size_t sz = 0;
                switch (ty) {
                case 'A': case 'c': case 'C': sz = 1; break;
                case 's': case 'S': sz = 2; break;
                case 'i': case 'I': case 'f': sz = 4; break;
                case 'Z': case 'H': {
                    const void *z = memchr(p + off, 0, bs - off);
                    if (!z) throw IoError(path + ": unterminated tag string");
                    const size_t n = (size_t)(static_cast<const unsigned char *>(z) - (p + off));
                    if (ty == 'Z' && is("TR")) tr.assign(reinterpret_cast<const char *>(p + off), n), have_tr = true;
                    sz = n + 1;
                    break;
                }
                case 'B': {
                    if (off + 5 > bs) throw IoError(path + ": corrupt tag array");
                    const char sub = (char)p[off];
                    const size_t el = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    sz = 5 + el * (size_t)le32(p + off + 1);
                    break;
                }
                default: throw IoError(path + ": unknown tag type");
                }
                if (off + sz > bs) throw IoError(path + ": corrupt tag");
                // src/read.rs:62-72: AL / HP are taken from a u8 ('C') or an i32 ('i') value only
                if (is("AL") || is("HP")) {
                    std::optional<int> v;
                    if (ty == 'C') v = (int)p[off];
                    else if (ty == 'i') v = (int)(uint8_t)(int32_t)le32(p + off);
                    if (is("AL")) r.classification = v; else r.haplotype = v;
                }
                off += sz;
            }
            if (!have_tr) {
                ++reads_without_tr;
                continue;
            }
            reads_[tr].push_back(std::move(r));
        }
    }
};

// The samples' files, each pair read on its own thread (they are independent and decompression is the cost).
inline std::vector<std::unique_ptr<SampleFiles>> open_samples(const std::vector<std::string> &prefixes) {
    std::vector<std::unique_ptr<SampleFiles>> out(prefixes.size());
    std::vector<std::string> err(prefixes.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < prefixes.size(); ++i)
        th.emplace_back([&, i]() {
            try {
                out[i].reset(new SampleFiles(SampleFiles::from_prefix(prefixes[i])));
            } catch (const std::exception &e) {
                err[i] = e.what();
            }
        });
    for (auto &t : th) t.join();
    for (const std::string &e : err)
        if (!e.empty()) throw IoError(e);
    return out;
}

// Every locus of the catalog with the inputs of its samples, ready for process_alleles.
struct FamilyInputs {
    std::vector<Locus> loci;
    std::vector<std::vector<std::optional<SampleInput>>> samples; // [sample][locus]
    std::vector<std::string> messages;                             // per-sample reasons for missing inputs
};
inline FamilyInputs gather(const std::vector<LocusRec> &catalog, const std::vector<SampleFiles *> &files, bool take = true) {
    FamilyInputs out;
    out.samples.resize(files.size());
    for (const LocusRec &l : catalog) {
        out.loci.push_back(l.locus);
        for (size_t s = 0; s < files.size(); ++s) {
            std::string why;
            out.samples[s].push_back(files[s]->sample_input(l, &why, take));
            if (!why.empty()) out.messages.push_back(l.locus.id + " sample " + std::to_string(s) + ": " + why);
        }
    }
    return out;
}

} // namespace io
} // namespace tdbh
