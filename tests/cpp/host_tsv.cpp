// Test driver for the C++ host side (include/trgt_denovo_b200.hpp): runs the per-locus pipeline of
// src/trio/allele.rs / src/duo/allele.rs on a family description and writes the TSV + read-id file, or
// aligns one pair through the WFAligner shim.  Built and driven by tests/test_cpp_host.py.
//
//   host_tsv trio|duo <family.txt> <out.tsv> <out_read_ids.txt> [n_gpus]
//                                                 n_gpus > 0: through tdbh::MultiGpu (one ctx + host thread per GPU,
//                                                 loci in balanced contiguous ranges, rows merged in locus order)
//   host_tsv align <pattern> <text> <clip>        -> "status score cigar clipped clipped_cigar"
//   host_tsv pack <in.txt> <out.txt>              packing at the source, no device needed: lines "A ascii" (ASCII
//                                                 appended to the batch), "P ascii" (PackedBases::from_ascii),
//                                                 "N hex l_seq" (PackedBases::from_bam: BAM 4-bit SEQ bytes),
//                                                 "B" (start a new shard; shards are merged with Batch::append)
//                                                 -> n_bases / words / lengths / exception indices / exception bytes
//
// family.txt (tab separated):
//   P clip_len parent_quantile partition_by_alignment quick_kind(-|AL|MC) quick_tol(-|float)
//   L chrom start end trid motif[,motif...]                     one per locus, followed by its samples
//   S role(F|M|C|A|B) present(0|1) karyotype
//   A seq motif_count allele_length genotype_index              alleles of the sample
//   R name bases classification(-1 = none) haplotype(0 = none)  reads of the sample
#include "trgt_denovo_b200.hpp"

#include <fstream>
#include <iostream>
#include <sstream>

using namespace tdbh;

static std::vector<std::string> split(const std::string &s, char d) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream is(s);
    while (std::getline(is, cur, d)) out.push_back(cur);
    if (!s.empty() && s.back() == d) out.push_back("");
    return out;
}

int main(int argc, char **argv) {
    try {
        if (argc >= 5 && std::string(argv[1]) == "align") {
            WFAligner a = create_aligner_with_scoring();
            const auto st = a.align_end_to_end(argv[2], argv[3]);
            const size_t clip = (size_t)std::stoul(argv[4]);
            std::cout << (int)st << " " << a.score() << " " << a.cigar() << " " << a.cigar_score_clipped(clip) << " "
                      << a.cigar(clip) << "\n";
            return 0;
        }
        if (argc >= 4 && std::string(argv[1]) == "pack") {
            std::ifstream in(argv[2]);
            std::vector<detail::Batch> shards(1);
            std::string line;
            while (std::getline(in, line)) {
                const auto f = split(line, '\t');
                if (f.empty()) continue;
                detail::Batch &b = shards.back();
                if (f[0] == "B") shards.emplace_back();
                else if (f[0] == "A") b.seq(f.size() > 1 ? f[1] : std::string());
                else if (f[0] == "P") b.seq(PackedBases::from_ascii(f.size() > 1 ? f[1] : std::string()));
                else if (f[0] == "N") {
                    std::vector<uint8_t> raw;
                    for (size_t i = 0; i + 1 < f[1].size(); i += 2) raw.push_back((uint8_t)std::stoul(f[1].substr(i, 2), nullptr, 16));
                    const PackedBases pb = PackedBases::from_bam(raw.data(), std::stoul(f[2]));
                    if (PackedBases::from_ascii(pb.ascii()).w != pb.w) throw Error(TDB_E_INVALID, "from_bam / ascii round trip");
                    b.seq(pb);
                }
            }
            detail::Batch all;
            for (const auto &sh : shards) all.append(sh);
            const tdb_packed_seqs v = all.view();
            std::ofstream o(argv[3]);
            o << v.n_bases << "\n";
            for (size_t i = 0; i < tdb_packed_words(v.n_bases); ++i) o << std::hex << v.packed[i] << (i + 1 < tdb_packed_words(v.n_bases) ? " " : "");
            o << std::dec << "\n";
            for (size_t i = 0; i < v.n_seqs; ++i) o << v.seq_len[i] << " ";
            o << "\n";
            for (size_t i = 0; i < v.n_exc; ++i) o << v.exc_seq[i] << " ";
            o << "\n" << all.exc_bytes << "\n";
            return 0;
        }
        if (argc < 5) {
            std::cerr << "usage: host_tsv trio|duo family.txt out.tsv out_ids.txt | host_tsv align P T clip\n";
            return 2;
        }
        const bool is_duo = std::string(argv[1]) == "duo";
        std::ifstream in(argv[2]);
        Params params;
        std::vector<Locus> loci;
        std::vector<std::optional<SampleInput>> fam[3]; // trio: F M C ; duo: A B -
        std::optional<SampleInput> *cur = nullptr;
        std::string line;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            const auto f = split(line, '\t');
            if (f[0] == "P") {
                params.clip_len = std::stoul(f[1]), params.parent_quantile = std::stod(f[2]);
                params.partition_by_alignment = f[3] == "1";
                if (f[4] != "-") {
                    QuickMode q;
                    q.kind = f[4] == "AL" ? QuickMode::AL : QuickMode::MC;
                    if (f[5] != "-") q.tolerance = std::stod(f[5]);
                    params.quick_mode = q;
                }
            } else if (f[0] == "L") {
                Locus l;
                l.chrom = f[1], l.start = std::stol(f[2]), l.end = std::stol(f[3]), l.id = f[4], l.motifs = split(f[5], ',');
                loci.push_back(l);
            } else if (f[0] == "S") {
                const int slot = f[1] == "F" || f[1] == "A" ? 0 : (f[1] == "M" || f[1] == "B" ? 1 : 2);
                fam[slot].emplace_back(std::nullopt);
                cur = &fam[slot].back();
                if (f[2] == "1") {
                    *cur = SampleInput();
                    (*cur)->karyotype = f[3];
                }
            } else if (f[0] == "A") {
                (*cur)->allele_seqs.push_back(f[1]), (*cur)->motif_counts.push_back(f[2]);
                (*cur)->allele_lengths.push_back(f[3]), (*cur)->genotype_indices.push_back(std::stoul(f[4]));
            } else if (f[0] == "R") {
                TrgtRead r;
                r.name = f[1];
                // every other read arrives packed at the source, the rest as ASCII: both forms go through one batch
                if ((*cur)->reads.size() % 2) r.packed = PackedBases::from_ascii(f[2]);
                else r.bases = f[2];
                if (f[3] != "-1") r.classification = std::stoi(f[3]);
                if (f[4] != "0") r.haplotype = std::stoi(f[4]);
                (*cur)->reads.push_back(r);
            }
        }
        std::string tsv, ids;
        const int n_gpus = argc >= 6 ? std::stoi(argv[5]) : 0;
        if (n_gpus > 0) {
            MultiGpu mg(n_gpus);
            const auto rows = is_duo ? mg.duo_process_alleles(loci, fam[0], fam[1], params)
                                     : mg.trio_process_alleles(loci, fam[0], fam[1], fam[2], params);
            tsv = is_duo ? write_tsv(rows, duo::columns(), duo::row_tsv) : write_tsv(rows, trio::columns(), trio::row_tsv);
            ids = write_read_ids(rows);
            std::ofstream(argv[3]) << tsv;
            std::ofstream(argv[4]) << ids;
            std::cout << "bounds";
            for (size_t b : mg.last_bounds()) std::cout << " " << b;
            std::cout << "\n";
            return 0;
        }
        WFAligner aligner = create_aligner_with_scoring();
        if (is_duo) {
            const auto rows = duo::process_alleles(aligner, loci, fam[0], fam[1], params);
            tsv = write_tsv(rows, duo::columns(), duo::row_tsv), ids = write_read_ids(rows);
        } else {
            const auto rows = trio::process_alleles(aligner, loci, fam[0], fam[1], fam[2], params);
            tsv = write_tsv(rows, trio::columns(), trio::row_tsv), ids = write_read_ids(rows);
        }
        std::ofstream(argv[3]) << tsv;
        std::ofstream(argv[4]) << ids;
        return 0;
    } catch (const Error &e) {
        std::cerr << e.what() << "\n";
        return e.code == TDB_E_NO_DEVICE ? 42 : 1;
    }
}
