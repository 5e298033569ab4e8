// Test driver for the real-input front end (include/trgt_denovo_b200_io.hpp): catalog BED + indexed FASTA + per
// sample TRGT VCF / spanning BAM -> the per-locus pipeline of the C++ host side.  Built and driven by
// tests/test_io_frontend.py.
//
//   files_tsv dump <ref.fa> <catalog.bed> <flank_len> <out_family.txt> <prefix>...     (no GPU needed)
//   files_tsv trio <ref.fa> <catalog.bed> <out.tsv> <out_read_ids.txt> <father> <mother> <child>
//   files_tsv duo  <ref.fa> <catalog.bed> <out.tsv> <out_read_ids.txt> <sampleA> <sampleB>
//
// `dump` writes what the readers produced in the family.txt format of host_tsv.cpp (roles F, M, C in
// the order of the prefixes), followed by "E <message>" lines for catalog errors and missing inputs.
#include "trgt_denovo_b200_io.hpp"

#include <chrono>
#include <iostream>
#include <memory>

using namespace tdbh;

int main(int argc, char **argv) {
    try {
        if (argc < 7) {
            std::cerr << "usage: files_tsv dump ref.fa catalog.bed flank out.txt prefix... | files_tsv trio|duo ref.fa catalog.bed out.tsv out_ids.txt prefixes...\n";
            return 2;
        }
        const std::string mode = argv[1];
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto t0 = now();
        io::Fasta genome(argv[2]);
        std::vector<std::string> errors;
        const size_t flank = mode == "dump" ? std::stoul(argv[4]) : 50;
        const auto catalog = io::read_catalog(argv[3], genome, flank, &errors);
        const auto t1 = now();
        std::vector<std::string> prefixes(argv + 6, argv + argc);
        const auto files = io::open_samples(prefixes);
        std::vector<io::SampleFiles *> ptrs;
        for (const auto &f : files) ptrs.push_back(f.get());
        const auto t2 = now();
        io::FamilyInputs fam = io::gather(catalog, ptrs);
        const auto t3 = now();
        if (mode == "dump") {
            std::ofstream o(argv[5]);
            o << "P\t50\t1.0\t0\t-\t-\n";
            const char *roles = "FMC";
            for (size_t i = 0; i < fam.loci.size(); ++i) {
                const Locus &l = fam.loci[i];
                o << "L\t" << l.chrom << "\t" << l.start << "\t" << l.end << "\t" << l.id << "\t";
                for (size_t k = 0; k < l.motifs.size(); ++k) o << (k ? "," : "") << l.motifs[k];
                o << "\n";
                for (size_t s = 0; s < fam.samples.size(); ++s) {
                    const auto &in = fam.samples[s][i];
                    o << "S\t" << roles[s % 3] << "\t" << (in ? 1 : 0) << "\t" << (in ? in->karyotype : "XX") << "\n";
                    if (!in) continue;
                    for (size_t k = 0; k < in->allele_seqs.size(); ++k)
                        o << "A\t" << in->allele_seqs[k] << "\t" << in->motif_counts[k] << "\t" << in->allele_lengths[k] << "\t"
                          << in->genotype_indices[k] << "\n";
                    for (const TrgtRead &r : in->reads)
                        o << "R\t" << r.name << "\t" << r.bases_ascii() << "\t" << (r.classification ? *r.classification : -1) << "\t"
                          << (r.haplotype ? *r.haplotype : 0) << "\n";
                }
            }
            for (const std::string &e : errors) o << "E\t" << e << "\n";
            for (const std::string &e : fam.messages) o << "E\t" << e << "\n";
            for (size_t s = 0; s < files.size(); ++s)
                o << "V\t" << s << "\t" << files[s]->trgt_version << "\t" << (files[s]->is_trgt_v1 ? 1 : 0) << "\t" << files[s]->karyotype
                  << "\t" << files[s]->reads_without_tr << "\n";
            return 0;
        }
        Params params;
        WFAligner aligner = create_aligner_with_scoring();
        const auto t4 = now();
        std::string tsv, ids;
        if (mode == "duo") {
            if (fam.samples.size() != 2) throw io::IoError("duo needs two sample prefixes");
            const auto rows = duo::process_alleles(aligner, fam.loci, std::move(fam.samples[0]), std::move(fam.samples[1]), params);
            tsv = write_tsv(rows, duo::columns(), duo::row_tsv), ids = write_read_ids(rows);
        } else {
            if (fam.samples.size() != 3) throw io::IoError("trio needs father, mother and child prefixes");
            const auto rows = trio::process_alleles(aligner, fam.loci, std::move(fam.samples[0]), std::move(fam.samples[1]), std::move(fam.samples[2]), params);
            tsv = write_tsv(rows, trio::columns(), trio::row_tsv), ids = write_read_ids(rows);
        }
        const auto t5 = now();
        std::ofstream(argv[4]) << tsv;
        std::ofstream(argv[5]) << ids;
        std::cerr << "timing_ms loci=" << fam.loci.size() << " catalog+flanks=" << ms(t0, t1) << " vcf+bam=" << ms(t1, t2)
                  << " gather=" << ms(t2, t3) << " context=" << ms(t3, t4) << " process_alleles=" << ms(t4, t5)
                  << " write=" << ms(t5, now()) << "\n";
        return 0;
    } catch (const Error &e) {
        std::cerr << e.what() << "\n";
        return e.code == TDB_E_NO_DEVICE ? 42 : 1;
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
}
