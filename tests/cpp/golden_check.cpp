// Evaluates the host-side functions the reference's own unit tests cover, on inputs read from stdin (one case per
// line, tab separated), so that tests/test_host_golden.py can assert the reference's expected values
// (tests/golden/host_vectors.json) against the C++ header.  Needs no device.
//   T fn(equiv|sim) kind(AL|MC) tol(-|float) father mother child     alleles as "mc:al,mc:al"  -> 0|1
//   D fn(equiv|sim) kind tol a b                                                             -> 0|1
//   B vcf_alleles(comma) genotype_indices(comma) left_flank right_flank                       -> alleles(comma) TAB indices(comma)
//   I encoding                                                                               -> name TAB value | ERR
#include "trgt_denovo_b200_io.hpp"

#include <iostream>
#include <sstream>

using namespace tdbh;

static std::vector<std::string> split_on(const std::string &s, char d) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream is(s);
    while (std::getline(is, cur, d)) out.push_back(cur);
    return out;
}
static AlleleSet make_set(const std::string &enc) {
    AlleleSet s;
    for (const std::string &a : split_on(enc, ',')) {
        const auto f = split_on(a, ':');
        Allele x;
        x.motif_count = f[0], x.allele_length = f[1];
        s.alleles.push_back(x);
    }
    return s;
}
static QuickMode make_mode(const std::string &kind, const std::string &tol) {
    QuickMode q;
    q.kind = kind == "AL" ? QuickMode::AL : QuickMode::MC;
    if (tol != "-") q.tolerance = std::stod(tol);
    return q;
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        const auto f = split_on(line, '\t');
        if (f.empty()) continue;
        if (f[0] == "T") {
            const QuickMode q = make_mode(f[2], f[3]);
            const AlleleSet fa = make_set(f[4]), mo = make_set(f[5]), ch = make_set(f[6]);
            std::cout << (f[1] == "equiv" ? trio::check_field_equivalence(fa, mo, ch, q) : trio::check_field_similarity(fa, mo, ch, q)) << "\n";
        } else if (f[0] == "D") {
            const QuickMode q = make_mode(f[2], f[3]);
            const AlleleSet a = make_set(f[4]), b = make_set(f[5]);
            std::cout << (f[1] == "equiv" ? duo::check_allele_field_equivalence(a, b, q) : duo::check_field_similarity(a, b, q)) << "\n";
        } else if (f[0] == "B") {
            std::vector<size_t> gi;
            for (const std::string &x : split_on(f[2], ',')) gi.push_back(std::stoul(x));
            const auto [alleles, idx] = io::build_allele_seqs(split_on(f[1], ','), gi, f[3], f[4]);
            for (size_t i = 0; i < alleles.size(); ++i) std::cout << (i ? "," : "") << alleles[i];
            std::cout << "\t";
            for (size_t i = 0; i < idx.size(); ++i) std::cout << (i ? "," : "") << idx[i];
            std::cout << "\n";
        } else if (f[0] == "I") {
            try {
                const auto [name, value] = io::decode_info_field(f[1]);
                std::cout << name << "\t" << value << "\n";
            } catch (const std::exception &) {
                std::cout << "ERR\n";
            }
        }
    }
    return 0;
}
