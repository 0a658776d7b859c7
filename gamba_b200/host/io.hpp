// Input grammar and byte-exact output layout of `gamba -i/-o`.
// Behaviour follows source/io.cpp:46-373 (reader) and :375-449 (writer) of the
// reference; see SURVEY.md §3.1 and Appendix A.1/A.2. Independent implementation:
// coefficients are reduced mod p on the fly (num * den^-1), which differs from the
// reference's "clear denominators" only by a non-zero scalar per polynomial and the
// driver makes every polynomial monic on import anyway.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace gb {

struct PolySystem {
    std::vector<std::string> vars;
    uint32_t p = 0;
    // generators: flat term lists
    std::vector<uint32_t> lens;        // terms per generator
    std::vector<uint32_t> coeffs;      // residues in [0,p)
    std::vector<uint16_t> exps;        // nvars per term

    void read(std::istream& in);
    void write(std::ostream& out) const;
};

bool is_prime_u32(uint64_t n);
uint32_t inv_mod(uint32_t a, uint32_t p);

}  // namespace gb
