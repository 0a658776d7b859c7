#include "io.hpp"

#include <atomic>
#include <charconv>
#include <thread>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <istream>
#include <numeric>
#include <ostream>
#include <sstream>
#include <stdexcept>

namespace gb {

static uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t m) {
    return (uint64_t)((unsigned __int128)a * b % m);
}
static uint64_t powmod64(uint64_t a, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m; a %= m;
    for (; e; e >>= 1, a = mulmod64(a, a, m)) if (e & 1) r = mulmod64(r, a, m);
    return r;
}
// deterministic Miller-Rabin, exact for n < 2^64 with these bases
bool is_prime_u32(uint64_t n) {
    if (n < 2) return false;
    for (uint64_t q : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % q == 0) return n == q;
    }
    uint64_t d = n - 1; int s = 0;
    while ((d & 1) == 0) { d >>= 1; ++s; }
    for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        uint64_t x = powmod64(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int r = 1; r < s; ++r) {
            x = mulmod64(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}

uint32_t inv_mod(uint32_t a, uint32_t p) {
    int64_t t = 0, nt = 1, r = p, nr = a % p;
    while (nr) {
        int64_t q = r / nr;
        int64_t x = t - q * nt; t = nt; nt = x;
        x = r - q * nr; r = nr; nr = x;
    }
    if (t < 0) t += p;
    return (uint32_t)t;
}

static std::string strip_spaces(const std::string& s) {
    std::string o; o.reserve(s.size());
    for (char c : s) if (!std::isspace((unsigned char)c)) o.push_back(c);
    return o;
}

namespace {
struct Parser {
    PolySystem& sys;
    std::vector<int> by_len;  // variable indices sorted by decreasing name length

    explicit Parser(PolySystem& s) : sys(s) {
        by_len.resize(sys.vars.size());
        std::iota(by_len.begin(), by_len.end(), 0);
        std::stable_sort(by_len.begin(), by_len.end(), [&](int a, int b) {
            return sys.vars[a].size() > sys.vars[b].size();
        });
    }
    [[noreturn]] static void bad(size_t gen) {
        throw std::runtime_error("Error parsing generator number " + std::to_string(gen) + ".");
    }
    // digits -> value mod p (p > 0)
    static bool digits_mod(const std::string& t, size_t& i, uint32_t p, uint32_t& out) {
        size_t s = i; uint64_t v = 0;
        while (i < t.size() && std::isdigit((unsigned char)t[i])) { v = (v * 10 + (t[i] - '0')) % p; ++i; }
        out = (uint32_t)v;
        return i > s;
    }
    void term(const std::string& t, size_t gen, std::vector<uint16_t>& e, uint32_t& cf) {
        const uint32_t p = sys.p;
        size_t i = 0; bool neg = false;
        if (t[i] == '+' || t[i] == '-') { neg = t[i] == '-'; ++i; }
        if (i >= t.size() || t[i] == '*' || t.back() == '*') bad(gen);
        uint32_t num = 1 % p, den = 1 % p;
        bool have_cf = false;
        if (std::isdigit((unsigned char)t[i])) {
            have_cf = true;
            digits_mod(t, i, p, num);
            if (i < t.size() && t[i] == '/') {
                ++i;
                size_t s = i;
                if (!digits_mod(t, i, p, den)) bad(gen);
                // a denominator that is 0 as an integer or divisible by p
                bool all_zero = true;
                for (size_t k = s; k < i; ++k) if (t[k] != '0') all_zero = false;
                if (all_zero || den == 0)
                    throw std::runtime_error("Division by zero in generator number " + std::to_string(gen) + ".");
            }
        } else if (t[i] == '/') bad(gen);
        cf = (uint32_t)((uint64_t)num * inv_mod(den, p) % p);
        if (neg && cf) cf = p - cf;
        if (have_cf && i < t.size()) { if (t[i] != '*') bad(gen); ++i; }
        std::fill(e.begin(), e.end(), 0);
        while (i < t.size()) {
            int found = -1;
            for (int v : by_len) {
                const std::string& nm = sys.vars[v];
                if (t.compare(i, nm.size(), nm) == 0) { found = v; break; }
            }
            if (found < 0) bad(gen);
            i += sys.vars[found].size();
            uint32_t ex = 1;
            if (i < t.size() && t[i] == '^') {
                ++i; size_t s = i; uint64_t v = 0;
                while (i < t.size() && std::isdigit((unsigned char)t[i])) { v = v * 10 + (t[i] - '0'); if (v > (1u << 20)) v = 1u << 20; ++i; }
                if (i == s) bad(gen);
                ex = (uint32_t)v;
            }
            if (ex + e[found] >= (1u << 15))
                throw std::runtime_error("Monomial exponent in generator " + std::to_string(gen)
                                         + " must be >= 0 or < 32768.");
            e[found] = (uint16_t)(e[found] + ex);
            if (i < t.size()) { if (t[i] != '*') bad(gen); ++i; if (i >= t.size()) bad(gen); }
        }
    }
    void generator(const std::string& g, size_t gen) {
        std::vector<uint16_t> e(sys.vars.size());
        uint32_t n = 0;
        size_t i = 0;
        while (i < g.size()) {
            size_t j = i + 1;
            while (j < g.size() && g[j] != '+' && g[j] != '-') ++j;
            std::string t = g.substr(i, j - i);
            if (t == "+" || t == "-") bad(gen);
            uint32_t cf;
            term(t, gen, e, cf);
            sys.coeffs.push_back(cf);
            sys.exps.insert(sys.exps.end(), e.begin(), e.end());
            ++n; i = j;
        }
        sys.lens.push_back(n);
    }
};
}  // namespace

void PolySystem::read(std::istream& in) {
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("Input file is empty.");
    line = strip_spaces(line);
    if (!line.empty() && line.back() == ',') line.pop_back();
    {
        std::stringstream ss(line); std::string nm;
        static const std::string restricted = "+-*^";
        // a trailing empty name cannot occur after removing the trailing comma
        while (std::getline(ss, nm, ',')) {
            if (nm.empty()) throw std::runtime_error("Empty variable name.");
            if (std::isdigit((unsigned char)nm[0]) || restricted.find(nm[0]) != std::string::npos)
                throw std::runtime_error("Invalid variable name: " + nm + ".");
            if (std::find(vars.begin(), vars.end(), nm) != vars.end())
                throw std::runtime_error("Duplicated variable name: " + nm + ".");
            vars.push_back(nm);
        }
        if (vars.empty()) throw std::runtime_error("Empty variable name.");
        if (vars.size() > 256) throw std::runtime_error("Too many variables (max 256).");
    }
    if (!std::getline(in, line)) throw std::runtime_error("Missing line containing field characteristic.");
    long long fc;
    try { fc = std::stoll(line); } catch (...) { throw std::runtime_error("Field characteristic not valid."); }
    if (fc < 0 || fc > 0xffffffffll) throw std::runtime_error("Field characteristic must be >= 0 and < 2^32.");
    if (fc != 0 && !is_prime_u32((uint64_t)fc)) throw std::runtime_error("Field characteristic is not a prime number.");
    p = (uint32_t)fc;
    if (p == 2) throw std::runtime_error("Field characteristic = 2 not supported.");
    if (p > 0x7fffffffu) throw std::runtime_error("Field characteristic too large.");
    if (p == 0) return;  // rationals: not implemented in the reference either (gamba.hpp:46-50)

    Parser ps(*this);
    std::string rest((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t gen = 1, i = 0;
    while (i <= rest.size()) {
        size_t j = rest.find(',', i);
        if (j == std::string::npos) j = rest.size();
        std::string g = strip_spaces(rest.substr(i, j - i));
        if (!g.empty()) ps.generator(g, gen);
        ++gen; i = j + 1;
    }
}

// One generator as text (source/io.cpp:375-449: terms joined by '+', coefficient 1 left out, exponent 1 left out).
static void format_generator(std::string& buf, const PolySystem& s, size_t off, size_t len, bool last) {
    const size_t nv = s.vars.size();
    char num[16];
    auto put = [&](uint32_t v) { auto r = std::to_chars(num, num + sizeof num, v); buf.append(num, r.ptr); };
    for (size_t j = 0; j < len; ++j) {
        const uint32_t cf = s.coeffs[off + j];
        if (cf == 0) continue;
        bool star = false;
        if (cf != 1) { put(cf); star = true; }
        bool any = false;
        const uint16_t* ex = s.exps.data() + (off + j) * nv;
        for (size_t k = 0; k < nv; ++k) {
            const uint16_t e = ex[k];
            if (!e) continue;
            if (star || any) buf += '*';
            buf += s.vars[k];
            if (e > 1) { buf += '^'; put(e); }
            any = true;
        }
        if (!any && cf == 1) buf += '1';
        if (j + 1 != len && s.coeffs[off + j + 1] > 0) buf += '+';
    }
    if (!last) buf += ',';
    buf += '\n';
}

// The basis of kat15-15 is 1.5 GB of text: generators are formatted by all host threads into per-chunk buffers (chunks of
// about equal numbers of terms, handed out dynamically) and written in order; a small system takes the serial path.
void PolySystem::write(std::ostream& out) const {
    for (auto const& v : vars) out << v << ",";
    out << "\n" << p << "\n";
    const size_t ng = lens.size();
    std::vector<size_t> off(ng + 1, 0);
    for (size_t g = 0; g < ng; ++g) off[g + 1] = off[g] + lens[g];
    const size_t terms = off[ng];
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* wt = std::getenv("GAMBA_WRITE_THREADS")) hw = (unsigned)std::max(1, std::atoi(wt));     // 1 = the serial path (tests)
    if (terms < 200000 || hw == 1) {
        std::string buf;
        for (size_t g = 0; g < ng; ++g) { buf.clear(); format_generator(buf, *this, off[g], lens[g], g + 1 == ng); out << buf; }
        out.flush();
        return;
    }
    // chunk boundaries: ~8 chunks per thread, cut at generator boundaries
    const size_t n_chunks = std::min<size_t>(ng, (size_t)hw * 8), per = (terms + n_chunks - 1) / n_chunks;
    std::vector<size_t> cut{0};
    for (size_t g = 0; g < ng; ++g)
        if (off[g + 1] >= cut.size() * per || g + 1 == ng) cut.push_back(g + 1);
    const size_t nc = cut.size() - 1;
    // workers take the chunks in order; this thread writes every chunk as soon as it is complete (formatting and the write of
    // the 2 GB file overlap), then frees it
    std::vector<std::string> bufs(nc);
    std::vector<std::atomic<int>> ready(nc);
    for (auto& r : ready) r.store(0, std::memory_order_relaxed);
    std::atomic<size_t> next{0};
    auto worker = [&] {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nc) break;
            std::string& b = bufs[c];
            b.reserve((off[cut[c + 1]] - off[cut[c]]) * 20 + 64);
            for (size_t g = cut[c]; g < cut[c + 1]; ++g) format_generator(b, *this, off[g], lens[g], g + 1 == ng);
            ready[c].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t + 1 < hw; ++t) th.emplace_back(worker);
    for (size_t c = 0; c < nc; ++c) {
        while (!ready[c].load(std::memory_order_acquire)) std::this_thread::yield();
        out.write(bufs[c].data(), (std::streamsize)bufs[c].size());
        std::string().swap(bufs[c]);
    }
    for (auto& t : th) t.join();
    out.flush();
}

}  // namespace gb
