// TEST INFRASTRUCTURE (oracle/_ref build). Stand-in for <gmpxx.h>: GMP's headers are not in this image (only the
// runtime library). The reference touches GMP only in its parser / writer (source/io.cpp:83-87,155-181,308-320,364-372),
// in basis.hpp's rational branch (:439,488-497, never instantiated for F_p) and field.hpp:186-200; every shipped example
// has coefficients below 2^64, so 128-bit integers are enough. Overflow is not detected.
#pragma once
#include <cstdint>
#include <iostream>
#include <numeric>
#include <string>

#define __GNU_MP_VERSION 0
#define __GNU_MP_VERSION_MINOR 0
#define __GNU_MP_VERSION_PATCHLEVEL 0

class mpz_class
{
public:
    using int_t = __int128;
    int_t v{0};
    mpz_class() = default;
    mpz_class(int x) : v(x) {}
    mpz_class(long x) : v(x) {}
    mpz_class(long long x) : v(x) {}
    mpz_class(unsigned x) : v(x) {}
    mpz_class(unsigned long x) : v(x) {}
    mpz_class(unsigned long long x) : v(x) {}
    static mpz_class from_int128(int_t x) { mpz_class r; r.v = x; return r; }
    explicit mpz_class(std::string const& s, int /*base*/ = 10)
    {
        size_t i = 0; bool neg = false;
        if (i < s.size() && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
        if (i >= s.size()) throw std::invalid_argument("mpz_class: empty");
        for (; i < s.size(); ++i)
        {
            if (s[i] < '0' || s[i] > '9') throw std::invalid_argument("mpz_class: bad digit");
            v = v * 10 + (s[i] - '0');
        }
        if (neg) v = -v;
    }
    mpz_class const* get_mpz_t() const { return this; }
    unsigned long get_ui() const { return static_cast<unsigned long>(v); }
    long get_si() const { return static_cast<long>(v); }
    std::string get_str(int = 10) const
    {
        if (v == 0) return "0";
        int_t x = v; bool neg = x < 0; if (neg) x = -x;
        std::string s;
        while (x > 0) { s.insert(s.begin(), char('0' + int(x % 10))); x /= 10; }
        return neg ? "-" + s : s;
    }
    mpz_class& operator+=(mpz_class const& o) { v += o.v; return *this; }
    mpz_class& operator-=(mpz_class const& o) { v -= o.v; return *this; }
    mpz_class& operator*=(mpz_class const& o) { v *= o.v; return *this; }
    mpz_class& operator/=(mpz_class const& o) { v /= o.v; return *this; }
    mpz_class& operator%=(mpz_class const& o) { v %= o.v; return *this; }
    mpz_class operator-() const { return from_int128(-v); }
    friend mpz_class operator+(mpz_class a, mpz_class const& b) { return a += b; }
    friend mpz_class operator-(mpz_class a, mpz_class const& b) { return a -= b; }
    friend mpz_class operator*(mpz_class a, mpz_class const& b) { return a *= b; }
    friend mpz_class operator/(mpz_class a, mpz_class const& b) { return a /= b; }
    friend mpz_class operator%(mpz_class a, mpz_class const& b) { return a %= b; }
    friend bool operator==(mpz_class const& a, mpz_class const& b) { return a.v == b.v; }
    friend bool operator!=(mpz_class const& a, mpz_class const& b) { return a.v != b.v; }
    friend bool operator<(mpz_class const& a, mpz_class const& b) { return a.v < b.v; }
    friend bool operator>(mpz_class const& a, mpz_class const& b) { return a.v > b.v; }
    friend bool operator<=(mpz_class const& a, mpz_class const& b) { return a.v <= b.v; }
    friend bool operator>=(mpz_class const& a, mpz_class const& b) { return a.v >= b.v; }
    friend std::ostream& operator<<(std::ostream& os, mpz_class const& a) { return os << a.get_str(); }
};

inline mpz_class abs(mpz_class const& a) { return a.v < 0 ? -a : a; }
inline mpz_class gcd(mpz_class a, mpz_class b)
{
    a = abs(a); b = abs(b);
    while (b.v != 0) { mpz_class t = a % b; a = b; b = t; }
    return a;
}
inline mpz_class lcm(mpz_class const& a, mpz_class const& b)
{
    if (a.v == 0 || b.v == 0) return mpz_class{0};
    return abs(a / gcd(a, b) * b);
}

class mpq_class
{
public:
    mpz_class num{0}, den{1};
    mpq_class() = default;
    mpq_class(long n, long d) : num(n), den(d) {}
    mpq_class(mpz_class const& n) : num(n), den(1) {}
    explicit mpq_class(std::string const& s, int base = 10)
    {
        size_t const slash = s.find('/');
        if (slash == std::string::npos) { num = mpz_class(s, base); den = mpz_class(1); }
        else { num = mpz_class(s.substr(0, slash), base); den = mpz_class(s.substr(slash + 1), base); }
    }
    void canonicalize()
    {
        if (den.v < 0) { num = -num; den = -den; }
        mpz_class const g = gcd(num, den);
        if (g.v > 1) { num /= g; den /= g; }
    }
    mpz_class const& get_num() const { return num; }
    mpz_class const& get_den() const { return den; }
    friend mpq_class operator*(mpq_class const& a, mpz_class const& b)
    {
        mpq_class r; r.num = a.num * b; r.den = a.den; r.canonicalize(); return r;
    }
    friend mpq_class operator*(mpq_class const& a, mpq_class const& b)
    {
        mpq_class r; r.num = a.num * b.num; r.den = a.den * b.den; r.canonicalize(); return r;
    }
    friend mpq_class operator-(mpq_class const& a) { mpq_class r = a; r.num = -r.num; return r; }
    mpq_class& operator*=(long const x) { num *= mpz_class(x); canonicalize(); return *this; }
    mpq_class& operator*=(mpz_class const& x) { num *= x; canonicalize(); return *this; }
    friend std::ostream& operator<<(std::ostream& os, mpq_class const& a) { return os << a.num << "/" << a.den; }
};

inline unsigned long mpz_fdiv_ui(mpz_class const* a, unsigned long d)
{
    __int128 r = a->v % static_cast<__int128>(d);
    if (r < 0) r += d;
    return static_cast<unsigned long>(r);
}

// deterministic Miller-Rabin for n < 2^64 (the reference asks for 50 rounds of the probabilistic test, io.cpp:84)
inline int mpz_probab_prime_p(mpz_class const* a, int /*reps*/)
{
    if (a->v < 2) return 0;
    unsigned long long const n = static_cast<unsigned long long>(a->v);
    for (unsigned long long q : {2ULL, 3ULL, 5ULL, 7ULL, 11ULL, 13ULL, 17ULL, 19ULL, 23ULL, 29ULL, 31ULL, 37ULL})
    {
        if (n == q) return 2;
        if (n % q == 0) return 0;
    }
    auto mulmod = [n](unsigned long long x, unsigned long long y) { return static_cast<unsigned long long>((unsigned __int128)x * y % n); };
    auto powmod = [&](unsigned long long b, unsigned long long e) { unsigned long long r = 1; b %= n; while (e) { if (e & 1) r = mulmod(r, b); b = mulmod(b, b); e >>= 1; } return r; };
    unsigned long long d = n - 1; int s = 0;
    while ((d & 1) == 0) { d >>= 1; ++s; }
    for (unsigned long long w : {2ULL, 3ULL, 5ULL, 7ULL, 11ULL, 13ULL, 17ULL, 19ULL, 23ULL, 29ULL, 31ULL, 37ULL})
    {
        unsigned long long x = powmod(w, d);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int r = 1; r < s && comp; ++r) { x = mulmod(x, x); if (x == n - 1) comp = false; }
        if (comp) return 0;
    }
    return 2;
}
