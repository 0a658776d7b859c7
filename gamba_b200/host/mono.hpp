// Monomials, monomial orders, divisibility masks and the open-addressing
// monomial table used by the host driver.
//
// Plays the role of the reference's monomial.hpp / order.hpp / divmask.hpp /
// flat_set.hpp (source/monomial.hpp:80-131, source/order.hpp:25-163,
// source/divmask.hpp, source/flat_set.hpp) but is an independent design:
// one flat exponent arena per table, 32-bit handles, additive hashing.
#pragma once
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

namespace gb {

// Layout of one exponent vector: [deg_total, deg_block1, e_0 .. e_{nv-1}]
struct MonoCtx {
    int nv = 0;        // number of variables
    int stride = 0;    // nv + 2
    int nelim = 0;     // size of the eliminated block (0 = plain grevlex)
    std::vector<uint32_t> w;  // hash weights
    int dm_bits = 1;   // divmask bits per variable
    int dm_vars = 0;   // variables covered by the divmask

    void init(int nvars, int num_elim, uint64_t seed) {
        nv = nvars; stride = nv + 2; nelim = num_elim;
        std::mt19937_64 rng(seed);
        w.resize(nv);
        for (auto& x : w) x = (uint32_t)rng() | 1u;
        dm_vars = nv < 64 ? nv : 64;
        dm_bits = dm_vars ? 64 / dm_vars : 1;
        if (dm_bits > 8) dm_bits = 8;
    }
    uint32_t hash(const uint16_t* e) const {
        uint32_t h = 0;
        for (int i = 0; i < nv; ++i) h += w[i] * e[2 + i];
        return h;
    }
    // bit (i*dm_bits + j) is set iff e_i > j; a | b implies mask(a) & ~mask(b) == 0
    uint64_t divmask(const uint16_t* e) const {
        uint64_t m = 0; int b = 0;
        for (int i = 0; i < dm_vars; ++i)
            for (int j = 0; j < dm_bits; ++j, ++b)
                if (e[2 + i] > j) m |= (1ull << b);
        return m;
    }
    // > 0 iff a > b in the monomial order (grevlex, or 2-block grevlex when
    // nelim > 0; same orders as source/order.hpp:25-66 and :101-163).
    int cmp(const uint16_t* a, const uint16_t* b) const {
        if (nelim > 0) {
            if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
            for (int i = nelim - 1; i >= 0; --i)
                if (a[2 + i] != b[2 + i]) return (int)b[2 + i] - (int)a[2 + i];
            int da = a[0] - a[1], db = b[0] - b[1];
            if (da != db) return da < db ? -1 : 1;
            for (int i = nv - 1; i >= nelim; --i)
                if (a[2 + i] != b[2 + i]) return (int)b[2 + i] - (int)a[2 + i];
            return 0;
        }
        if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
        for (int i = nv - 1; i >= 0; --i)
            if (a[2 + i] != b[2 + i]) return (int)b[2 + i] - (int)a[2 + i];
        return 0;
    }
    bool divides(const uint16_t* a, const uint16_t* b) const {  // a | b
        for (int i = 0; i < nv; ++i) if (a[2 + i] > b[2 + i]) return false;
        return true;
    }
};

// Open-addressing set of monomials. Handles are dense indices 0..size()-1. Storage is
// capacity-managed so that a parallel phase (insert_mt, after reserve_more) never reallocates.
struct MonoTable {
    static constexpr uint32_t PENDING = 0xffffffffu;
    const MonoCtx* cx = nullptr;
    std::vector<uint16_t> ex;     // capacity * stride
    std::vector<uint32_t> hv;     // capacity
    std::vector<uint32_t> slots;  // handle + 1, 0 = empty, PENDING = being written
    uint32_t mask = 0;
    uint32_t n = 0;               // monomials stored

    explicit MonoTable(const MonoCtx* c = nullptr, size_t cap = 1 << 16) : cx(c) {
        slots.assign(cap, 0); mask = (uint32_t)cap - 1;
    }
    size_t size() const { return n; }
    const uint16_t* e(uint32_t h) const { return &ex[(size_t)h * cx->stride]; }
    void clear() {
        n = 0;
        std::fill(slots.begin(), slots.end(), 0u);
    }
    void rehash(size_t cap) {
        slots.assign(cap, 0); mask = (uint32_t)cap - 1;
        for (uint32_t h = 0; h < n; ++h) {
            uint32_t s = hv[h] & mask;
            while (slots[s]) s = (s + 1) & mask;
            slots[s] = h + 1;
        }
    }
    // Room for `extra` more monomials without any reallocation or rehash.
    void reserve_more(size_t extra) {
        const size_t want = (size_t)n + extra;
        if (hv.size() < want) {
            size_t cap = std::max<size_t>(want, hv.size() + hv.size() / 2 + 1024);
            hv.resize(cap); ex.resize(cap * cx->stride);
        }
        if (want * 2 > slots.size()) {
            size_t cap = slots.size();
            while (want * 2 > cap) cap *= 2;
            rehash(cap);
        }
    }
    // Insert the monomial with exponent vector e (full stride, degrees filled
    // in) and precomputed hash; returns its handle.
    uint32_t insert(const uint16_t* e_, uint32_t h) {
        const int st = cx->stride;
        uint32_t s = h & mask;
        while (uint32_t v = slots[s]) {
            uint32_t cand = v - 1;
            if (hv[cand] == h && std::memcmp(&ex[(size_t)cand * st], e_, st * sizeof(uint16_t)) == 0)
                return cand;
            s = (s + 1) & mask;
        }
        if (hv.size() <= n || ((size_t)n + 1) * 2 > slots.size()) {
            // e_ may point into ex: copy before growing
            uint16_t tmp[256 + 2];
            std::memcpy(tmp, e_, st * sizeof(uint16_t));
            reserve_more(1);
            return insert(tmp, h);
        }
        const uint32_t id = n++;
        slots[s] = id + 1;
        hv[id] = h;
        std::memcpy(&ex[(size_t)id * st], e_, st * sizeof(uint16_t));
        return id;
    }
    // Concurrent insert (several threads, after reserve_more covered every insert of the phase).
    // A slot is claimed with a PENDING mark, filled, then published; readers of a pending slot wait.
    uint32_t insert_mt(const uint16_t* e_, uint32_t h) {
        const int st = cx->stride;
        uint32_t s = h & mask;
        for (;;) {
            uint32_t v = __atomic_load_n(&slots[s], __ATOMIC_ACQUIRE);
            if (v == 0) {
                uint32_t expect = 0;
                if (__atomic_compare_exchange_n(&slots[s], &expect, PENDING, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
                    const uint32_t id = __atomic_fetch_add(&n, 1u, __ATOMIC_RELAXED);
                    hv[id] = h;
                    std::memcpy(&ex[(size_t)id * st], e_, st * sizeof(uint16_t));
                    __atomic_store_n(&slots[s], id + 1, __ATOMIC_RELEASE);
                    return id;
                }
                v = expect;
            }
            while (v == PENDING) {
                __builtin_ia32_pause();
                v = __atomic_load_n(&slots[s], __ATOMIC_ACQUIRE);
            }
            const uint32_t cand = v - 1;
            if (hv[cand] == h && std::memcmp(&ex[(size_t)cand * st], e_, st * sizeof(uint16_t)) == 0) return cand;
            s = (s + 1) & mask;
        }
    }
    // Insert the product a*b of two exponent vectors (a, b may live anywhere).
    uint32_t insert_product(const uint16_t* a, uint32_t ha, const uint16_t* b, uint32_t hb) {
        const int st = cx->stride;
        uint16_t tmp[256 + 2];
        for (int i = 0; i < st; ++i) tmp[i] = (uint16_t)(a[i] + b[i]);
        return insert(tmp, ha + hb);
    }
};

}  // namespace gb
