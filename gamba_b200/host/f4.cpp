#include "f4.hpp"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <stdexcept>

namespace gb {

static inline uint32_t mulmod(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)((uint64_t)a * b % p); }

F4::F4(const PolySystem& in, const Params& prm, Engine& eng)
    : in_(in), prm_(prm), eng_(eng), p_(in.p) {
    if (prm_.num_elim < 0 || (prm_.num_elim != 0 && (size_t)prm_.num_elim >= in.vars.size()))
        throw std::runtime_error("Elimination block size must be smaller than the number of variables.");
    nthreads_ = prm_.threads > 0 ? prm_.threads : std::max(1, std::min((int)std::thread::hardware_concurrency(), 16));
    pool_ = std::make_unique<ThreadPool>(nthreads_);
    cx_.init((int)in.vars.size(), prm_.num_elim, prm_.seed);
    bt_ = MonoTable(&cx_, 1 << 16);
    st_ = MonoTable(&cx_, 1 << 16);
    rg_ = eng_.rowgen();
    import_generators();
}

// Sort terms decreasing, merge equal monomials, drop zeros / zero polynomials,
// make monic (behaviour of source/basis.hpp:450-509).
void F4::import_generators() {
    const int nv = cx_.nv, st = cx_.stride;
    size_t off = 0;
    std::vector<uint16_t> e(st);
    for (size_t g = 0; g < in_.lens.size(); ++g) {
        std::vector<std::pair<uint32_t, uint32_t>> terms;  // (handle, coeff)
        for (size_t j = 0; j < in_.lens[g]; ++j) {
            const uint16_t* ex = &in_.exps[(off + j) * nv];
            uint32_t d = 0, d1 = 0;
            for (int k = 0; k < nv; ++k) { e[2 + k] = ex[k]; d += ex[k]; if (k < cx_.nelim) d1 += ex[k]; }
            if (d >= (1u << 16)) throw std::runtime_error("Monomial degree too large.");
            e[0] = (uint16_t)d; e[1] = (uint16_t)d1;
            terms.emplace_back(bt_.insert(e.data(), cx_.hash(e.data())), in_.coeffs[off + j] % p_);
        }
        off += in_.lens[g];
        std::stable_sort(terms.begin(), terms.end(), [&](auto const& a, auto const& b) {
            return cx_.cmp(bt_.e(a.first), bt_.e(b.first)) > 0;
        });
        Poly f;
        for (size_t j = 0; j < terms.size();) {
            size_t k = j; uint64_t s = 0;
            while (k < terms.size() && terms[k].first == terms[j].first) { s += terms[k].second; ++k; }
            s %= p_;
            if (s) { f.mon.push_back(terms[j].first); f.cf.push_back((uint32_t)s); }
            j = k;
        }
        if (f.mon.empty()) continue;
        uint32_t inv = inv_mod(f.cf[0], p_);
        for (auto& c : f.cf) c = mulmod(c, inv, p_);
        f.deg = 0;
        for (auto m : f.mon) f.deg = std::max<uint32_t>(f.deg, bt_.e(m)[0]);
        f.lm_mask = cx_.divmask(bt_.e(f.mon[0]));
        if (bt_.e(f.mon[0])[0] == 0) trivial_ = true;
        G_.push_back(std::move(f));
    }
    // generators sorted by increasing leading monomial (ties: input order)
    std::stable_sort(G_.begin(), G_.end(), [&](Poly const& a, Poly const& b) {
        return cx_.cmp(bt_.e(a.mon[0]), bt_.e(b.mon[0])) < 0;
    });
}

void F4::make_trivial() {
    trivial_ = true;
    G_.clear(); P_.clear();
    std::vector<uint16_t> e(cx_.stride, 0);
    Poly one;
    one.mon.push_back(bt_.insert(e.data(), 0));
    one.cf.push_back(1);
    one.lm_mask = 0;
    G_.push_back(std::move(one));
}

// Gebauer-Moeller installation of generator h (criteria as in
// source/update.hpp:29-269; independent code).
void F4::update(uint32_t h) {
    const int nv = cx_.nv, st = cx_.stride;
    const uint16_t* eh = bt_.e(G_[h].mon[0]);
    std::vector<uint16_t> ehv(eh, eh + st);  // bt_ may reallocate below
    eh = ehv.data();
    const uint64_t mh = G_[h].lm_mask;

    double tu = now_s();
    // B criterion on the old pairs
    {
        size_t w = 0;
        for (size_t k = 0; k < P_.size(); ++k) {
            const Pair& pr = P_[k];
            const uint16_t* el = bt_.e(pr.lcm);
            bool kill = false;
            if (cx_.divides(eh, el)) {
                const uint16_t* ei = bt_.e(G_[pr.i].mon[0]);
                const uint16_t* ej = bt_.e(G_[pr.j].mon[0]);
                bool eq_i = true, eq_j = true;
                for (int v = 0; v < nv; ++v) {
                    if (std::max(ei[2 + v], eh[2 + v]) != el[2 + v]) eq_i = false;
                    if (std::max(ej[2 + v], eh[2 + v]) != el[2 + v]) eq_j = false;
                }
                kill = !eq_i && !eq_j;
            }
            if (!kill) P_[w++] = pr; else ++stats.gm_removed;
        }
        P_.resize(w);
    }

    upd_t_[0] += now_s() - tu; tu = now_s();
    // Candidate pairs (i, h). Their lcms stay in a local, contiguous arena until the criteria have run:
    // ~99 % of the candidates are discarded, and only the survivors' lcms enter the basis table.
    struct NP { uint32_t i, at; int32_t deg; uint64_t mask, key; bool coprime, dead; };
    std::vector<NP> np;
    std::vector<uint16_t>& L = upd_lcm_;
    // one slot per older generator, filled by all threads, compacted afterwards (slot i keeps its lcm at L[i * st])
    std::vector<NP> slot(h);
    L.resize((size_t)h * st);
    std::atomic<bool> keys_bad{false}, overflow{false};
    pool_->parallel_for(h, 512, [&](size_t i0, size_t i1) {
        for (uint32_t i = (uint32_t)i0; i < (uint32_t)i1; ++i) {
            slot[i].dead = true;
            if (G_[i].redundant) continue;
            const uint16_t* ei = bt_.e(G_[i].mon[0]);
            uint16_t* el = &L[(size_t)i * st];
            uint32_t d = 0, d1 = 0; bool cop = true;
            for (int v = 0; v < nv; ++v) {
                uint16_t a = ei[2 + v], b = eh[2 + v];
                if (a && b) cop = false;
                uint16_t m = std::max(a, b);
                el[2 + v] = m; d += m; if (v < cx_.nelim) d1 += m;
            }
            if (d >= (1u << 16)) { overflow = true; continue; }
            el[0] = (uint16_t)d; el[1] = (uint16_t)d1;
            // order-preserving 64-bit prefix of the grevlex comparison (degree, then the last 7 variables reversed)
            uint64_t key = 0;
            if (cx_.nelim == 0) {
                if (d >= 256) keys_bad = true;
                key = (uint64_t)(d & 0xff) << 56;
                for (int v = nv - 1, sh = 48; v >= 0 && sh >= 0; --v, sh -= 8) key |= (uint64_t)(0xffu - (el[2 + v] & 0xffu)) << sh;
            }
            slot[i] = NP{i, i, (int32_t)d, G_[i].lm_mask | mh, key, cop, false};
        }
    });
    if (overflow) throw std::overflow_error("degree overflow");
    const bool keys_ok = cx_.nelim == 0 && !keys_bad;
    np.reserve(h);
    for (uint32_t i = 0; i < h; ++i) if (!slot[i].dead) np.push_back(slot[i]);
    auto ex_of = [&](const NP& q) { return &L[(size_t)q.at * st]; };
    auto same = [&](const NP& a, const NP& b) { return std::memcmp(ex_of(a), ex_of(b), st * sizeof(uint16_t)) == 0; };
    upd_t_[1] += now_s() - tu; tu = now_s();
    // increasing lcm; equal lcm: coprime first, then older generator first
    parallel_sort(*pool_, np, [&](NP const& a, NP const& b) {
        if (keys_ok && a.key != b.key) return a.key < b.key;
        const int c = cx_.cmp(ex_of(a), ex_of(b));
        if (c != 0) return c < 0;
        if (a.coprime != b.coprime) return a.coprime;
        return a.i < b.i;
    });
    upd_t_[2] += now_s() - tu; tu = now_s();
    // M criterion: a strictly smaller lcm divides
    for (size_t a = 0; a < np.size(); ++a) {
        if (np[a].coprime) continue;
        const uint16_t* ea = ex_of(np[a]);
        const uint64_t nm = ~np[a].mask;
        for (size_t b = 0; b < a; ++b) {
            if (np[b].mask & nm) continue;
            if (same(np[b], np[a])) continue;
            if (cx_.divides(ex_of(np[b]), ea)) { np[a].dead = true; break; }
        }
    }
    upd_t_[3] += now_s() - tu; tu = now_s();
    // F criterion: keep the first pair of every run of equal lcm
    for (size_t a = 1; a < np.size(); ++a) {
        if (np[a].coprime || np[a].dead) continue;
        for (size_t b = a; b > 0 && same(np[b - 1], np[a]); --b)
            if (!np[b - 1].dead) { np[a].dead = true; break; }
    }
    for (auto const& q : np) {
        if (q.dead || q.coprime) { ++stats.gm_removed; continue; }
        const uint16_t* el = ex_of(q);
        P_.push_back({q.i, h, bt_.insert(el, cx_.hash(el)), q.deg});
    }
    upd_t_[4] += now_s() - tu; tu = now_s();
    // older generators whose lead is divisible by the new lead become redundant
    for (uint32_t i = 0; i < h; ++i) {
        if (G_[i].redundant) continue;
        if (mh & ~G_[i].lm_mask) continue;
        if (cx_.divides(eh, bt_.e(G_[i].mon[0]))) G_[i].redundant = true;
    }
}

// Gebauer-Moeller installation of the generators [from, n) as ONE batch - the same pairs, in the same order, as update(from),
// update(from + 1), ... one after the other, but the generators are worked on by all host threads at once (a step of kat15-15
// brings up to 1 700 new generators with ~8 000 candidate pairs each: 3 s of a 16 s run when taken one at a time):
//   0. when each older generator turns redundant: red_at[i] = the first new generator h > i whose lead divides lead(i) - update(h')
//      sees generator i iff red_at[i] >= h';
//   1. per new generator h, one thread: candidate pairs (i, h), sort by lcm, M and F criteria -> its surviving pairs with their lcms
//      in a local arena (nothing is written to the basis table here);
//   2. B criterion: a pair (old, or new from h') dies iff some new generator h (h > h') kills it - independent per pair;
//   3. the survivors enter the queue in the sequential order (old pairs, then the pairs of from, from + 1, ...), their lcms the table.
void F4::update_batch(size_t from) {
    const size_t n = G_.size();
    const int nv = cx_.nv, st = cx_.stride;
    constexpr uint32_t NONE = 0xffffffffu;
    double tu = now_s();
    // ---- 0. redundancy times ------------------------------------------------------------------------------------------
    std::vector<uint32_t> red_at(n, NONE);
    pool_->parallel_for(n, 64, [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; ++i) {
            if (G_[i].redundant) { red_at[i] = 0; continue; }
            const uint16_t* ei = bt_.e(G_[i].mon[0]);
            const uint64_t mi = G_[i].lm_mask;
            for (size_t h = std::max(from, i + 1); h < n; ++h) {
                if (G_[h].lm_mask & ~mi) continue;
                if (cx_.divides(bt_.e(G_[h].mon[0]), ei)) { red_at[i] = (uint32_t)h; break; }
            }
        }
    });
    upd_t_[0] += now_s() - tu; tu = now_s();
    // ---- 1. candidate pairs of every new generator ---------------------------------------------------------------------
    struct Out { std::vector<uint32_t> i; std::vector<int32_t> deg; std::vector<uint16_t> ex; uint64_t removed = 0; };
    std::vector<Out> outs(n - from);
    std::atomic<bool> overflow{false};
    pool_->parallel_for(n - from, 1, [&](size_t a0, size_t a1) {
        struct NP { uint32_t i, at; int32_t deg; uint64_t mask, key; bool coprime, dead; };
        std::vector<NP> np;
        std::vector<uint16_t> L;
        for (size_t a = a0; a < a1; ++a) {
            const uint32_t h = (uint32_t)(from + a);
            Out& o = outs[a];
            const uint16_t* eh = bt_.e(G_[h].mon[0]);
            const uint64_t mh = G_[h].lm_mask;
            np.clear();
            L.resize((size_t)h * st);
            bool keys_bad = false;
            for (uint32_t i = 0; i < h; ++i) {
                if (red_at[i] < h) continue;
                const uint16_t* ei = bt_.e(G_[i].mon[0]);
                uint16_t* el = &L[(size_t)i * st];
                uint32_t d = 0, d1 = 0; bool cop = true;
                for (int v = 0; v < nv; ++v) {
                    const uint16_t x = ei[2 + v], y = eh[2 + v];
                    if (x && y) cop = false;
                    const uint16_t m = std::max(x, y);
                    el[2 + v] = m; d += m; if (v < cx_.nelim) d1 += m;
                }
                if (d >= (1u << 16)) { overflow = true; continue; }
                el[0] = (uint16_t)d; el[1] = (uint16_t)d1;
                uint64_t key = 0;      // order-preserving 64-bit prefix of the grevlex comparison, as in update()
                if (cx_.nelim == 0) {
                    if (d >= 256) keys_bad = true;
                    key = (uint64_t)(d & 0xff) << 56;
                    for (int v = nv - 1, sh = 48; v >= 0 && sh >= 0; --v, sh -= 8) key |= (uint64_t)(0xffu - (el[2 + v] & 0xffu)) << sh;
                }
                np.push_back(NP{i, i, (int32_t)d, G_[i].lm_mask | mh, key, cop, false});
            }
            const bool keys_ok = cx_.nelim == 0 && !keys_bad;
            auto ex_of = [&](const NP& q) { return &L[(size_t)q.at * st]; };
            auto same = [&](const NP& x, const NP& y) { return std::memcmp(ex_of(x), ex_of(y), st * sizeof(uint16_t)) == 0; };
            std::sort(np.begin(), np.end(), [&](NP const& x, NP const& y) {
                if (keys_ok && x.key != y.key) return x.key < y.key;
                const int c = cx_.cmp(ex_of(x), ex_of(y));
                if (c != 0) return c < 0;
                if (x.coprime != y.coprime) return x.coprime;
                return x.i < y.i;
            });
            for (size_t x = 0; x < np.size(); ++x) {           // M criterion: a strictly smaller lcm divides
                if (np[x].coprime) continue;
                const uint16_t* ea = ex_of(np[x]);
                const uint64_t nm = ~np[x].mask;
                for (size_t y = 0; y < x; ++y) {
                    if (np[y].mask & nm) continue;
                    if (same(np[y], np[x])) continue;
                    if (cx_.divides(ex_of(np[y]), ea)) { np[x].dead = true; break; }
                }
            }
            for (size_t x = 1; x < np.size(); ++x) {           // F criterion: the first pair of every run of equal lcm stays
                if (np[x].coprime || np[x].dead) continue;
                for (size_t y = x; y > 0 && same(np[y - 1], np[x]); --y)
                    if (!np[y - 1].dead) { np[x].dead = true; break; }
            }
            for (auto const& q : np) {
                if (q.dead || q.coprime) { ++o.removed; continue; }
                o.i.push_back(q.i); o.deg.push_back(q.deg);
                const uint16_t* el = ex_of(q);
                o.ex.insert(o.ex.end(), el, el + st);
            }
        }
    });
    if (overflow) throw std::overflow_error("degree overflow");
    upd_t_[1] += now_s() - tu; tu = now_s();
    // ---- 2. B criterion --------------------------------------------------------------------------------------------------
    // pair with lcm el, generators (pi, pj), dies by h iff lead(h) | el and lcm(pi, h) != el and lcm(pj, h) != el
    auto killed_by = [&](const uint16_t* el, uint64_t ml, uint32_t pi, uint32_t pj, size_t h_lo) {
        const uint16_t* ei = bt_.e(G_[pi].mon[0]);
        const uint16_t* ej = bt_.e(G_[pj].mon[0]);
        for (size_t h = h_lo; h < n; ++h) {
            if (G_[h].lm_mask & ~ml) continue;
            const uint16_t* eh = bt_.e(G_[h].mon[0]);
            if (!cx_.divides(eh, el)) continue;
            bool eq_i = true, eq_j = true;
            for (int v = 0; v < nv; ++v) {
                if (std::max(ei[2 + v], eh[2 + v]) != el[2 + v]) eq_i = false;
                if (std::max(ej[2 + v], eh[2 + v]) != el[2 + v]) eq_j = false;
            }
            if (!eq_i && !eq_j) return true;
        }
        return false;
    };
    std::vector<uint8_t> dead_old(P_.size(), 0);
    pool_->parallel_for(P_.size(), 256, [&](size_t k0, size_t k1) {
        for (size_t k = k0; k < k1; ++k) {
            const uint16_t* el = bt_.e(P_[k].lcm);
            dead_old[k] = killed_by(el, cx_.divmask(el), P_[k].i, P_[k].j, from) ? 1 : 0;
        }
    });
    std::vector<std::vector<uint8_t>> dead_new(n - from);
    pool_->parallel_for(n - from, 1, [&](size_t a0, size_t a1) {
        for (size_t a = a0; a < a1; ++a) {
            const Out& o = outs[a];
            dead_new[a].assign(o.i.size(), 0);
            for (size_t k = 0; k < o.i.size(); ++k) {
                const uint16_t* el = &o.ex[k * st];
                dead_new[a][k] = killed_by(el, cx_.divmask(el), o.i[k], (uint32_t)(from + a), from + a + 1) ? 1 : 0;
            }
        }
    });
    upd_t_[2] += now_s() - tu; tu = now_s();
    // ---- 3. the queue, in the order the one-at-a-time installation leaves it ---------------------------------------------
    {
        size_t w = 0;
        for (size_t k = 0; k < P_.size(); ++k) { if (!dead_old[k]) P_[w++] = P_[k]; else ++stats.gm_removed; }
        P_.resize(w);
    }
    for (size_t a = 0; a < n - from; ++a) {
        const Out& o = outs[a];
        stats.gm_removed += o.removed;
        for (size_t k = 0; k < o.i.size(); ++k) {
            if (dead_new[a][k]) { ++stats.gm_removed; continue; }
            const uint16_t* el = &o.ex[k * st];
            P_.push_back({o.i[k], (uint32_t)(from + a), bt_.insert(el, cx_.hash(el)), o.deg[k]});
        }
    }
    for (size_t i = 0; i < n; ++i) if (red_at[i] != NONE) G_[i].redundant = true;
    upd_t_[4] += now_s() - tu;
}

void F4::update_all(size_t from) {
    double t0 = now_s();
    // a few new generators: one at a time, each with all threads; a batch of them: all at once, one thread each
    static const bool batch = [] { const char* e = std::getenv("GAMBA_UPDATE_BATCH"); return !e || std::atoi(e) != 0; }();     // 0: tests
    if (batch && G_.size() - from >= 8 && nthreads_ > 1) update_batch(from);
    else for (size_t h = from; h < G_.size(); ++h) update((uint32_t)h);
    stats.t_update += now_s() - t0;
}

// Normal strategy: pairs of minimal degree, at most max_spairs of them, in
// increasing lcm order (source/spair.hpp:190-264).
long F4::select_pairs(int32_t& deg) {
    double t0 = now_s();
    long n;
    if (prm_.all_pairs) {
        std::sort(P_.begin(), P_.end(), [&](Pair const& a, Pair const& b) {
            if (a.lcm == b.lcm) return false;
            return cx_.cmp(bt_.e(a.lcm), bt_.e(b.lcm)) < 0;
        });
        n = (long)P_.size(); deg = 0xffff;
    } else {
        std::sort(P_.begin(), P_.end(), [&](Pair const& a, Pair const& b) {
            if (a.deg != b.deg) return a.deg < b.deg;
            if (a.lcm == b.lcm) return false;
            return cx_.cmp(bt_.e(a.lcm), bt_.e(b.lcm)) < 0;
        });
        deg = P_[0].deg;
        long cnt = 0;
        while ((size_t)cnt < P_.size() && P_[cnt].deg == deg) ++cnt;
        n = cnt;
        if (prm_.max_spairs > 0 && n > prm_.max_spairs) n = prm_.max_spairs;
    }
    stats.pairs_reduced += n;
    stats.t_select += now_s() - t0;
    return n;
}

void F4::clear_step() {
    st_.clear();
    row_mon_.clear(); row_off_.assign(1, 0); row_gen_.clear(); row_top_.clear(); row_lead_.clear();
    has_red_.clear();
    if (rg_) rg_->begin((uint32_t)cx_.stride);
}

// Append the rows gen[i] * x^q[i] (q: flat exponent vectors of full stride) to the step: every term
// is multiplied and looked up / inserted in the step table. The rows are independent, so they are
// generated by all threads against the concurrent table (mono.hpp: insert_mt); row numbering is the
// order of the batch, hence deterministic, and nothing downstream depends on the handle numbering.
void F4::add_rows(const std::vector<uint32_t>& gens, const std::vector<uint16_t>& qs, const std::vector<uint32_t>& hqs,
                  const std::vector<uint8_t>& tops) {
    const int st = cx_.stride;
    const size_t nb = gens.size();
    if (nb == 0) return;
    const size_t r0 = row_gen_.size();
    uint64_t terms = 0;
    row_off_.reserve(row_off_.size() + nb);
    for (size_t i = 0; i < nb; ++i) { terms += G_[gens[i]].mon.size(); row_off_.push_back(row_off_[r0] + terms); }
    if (!rg_) row_mon_.resize(row_off_.back());
    row_gen_.insert(row_gen_.end(), gens.begin(), gens.end());
    row_top_.insert(row_top_.end(), tops.begin(), tops.end());
    if (!rg_) {   // contiguous term data of the generators of this batch (once per generator)
        std::vector<uint32_t> need;
        for (uint32_t g : gens) if (G_[g].hv.empty()) { G_[g].hv.resize(1); need.push_back(g); }
        pool_->parallel_for(need.size(), 4, [&](size_t a, size_t b) {
            for (size_t x = a; x < b; ++x) {
                Poly& f = G_[need[x]];
                const size_t n = f.mon.size();
                f.ex.resize(n * st); f.hv.resize(n);
                for (size_t k = 0; k < n; ++k) {
                    std::memcpy(&f.ex[k * st], bt_.e(f.mon[k]), st * sizeof(uint16_t));
                    f.hv[k] = bt_.hv[f.mon[k]];
                }
            }
        });
    }
    row_lead_.resize(r0 + nb);
    if (rg_) {
        // Rows on the engine's device: the basis polynomials' terms are uploaded once, the products are looked up / inserted in
        // the device's monomial table and the handles stay there (the staging kernels read them in place). Back come one lead
        // handle per row and the monomials the batch introduced - mirrored into st_ (exponents and hashes only: nothing is
        // looked up on the host any more) for the reducer search, the column sort and the new basis elements.
        double tA = now_s();
        std::vector<uint64_t> first(nb);
        std::vector<uint32_t> nt(nb);
        {   // the basis table's new monomials, then the polynomials not yet in the device's term store (as handles into that
            // table: 4 bytes per term): one upload each
            if (bt_dev_n_ < bt_.n) {
                const uint64_t first_m = rg_->add_monomials(bt_.n - bt_dev_n_, bt_.e(bt_dev_n_), &bt_.hv[bt_dev_n_]);
                if (first_m != bt_dev_n_) throw std::logic_error("device basis table out of step");
                bt_dev_n_ = bt_.n;
            }
            std::vector<uint32_t> ug, un; std::vector<const uint32_t*> um;
            for (size_t i = 0; i < nb; ++i) {
                Poly& f = G_[gens[i]];
                if (f.dev_first == ~0ull) { f.dev_first = ~0ull - 1; ug.push_back(gens[i]); un.push_back((uint32_t)f.mon.size()); um.push_back(f.mon.data()); }
            }
            if (!ug.empty()) {
                std::vector<uint64_t> uf(ug.size());
                rg_->add_polys((uint32_t)ug.size(), un.data(), um.data(), uf.data());
                for (size_t k = 0; k < ug.size(); ++k) G_[ug[k]].dev_first = uf[k];
            }
        }
        for (size_t i = 0; i < nb; ++i) { const Poly& f = G_[gens[i]]; first[i] = f.dev_first; nt[i] = (uint32_t)f.mon.size(); }
        const uint32_t before = st_.n;
        const uint32_t after = rg_->add_rows((uint32_t)nb, first.data(), nt.data(), qs.data(), hqs.data(), &row_lead_[r0]);
        if (after > before) {
            if (st_.hv.size() < after) { const size_t cap = std::max<size_t>(after, st_.hv.size() + st_.hv.size() / 2 + 1024); st_.hv.resize(cap); st_.ex.resize(cap * st); }
            rg_->fetch(before, after - before, &st_.ex[(size_t)before * st], &st_.hv[before]);
            st_.n = after;
        }
        dbg_t_[1] += now_s() - tA;
    } else
    // Sub-batches: the table must have room for every term of a batch being NEW (no growth inside the
    // parallel region), while almost all terms are hits - so a batch is capped near the table's size.
    for (size_t b0 = 0; b0 < nb;) {
        double tA = now_s();
        static const uint64_t cap_min = std::getenv("GAMBA_ROWGEN_CAP") ? std::strtoull(std::getenv("GAMBA_ROWGEN_CAP"), nullptr, 10) : (1u << 20);
        const uint64_t cap = std::max<uint64_t>(cap_min, st_.size());
        size_t b1 = b0; uint64_t bt = 0;
        while (b1 < nb && (b1 == b0 || bt + G_[gens[b1]].mon.size() <= cap)) bt += G_[gens[b1++]].mon.size();
        st_.reserve_more(bt);
        dbg_t_[0] += now_s() - tA; tA = now_s();
        pool_->parallel_for(b1 - b0, 8, [&](size_t i0, size_t i1) {
            for (size_t i = b0 + i0; i < b0 + i1; ++i) {
                const Poly& f = G_[gens[i]];
                const uint16_t* q = &qs[i * st];
                const uint32_t hq = hqs[i];
                uint32_t* out = &row_mon_[row_off_[r0 + i]];
                uint16_t tmp[258];
                const uint16_t* a = f.ex.data();
                const uint32_t* fh = f.hv.data();
                // A lookup is three dependent cache misses (slot -> hash of the candidate -> its exponents) in a table of
                // millions of monomials: the slot of term k + 16 and the candidate of term k + 8 are prefetched
                const uint32_t msk = st_.mask;
                const uint32_t* sl = st_.slots.data();
                const size_t n = f.mon.size();
                for (size_t k = 0; k < n; ++k, a += st) {
                    if (k + 16 < n) __builtin_prefetch(&sl[(fh[k + 16] + hq) & msk]);
                    if (k + 8 < n) {
                        const uint32_t v = __atomic_load_n(&sl[(fh[k + 8] + hq) & msk], __ATOMIC_RELAXED);
                        if (v != 0 && v != MonoTable::PENDING) {
                            __builtin_prefetch(&st_.hv[v - 1]);
                            __builtin_prefetch(&st_.ex[(size_t)(v - 1) * st]);
                        }
                    }
                    for (int j = 0; j < st; ++j) tmp[j] = (uint16_t)(a[j] + q[j]);
                    out[k] = st_.insert_mt(tmp, fh[k] + hq);
                }
            }
        });
        dbg_t_[1] += now_s() - tA;
        b0 = b1;
    }
    if (!rg_) for (size_t i = 0; i < nb; ++i) row_lead_[r0 + i] = row_mon_[row_off_[r0 + i]];
    if (has_red_.size() < st_.size()) has_red_.resize(st_.size(), 0);
    for (size_t i = 0; i < nb; ++i)
        if (tops[i]) has_red_[row_lead_[r0 + i]] = 1;
}

// Rows of the selected pairs (source/matrix.hpp:215-347): per distinct lcm the
// shortest generator gives the pivot row, all others rows to reduce.
void F4::build_pair_rows(long n_sel) {
    double t0 = now_s();
    const int st = cx_.stride;
    std::vector<uint32_t> gens, bg, bh;
    std::vector<uint16_t> bq;
    std::vector<uint8_t> bt;
    std::vector<uint16_t> q(st);
    for (long a = 0; a < n_sel;) {
        long b = a; gens.clear();
        while (b < n_sel && P_[b].lcm == P_[a].lcm) { gens.push_back(P_[b].i); gens.push_back(P_[b].j); ++b; }
        std::sort(gens.begin(), gens.end());
        gens.erase(std::unique(gens.begin(), gens.end()), gens.end());
        size_t best = 0;
        for (size_t k = 1; k < gens.size(); ++k) {
            const Poly& x = G_[gens[k]]; const Poly& y = G_[gens[best]];
            if (x.mon.size() < y.mon.size() || (x.mon.size() == y.mon.size() && x.deg < y.deg)) best = k;
        }
        std::swap(gens[0], gens[best]);
        const uint16_t* el = bt_.e(P_[a].lcm);
        for (size_t k = 0; k < gens.size(); ++k) {
            const uint16_t* lm = bt_.e(G_[gens[k]].mon[0]);
            for (int i = 0; i < st; ++i) q[i] = (uint16_t)(el[i] - lm[i]);
            bg.push_back(gens[k]); bq.insert(bq.end(), q.begin(), q.end()); bh.push_back(cx_.hash(q.data())); bt.push_back(k == 0);
        }
        a = b;
    }
    add_rows(bg, bq, bh, bt);
    P_.erase(P_.begin(), P_.begin() + n_sel);
    stats.t_rows += now_s() - t0;
}

std::vector<uint32_t> F4::nonredundant() const {
    std::vector<uint32_t> v;
    for (uint32_t i = 0; i < G_.size(); ++i) if (!G_[i].redundant) v.push_back(i);
    return v;
}

// For every monomial of the step without a reducer, pick the first generator
// (shortest first) whose lead divides it and add the multiplied row as a pivot
// row (source/matrix.hpp:469-600).
void F4::symbolic() {
    double t0 = now_s();
    const int st = cx_.stride, nv = cx_.nv;
    std::vector<uint32_t> cand = nonredundant();
    std::stable_sort(cand.begin(), cand.end(), [&](uint32_t a, uint32_t b) {
        if (G_[a].mon.size() != G_[b].mon.size()) return G_[a].mon.size() < G_[b].mon.size();
        return G_[a].deg < G_[b].deg;
    });
    const size_t nc = cand.size();
    std::vector<uint64_t> cmask(nc);
    std::vector<uint16_t> cexp(nc * nv);
    for (size_t c = 0; c < nc; ++c) {
        cmask[c] = G_[cand[c]].lm_mask;
        std::memcpy(&cexp[c * nv], bt_.e(G_[cand[c]].mon[0]) + 2, nv * sizeof(uint16_t));
    }
    // Waves: all monomials that appeared in the previous wave are searched for a reducer in parallel
    // (read-only), then the multiplied reducer rows are generated in parallel. The closure - hence the
    // matrix - is the same as with the reference's one-monomial-at-a-time worklist.
    std::vector<uint32_t> hit, bg, bh;
    std::vector<uint16_t> bq;
    std::vector<uint8_t> bt;
    for (size_t m0 = 0; m0 < st_.size();) {
        const size_t m1 = st_.size();
        if (has_red_.size() < m1) has_red_.resize(m1, 0);
        hit.assign(m1 - m0, (uint32_t)nc);
        double tB = now_s();
        pool_->parallel_for(m1 - m0, 64, [&](size_t a0, size_t a1) {
        for (size_t m = m0 + a0; m < m0 + a1; ++m) {
            if (has_red_[m]) continue;
            const uint16_t* em = st_.e((uint32_t)m);
            const uint64_t nm = ~cx_.divmask(em);
            for (size_t c = 0; c < nc; ++c) {
                if (cmask[c] & nm) continue;
                const uint16_t* le = &cexp[c * nv];
                bool ok = true;
                for (int v = 0; v < nv; ++v) if (le[v] > em[2 + v]) { ok = false; break; }
                if (ok) { hit[m - m0] = (uint32_t)c; break; }
            }
        }
        });
        dbg_t_[2] += now_s() - tB; dbg_waves_++;
        bg.clear(); bq.clear(); bh.clear(); bt.clear();
        for (size_t m = m0; m < m1; ++m) {
            const uint32_t c = hit[m - m0];
            if (c == nc) continue;
            const uint16_t* em = st_.e((uint32_t)m);
            const uint16_t* lm = bt_.e(G_[cand[c]].mon[0]);
            const size_t o = bq.size();
            bq.resize(o + st);
            for (int i = 0; i < st; ++i) bq[o + i] = (uint16_t)(em[i] - lm[i]);
            bg.push_back(cand[c]); bh.push_back(cx_.hash(&bq[o])); bt.push_back(1);
        }
        add_rows(bg, bq, bh, bt);
        m0 = m1;
    }
    if (has_red_.size() < st_.size()) has_red_.resize(st_.size(), 0);
    stats.t_symbolic += now_s() - t0;
}

// Columns = monomials in decreasing order; then block form (engine.hpp):
// pivot columns first, pivot row i leads at column i, rows to reduce sorted by
// decreasing lead (source/matrix.hpp:603-673 + the A|B / C|D split of SURVEY §8(b)).
void F4::convert() {
    double t0 = now_s();
    const uint32_t nc = (uint32_t)st_.size();
    order_.resize(nc);
    std::iota(order_.begin(), order_.end(), 0u);
    parallel_sort(*pool_, order_, [&](uint32_t a, uint32_t b) { return cx_.cmp(st_.e(a), st_.e(b)) > 0; });
    std::vector<uint32_t> colof(nc);
    for (uint32_t c = 0; c < nc; ++c) colof[order_[c]] = c;

    const size_t nrows = row_gen_.size();
    std::vector<uint32_t> top, bot;
    for (uint32_t r = 0; r < nrows; ++r) (row_top_[r] ? top : bot).push_back(r);
    auto lead = [&](uint32_t r) { return colof[row_lead_[r]]; };
    std::sort(top.begin(), top.end(), [&](uint32_t a, uint32_t b) { return lead(a) < lead(b); });
    std::stable_sort(bot.begin(), bot.end(), [&](uint32_t a, uint32_t b) { return lead(a) > lead(b); });

    std::vector<uint8_t> is_piv(nc, 0);
    for (uint32_t r : top) {
        if (is_piv[lead(r)]) throw std::logic_error("two pivot rows share a lead column");
        is_piv[lead(r)] = 1;
    }
    std::vector<uint32_t> newcol(nc);
    nonpiv_col_.clear();
    uint32_t np = 0, nn = 0;
    const uint32_t ntop = (uint32_t)top.size();
    for (uint32_t c = 0; c < nc; ++c) {
        if (is_piv[c]) newcol[c] = np++;
        else { newcol[c] = ntop + nn++; nonpiv_col_.push_back(c); }
    }
    // handle -> block-form column in one table
    std::vector<uint32_t> h2c(nc);
    for (uint32_t h = 0; h < nc; ++h) h2c[h] = newcol[colof[h]];

    M_.p = p_; M_.n_top = ntop; M_.n_bot = (uint32_t)bot.size(); M_.n_cols = nc;
    // Indirect hand-over (include/gamba_cuda.h): the rows stay where they were generated (row_mon_, as
    // step-table handles); the engine gets their offsets in block-form order and the handle -> column map.
    M_.row_ptr.assign(nrows + 1, 0);
    M_.row_src.assign(nrows, 0);
    M_.coeff_id.assign(nrows, 0); M_.cf_ptr.clear(); M_.cf_len.clear();
    static std::vector<int32_t> gmap;
    gmap.assign(G_.size(), -1);
    uint64_t w = 0;
    uint32_t i = 0;
    auto place = [&](uint32_t r) {
        w += row_off_[r + 1] - row_off_[r];
        M_.row_ptr[i + 1] = w;
        M_.row_src[i] = row_off_[r];
        const uint32_t g = row_gen_[r];
        if (gmap[g] < 0) {
            gmap[g] = (int32_t)M_.cf_ptr.size();
            M_.cf_ptr.push_back(G_[g].cf.data());
            M_.cf_len.push_back((uint32_t)G_[g].cf.size());
        }
        M_.coeff_id[i++] = (uint32_t)gmap[g];
    };
    for (uint32_t r : top) place(r);
    for (uint32_t r : bot) place(r);
    if (rg_) { uint64_t n = 0; M_.src = rg_->rows_device(n); M_.src_len = n; M_.src_on_device = true; }
    else { M_.src = row_mon_.data(); M_.src_len = row_mon_.size(); M_.src_on_device = false; }
    M_.col_map.swap(h2c);
    stats.t_convert += now_s() - t0;
    stats.max_rows = std::max<uint64_t>(stats.max_rows, nrows);
    stats.max_cols = std::max<uint64_t>(stats.max_cols, nc);
    stats.max_nnz = std::max<uint64_t>(stats.max_nnz, w);
}

void F4::dump_step(const char* tag) {
    if (prm_.dump_dir.empty() || (long)M_.nnz() < prm_.dump_min_nnz) return;
    const bool is_final = std::string(tag) == "final";
    // --dump-every N: an unbiased sample of a long run (bench.py: every 6th of kat15-15's 60 steps)
    if (prm_.dump_every > 1 && !is_final && (stats.steps % (uint64_t)prm_.dump_every) != 0) return;
    if (prm_.dump_top > 0 && !is_final) {
        // --dump-top N: only the N steps with the most non-zeros stay on disk (the big systems' runs have 60+ steps of
        // hundreds of MB each); a step smaller than all N kept ones is not written at all
        if ((long)dump_kept_.size() >= prm_.dump_top) {
            auto mn = std::min_element(dump_kept_.begin(), dump_kept_.end());
            if (mn->first >= M_.nnz()) return;
            std::remove((prm_.dump_dir + "/step_" + mn->second + ".bin").c_str());
            std::remove((prm_.dump_dir + "/rows_" + mn->second + ".bin").c_str());
            dump_kept_.erase(mn);
        }
        dump_kept_.emplace_back(M_.nnz(), tag);
    }
    char name[64];
    std::snprintf(name, sizeof name, "/step_%s.bin", tag);
    if (M_.src_on_device) {       // the rows live on the GPU: bring them over for the file
        row_mon_.resize(M_.src_len);
        rg_->download_rows(0, M_.src_len, row_mon_.data());
        const uint32_t* dev = M_.src;
        M_.src = row_mon_.data(); M_.src_on_device = false;
        write_step_file(prm_.dump_dir + name, M_);
        M_.src = dev; M_.src_on_device = true;
    } else
    write_step_file(prm_.dump_dir + name, M_);
    std::snprintf(name, sizeof name, "/rows_%s.bin", tag);
    write_rows_file(prm_.dump_dir + name, R_);
}

// New pivot rows -> basis elements (source/matrix.hpp:799-895 +
// source/basis.hpp:697-769): relative non-pivot index -> column -> monomial.
void F4::insert_new_rows() {
    double t0 = now_s();
    // the new rows only hold non-pivot columns (a few thousand distinct monomials for up to 10^8 terms): every column's
    // monomial enters the basis table once, the terms index that map
    constexpr uint32_t UNSET = 0xffffffffu;
    std::vector<uint32_t> col2bt(nonpiv_col_.size(), UNSET);
    std::vector<uint8_t> used(nonpiv_col_.size(), 0);
    pool_->parallel_for(R_.n_new, 1, [&](size_t r0, size_t r1) {
        for (uint64_t k = R_.row_ptr[r0]; k < R_.row_ptr[r1]; ++k) __atomic_store_n(&used[R_.col[k]], (uint8_t)1, __ATOMIC_RELAXED);
    });
    for (size_t c = 0; c < used.size(); ++c)
        if (used[c]) { const uint32_t h = order_[nonpiv_col_[c]]; col2bt[c] = bt_.insert(st_.e(h), st_.hv[h]); }
    // the polynomials themselves (up to 10^8 terms in a step of kat15-15) are filled by all threads
    std::vector<Poly> fs(R_.n_new);
    pool_->parallel_for(R_.n_new, 1, [&](size_t r0, size_t r1) {
        for (size_t r = r0; r < r1; ++r) {
            Poly& f = fs[r];
            const uint64_t a = R_.row_ptr[r], b = R_.row_ptr[r + 1];
            f.mon.resize(b - a); f.cf.assign(R_.val.begin() + a, R_.val.begin() + b);
            for (uint64_t k = a; k < b; ++k) {
                const uint32_t c = R_.col[k];
                f.mon[k - a] = col2bt[c];
                f.deg = std::max<uint32_t>(f.deg, st_.e(order_[nonpiv_col_[c]])[0]);
            }
            if (b > a) f.lm_mask = cx_.divmask(bt_.e(f.mon[0]));
        }
    });
    for (uint32_t r = 0; r < R_.n_new; ++r) {
        Poly& f = fs[r];
        if (f.cf.empty() || f.cf[0] != 1) throw std::logic_error("engine returned a non-monic row");
        if (bt_.e(f.mon[0])[0] == 0) { make_trivial(); break; }
        G_.push_back(std::move(f));
    }
    stats.t_insert += now_s() - t0;
}

void F4::final_reduce() {
    double t0 = now_s();
    clear_step();
    std::vector<uint32_t> nr = nonredundant();
    std::vector<uint16_t> one(cx_.stride, 0);
    {
        std::vector<uint16_t> qs(nr.size() * cx_.stride, 0);
        std::vector<uint32_t> hs(nr.size(), 0);
        std::vector<uint8_t> tp(nr.size(), 0);
        add_rows(nr, qs, hs, tp);
        for (size_t i = 0; i < nr.size(); ++i) has_red_[row_lead_[i]] = 1;  // leads need no reducer
    }
    symbolic();
    convert();
    double t1 = now_s();
    eng_.reduce_final(M_, R_);
    stats.t_final_linalg += now_s() - t1;
    dump_step("final");
    if (R_.n_new != M_.n_bot) throw std::logic_error("final reduction lost a basis element");
    std::vector<Poly> out;
    std::vector<uint32_t> col2bt(nonpiv_col_.size(), 0xffffffffu);
    for (uint32_t r = 0; r < R_.n_new; ++r) {
        Poly f;
        uint64_t a = R_.row_ptr[r], b = R_.row_ptr[r + 1];
        f.mon.resize(b - a); f.cf.assign(R_.val.begin() + a, R_.val.begin() + b);
        for (uint64_t k = a; k < b; ++k) {
            const uint32_t c = R_.col[k], h = order_[nonpiv_col_[c]];
            if (col2bt[c] == 0xffffffffu) col2bt[c] = bt_.insert(st_.e(h), st_.hv[h]);
            f.mon[k - a] = col2bt[c];
        }
        f.lm_mask = cx_.divmask(bt_.e(f.mon[0]));
        out.push_back(std::move(f));
    }
    G_ = std::move(out);
    stats.t_final += now_s() - t0;
}

void F4::run() {
    double t_start = now_s();
    const bool rounds = prm_.verbose > 1;
    if (trivial_) make_trivial();
    if (!trivial_) update_all(0);
    if (rounds)
        std::printf("%4s %4s %14s %22s %8s %9s %10s %10s\n", "step", "deg", "pairs/queue",
                    "matrix (rows x cols)", "dens %", "new rows", "linalg s", "step s");
    bool stopped = false;
    while (!P_.empty() && !trivial_) {
        if (prm_.stop_after > 0 && (long)stats.steps >= prm_.stop_after) { stopped = true; break; }
        double ts = now_s();
        size_t prev = G_.size();
        int32_t deg; size_t queue = P_.size();
        long nsel = select_pairs(deg);
        clear_step();
        build_pair_rows(nsel);
        symbolic();
        stats.rows_reduced += std::count(row_top_.begin(), row_top_.end(), 0);
        convert();
        double t1 = now_s();
        eng_.reduce_step(M_, R_);
        double tl = now_s() - t1;
        stats.t_linalg += tl; stats.t_linalg_kernel += R_.kernel_seconds;
        stats.nnz_in += M_.nnz(); stats.fma_top += R_.fma_top;
        stats.zero_reductions += M_.n_bot - R_.n_new;
        ++stats.steps;
        char tag[32]; std::snprintf(tag, sizeof tag, "%04llu", (unsigned long long)stats.steps);
        dump_step(tag);
        insert_new_rows();
        if (!trivial_) update_all(prev);
        if (rounds) {
            double dens = 100.0 * (double)M_.nnz() / ((double)M_.n_rows() * (double)M_.n_cols);
            std::printf("%4llu %4d %7ld/%-6zu %10u x %-9u %8.2f %9u %10.3f %10.3f   %s\n",
                        (unsigned long long)stats.steps, deg, nsel, queue, M_.n_rows(), M_.n_cols, dens,
                        R_.n_new, tl, now_s() - ts, R_.note.c_str());
            std::fflush(stdout);
        }
    }
    // keep only generators whose lead is minimal (source/basis.hpp:822-882)
    if (!trivial_) {
        std::vector<uint32_t> nr = nonredundant();
        for (uint32_t i : nr)
            for (uint32_t j : nr) {
                if (i == j || G_[j].redundant) continue;
                if (G_[j].lm_mask & ~G_[i].lm_mask) continue;
                if (G_[i].mon[0] != G_[j].mon[0] && cx_.divides(bt_.e(G_[j].mon[0]), bt_.e(G_[i].mon[0]))) {
                    G_[i].redundant = true; break;
                }
            }
        if (!prm_.no_reduce && !stopped) final_reduce();
        else {
            std::vector<Poly> out;
            for (uint32_t i : nonredundant()) out.push_back(std::move(G_[i]));
            G_ = std::move(out);
        }
    }
    stats.t_total = now_s() - t_start;
    if (getenv("GAMBA_DEBUG_TIMERS")) std::fprintf(stderr, "dbg: reserve %.3f rowgen %.3f search %.3f waves %ld\n", dbg_t_[0], dbg_t_[1], dbg_t_[2], dbg_waves_),
        std::fprintf(stderr, "dbg update: B %.3f build %.3f sort %.3f M %.3f F+push %.3f\n", upd_t_[0], upd_t_[1], upd_t_[2], upd_t_[3], upd_t_[4]);
}

// Generators in increasing order of leading monomial (source/basis.hpp:923-972).
void F4::export_basis(PolySystem& out) const {
    out.vars = in_.vars; out.p = p_;
    out.lens.clear(); out.coeffs.clear(); out.exps.clear();
    std::vector<uint32_t> idx(G_.size());
    std::iota(idx.begin(), idx.end(), 0u);
    std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
        return cx_.cmp(bt_.e(G_[a].mon[0]), bt_.e(G_[b].mon[0])) < 0;
    });
    // kat15-15: 8 k generators, ~10^8 terms - filled in parallel at precomputed offsets
    const size_t ng = idx.size(), nv = cx_.nv;
    std::vector<size_t> off(ng + 1, 0);
    out.lens.resize(ng);
    for (size_t g = 0; g < ng; ++g) { out.lens[g] = (uint32_t)G_[idx[g]].mon.size(); off[g + 1] = off[g] + out.lens[g]; }
    out.coeffs.resize(off[ng]);
    out.exps.resize(off[ng] * nv);
    pool_->parallel_for(ng, 1, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            const Poly& f = G_[idx[g]];
            uint32_t* cf = out.coeffs.data() + off[g];
            uint16_t* ex = out.exps.data() + off[g] * nv;
            for (size_t k = 0; k < f.mon.size(); ++k) {
                cf[k] = f.cf[k];
                const uint16_t* e = bt_.e(f.mon[k]);
                std::copy(e + 2, e + 2 + nv, ex + k * nv);
            }
        }
    });
}

}  // namespace gb
