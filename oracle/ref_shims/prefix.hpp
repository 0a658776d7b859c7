// TEST INFRASTRUCTURE (oracle/_ref build). The reference builds with a precompiled header that supplies all
// standard includes (CMakeLists.txt:138-177); this prefix header plays that role (-include prefix.hpp).
#pragma once
#include <algorithm>
#include <array>
#include <bit>
#include <cassert>
#include <chrono>
#include <climits>
#include <cmath>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <format>
#include <fstream>
#include <functional>
#include <immintrin.h>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <new>
#include <numeric>
#include <random>
#include <ranges>
#include <span>
#include <sstream>
#include <string>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <tuple>
#include <type_traits>
#include <unistd.h>
#include <unordered_map>
#include <utility>
#include <vector>
#include <xmmintrin.h>
#include <gmpxx.h>

// source/matrix.hpp:514-524 sorts three parallel arrays with std::ranges::stable_sort over a zip view, which libstdc++ 13
// cannot instantiate (its std::stable_sort wants a classic iterator category; the reference builds with clang/libc++,
// CMakeLists.txt:110-134). All standard headers are in by now, so the name is redirected for the reference's sources to
// an equivalent that sorts an index permutation stably and writes the elements back through the proxy references.
namespace std::ranges
{
struct gamba_ref_stable_sort_fn
{
    template <class R, class Comp = std::ranges::less, class Proj = std::identity>
    void operator()(R&& r, Comp comp = {}, Proj proj = {}) const
    {
        using value_t = std::ranges::range_value_t<R>;
        std::vector<value_t> tmp;
        for (auto&& e : r) tmp.emplace_back(e);
        std::vector<size_t> perm(tmp.size());
        std::iota(perm.begin(), perm.end(), size_t{0});
        std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
            return std::invoke(comp, std::invoke(proj, tmp[a]), std::invoke(proj, tmp[b]));
        });
        size_t i = 0;
        for (auto&& e : r) e = tmp[perm[i++]];
    }
};
inline constexpr gamba_ref_stable_sort_fn gamba_ref_stable_sort{};
}  // namespace std::ranges
#define stable_sort gamba_ref_stable_sort
