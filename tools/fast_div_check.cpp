// tools/fast_div_check.cpp — fast_div (wavefront.cuh: n / d = hi64(n * ceil(2^64 / d))) against n / d on 104 M (n, d) pairs: film widths, powers of two,
// multiples +- 2, the top of the 32-bit range, random pairs.  g++ -O2 -o fdc tools/fast_div_check.cpp && ./fdc
#include <cstdint>
#include <cstdio>
#include <random>
static unsigned long long magic(uint32_t d) { return d <= 1u ? 0ull : 0xFFFFFFFFFFFFFFFFull / d + 1ull; }
static uint32_t fdiv(uint32_t n, unsigned long long m) { return m ? (uint32_t)(((unsigned __int128)n * m) >> 64) : n; }
int main() {
    std::mt19937_64 g(1);
    unsigned long long bad = 0, cnt = 0;
    const uint32_t ds[] = {1, 2, 3, 5, 7, 16, 24, 96, 128, 160, 512, 1000, 1920, 3840, 16384, 65535, 65536, 262144, 2073600, 8294400, 0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFFu};
    for (uint32_t d : ds) {
        const unsigned long long m = magic(d);
        for (uint64_t k = 0; k < 4000000; ++k) { const uint32_t n = (uint32_t)g(); ++cnt; if (fdiv(n, m) != n / d) ++bad; }
        for (uint64_t q = 0; q < 100000; ++q) for (int o = -2; o <= 2; ++o) { const uint64_t n = q * d + o; if (n > 0xFFFFFFFFull) continue; ++cnt; if (fdiv((uint32_t)n, m) != (uint32_t)n / d) ++bad; }
        for (uint64_t n = 0xFFFFFFFFull - 100000; n <= 0xFFFFFFFFull; ++n) { ++cnt; if (fdiv((uint32_t)n, m) != (uint32_t)n / d) ++bad; }
    }
    for (int k = 0; k < 2000000; ++k) { uint32_t d = (uint32_t)g(); if (!d) d = 1; const uint32_t n = (uint32_t)g(); ++cnt; if (fdiv(n, magic(d)) != n / d) ++bad; }
    printf("%llu checks, %llu differences\n", cnt, bad);
    return bad != 0;
}
