// tools/nudge_check.cpp — nudge(po, off) (pb2_math.cuh, the branch-free step of offset_ray_origin) against the reference's
// next_float_up / next_float_down (pbrt.rs:43-77) for every binary32 value of po (or every stride-th one: argv[1]) and
// off in {+, -, +0, -0}.  Exhaustive run: 2^34 comparisons, ~15 s on 8 threads.  Exit code 1 on any difference.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
using std::isinf;
#include "../pbrt-rs_b200/csrc/pb2_math.cuh"
using namespace pb2;
static float ref(float po, float off) {
    if (off > 0.0f) return next_up(po);
    if (off < 0.0f) return next_down(po);
    return po;
}
int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1;
    std::atomic<uint64_t> bad{0}, n{0};
    auto check = [&](uint32_t u) {
        const float po = u2f(u);
        for (float off : {1.0f, -1.0f, 0.0f, -0.0f, 1e-30f, -1e30f})
            if (f2u(ref(po, off)) != f2u(nudge(po, off)) && bad++ < 10) std::printf("diff po=%08x off=%g\n", u, off);
    };
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] {
        uint64_t c = 0;
        for (uint64_t u = t * stride; u < (1ull << 32); u += T * stride) { check((uint32_t)u); ++c; }
        n += c;
    });
    for (auto& x : th) x.join();
    for (uint32_t u : {0u, 0x80000000u, 1u, 0x80000001u, 0x7f800000u, 0xff800000u, 0x7f7fffffu, 0xff7fffffu, 0x7fc00000u, 0xffc00000u, 0x00800000u, 0x80800000u}) check(u);
    std::printf("%llu values checked, %llu differences\n", (unsigned long long)n.load(), (unsigned long long)bad.load());
    return bad != 0;
}
