// wavefront.cu — the wavefront PathIntegrator: ray-gen -> extend (closest hit) -> material-sorted shade ->
// shadow (any hit) + MIS (closest hit), which fold their answers into L -> next bounce, and the Film accumulation kernels (sm_100a).
//
// Replaces SamplerIntegrator::render (src/core/integrator.rs:399-480), PathIntegrator::li (src/integrators/path.rs:65-213),
// uniform_sample_one_light / estimate_direct (src/core/integrator.rs:92-266), DiffuseAreaLight / PointLight sampling
// (src/lights/diffuse.rs:60-90,150-156, src/lights/point.rs:47-74, src/core/shape.rs:38-69, src/shapes/triangle.rs:323-348),
// Interaction::spawn_ray / spawn_ray_to (src/core/interaction.rs:132-153) and FilmTile::add_sample / merge_film_tile
// (src/core/film.rs:252-295,111-123).  Each (pixel, sample) owns the sampler stream RNG::new(pixel_index*spp + sample)
// and draws from it in the reference's order (5 camera dims; per vertex: light pick, u_light, u_scattering, BSDF sample,
// Russian roulette), so a path's radiance does not depend on scheduling.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "camera.cuh"
#include "trace_persistent.cuh"
#include "wavefront.cuh"

namespace pb2 {

TraceTuning trace_tuning();

namespace {

constexpr float kInf = __builtin_huge_valf();
constexpr int kThreads = 256;

// ---- order-preserving queues -------------------------------------------------------------------------------------------------
// The stages talk through queues of path slots.  The visibility / shade kernels do not append to them: each leaves one byte per
// path (PathBuffers::state) and compact_queues() rebuilds the queues from the bounce's active queue with a stable three-way
// select, so every queue stays sorted by slot and neighbouring lanes of the next kernel read neighbouring path state.
// (Appending in completion order of the persistent traversal — one warp-aggregated atomic per group of lanes — scattered a
// warp's slots over a ~150 K-slot window; with the queues sorted k_shade runs 34-39 % faster, k_shadow 11-22 %, k_extend 3-8 %:
// profiles/r01_tuning.md, session 4.)
constexpr unsigned kStateDead = 3u;            // bits 0-1: shading class of the hit, 3 = the ray escaped
constexpr unsigned kStateContinues = 4u;       // bit 2: the path continues with the ray k_shade wrote
                                               // bits 3, 4: shadow ray / MIS ray of the NEE record pending

// compact_queues(): ONE launch — a stable three-way select with decoupled look-back.  The active queue is cut into tiles of
// kSelTile entries; a CTA takes the next tile (ticket counter, so a tile's predecessors are always running or done), counts its
// entries per output, publishes the three counts, obtains its exclusive offsets by looking back over the predecessors'
// published counts / inclusive prefixes (one 64-bit word per tile and output: {launch epoch : 32 | flag : 2 | value : 30}, so a
// word is valid only for the launch that wrote it and the array never needs clearing), and scatters.  The input is read once.
// The CTA of the last tile leaves the totals in the queue counters and does the per-bounce bookkeeping k_iter_begin used to do in
// a launch of its own.  (Round 1: count -> one-CTA scan -> scatter, three launches that read the queue and the state bytes twice.)
constexpr int kSelThreads = 256, kSelItems = 16, kSelTile = kSelThreads * kSelItems;
struct SelectJob {
    const uint32_t* in;                        // the bounce's active queue (sorted by slot)
    const unsigned long long* n_in;
    const uint8_t* state;
    uint32_t* out[3];
    unsigned long long* n_out[3];
    int by_class;                              // 1: out[k] takes class k (after k_extend); 0: out[k] takes bit 2 + k (after k_shade)
    unsigned long long* status;                // [3][max_tiles]
    uint32_t max_tiles;
    unsigned* ticket;                          // this launch's tile counter (zero on entry)
    unsigned* ticket_next;                     // the next launch's: zeroed here
    unsigned epoch;                            // > 0, different for every launch
    unsigned long long* counters;
};
__device__ __forceinline__ bool select_pred(int by_class, unsigned st, int k) {
    return by_class ? (st & 3u) == (unsigned)k : ((st >> (2 + k)) & 1u) != 0u;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// After the queues of a bounce are rebuilt from the shade stage's state bytes: the ray totals of the frame counters and the
// work counters of the next two traversal launches (k_shadow of this bounce, k_extend of the next).
__device__ __forceinline__ void select_finish(const SelectJob& j, unsigned long long t0, unsigned long long t1, unsigned long long t2) {
    *j.n_out[0] = t0; *j.n_out[1] = t1; *j.n_out[2] = t2;
    if (!j.by_class) {
        j.counters[T_EXTEND] += t0;
        j.counters[T_SHADOW] += t1;
        j.counters[T_MIS] += t2;
        j.counters[C_WORK_EXTEND] = 0;
        j.counters[C_WORK_SHADOW] = 0;
    }
}
// Exclusive offset of `tile` for one output: the sum of the predecessors' counts, walking back 64 tiles per round (two
// descriptors per lane in flight) until one that already holds its inclusive prefix.  Run by one warp per output; the
// tiles in flight started at about the same time, so the walk covers most of them and its depth is what the pass costs.
__device__ __forceinline__ unsigned select_look_back(const unsigned long long* st, unsigned tile, unsigned epoch, unsigned lane) {
    const unsigned long long tag = (unsigned long long)epoch << 32;
    unsigned excl = 0u;
    int look = (int)tile - 1;
    for (;;) {
        const int i0 = look - (int)lane, i1 = look - 32 - (int)lane;            // nearest first: lane 0 of the first word
        unsigned long long w0, w1;
        unsigned incl0, incl1;
        for (;;) {                                   // wait for every predecessor up to the nearest inclusive prefix
            w0 = i0 >= 0 ? ld_status(st + i0) : (tag | (2ull << 30));
            w1 = i1 >= 0 ? ld_status(st + i1) : (tag | (2ull << 30));
            const bool v0 = (w0 >> 32) == (unsigned long long)epoch && ((w0 >> 30) & 3ull) != 0ull;
            const bool v1 = (w1 >> 32) == (unsigned long long)epoch && ((w1 >> 30) & 3ull) != 0ull;
            const unsigned vm0 = __ballot_sync(0xFFFFFFFFu, v0), vm1 = __ballot_sync(0xFFFFFFFFu, v1);
            incl0 = __ballot_sync(0xFFFFFFFFu, v0 && ((w0 >> 30) & 3ull) == 2ull);
            incl1 = __ballot_sync(0xFFFFFFFFu, v1 && ((w1 >> 30) & 3ull) == 2ull);
            if (incl0) {
                const unsigned need = (1u << (__ffs(incl0) - 1)) - 1u;
                if ((vm0 & need) == need) break;
            } else if (vm0 == 0xFFFFFFFFu) {
                const unsigned need = incl1 ? ((1u << (__ffs(incl1) - 1)) - 1u) : 0xFFFFFFFFu;
                if ((vm1 & need) == need) break;
            }
        }
        // sum the counts up to and including the nearest predecessor that holds an inclusive prefix
        const int stop0 = incl0 ? __ffs(incl0) - 1 : 31;
        const int stop1 = incl0 ? -1 : (incl1 ? __ffs(incl1) - 1 : 31);
        unsigned v = ((int)lane <= stop0 ? (unsigned)(w0 & 0x3FFFFFFFull) : 0u) + ((int)lane <= stop1 ? (unsigned)(w1 & 0x3FFFFFFFull) : 0u);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        excl += v;
        if (incl0 || incl1) break;
        look -= 64;
    }
    return excl;
}
__global__ void __launch_bounds__(kSelThreads) k_select3(SelectJob j) {
    __shared__ unsigned s_tile;
    __shared__ unsigned s_warp[3][kSelThreads / 32];
    __shared__ unsigned s_base[3];
    const unsigned long long n = *j.n_in;
    const unsigned n_tiles = (unsigned)((n + kSelTile - 1) / kSelTile);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long tag = (unsigned long long)j.epoch << 32;
    for (;;) {
        if (threadIdx.x == 0) {
            s_tile = atomicAdd(j.ticket, 1u);
            if (s_tile == 0u) {
                *j.ticket_next = 0u;
                if (n_tiles == 0u) select_finish(j, 0ull, 0ull, 0ull);
            }
        }
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= n_tiles) return;
        // ---- load kSelItems consecutive entries per thread (blocked: the select is stable) ----
        const unsigned long long i0 = (unsigned long long)tile * kSelTile + (unsigned long long)threadIdx.x * kSelItems;
        uint32_t slot[kSelItems];
        if (i0 + kSelItems <= n) {
#pragma unroll
            for (int k = 0; k < kSelItems; k += 4) {
                const uint4 a = *reinterpret_cast<const uint4*>(j.in + i0 + k);
                slot[k] = a.x; slot[k + 1] = a.y; slot[k + 2] = a.z; slot[k + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kSelItems; ++k) slot[k] = i0 + k < n ? j.in[i0 + k] : 0u;
        }
        unsigned bits[3] = {0u, 0u, 0u};           // bit k of bits[q]: entry k goes to output q
#pragma unroll
        for (int k = 0; k < kSelItems; ++k) {
            const unsigned st = i0 + k < n ? j.state[slot[k]] : (j.by_class ? kStateDead : 0u);
#pragma unroll
            for (int q = 0; q < 3; ++q) bits[q] |= (select_pred(j.by_class, st, q) ? 1u : 0u) << k;
        }
        unsigned cnt[3], inc[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) inc[q] = cnt[q] = (unsigned)__popc(bits[q]);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const unsigned v = __shfl_up_sync(0xFFFFFFFFu, inc[q], o);
                if (lane >= (unsigned)o) inc[q] += v;
            }
        if (lane == 31u)
#pragma unroll
            for (int q = 0; q < 3; ++q) s_warp[q][warp] = inc[q];
        __syncthreads();
        // ---- warps 0-2: tile total of output `warp`, publish it, look back ----
        if (warp < 3u) {
            const int q = (int)warp;
            unsigned total = 0u;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; ++w) total += s_warp[q][w];
            unsigned long long* st = j.status + (size_t)q * j.max_tiles;
            if (tile == 0u) {
                if (lane == 0u) { st_status(st, tag | (2ull << 30) | total); s_base[q] = 0u; }
            } else {
                if (lane == 0u) st_status(st + tile, tag | (1ull << 30) | total);
                const unsigned excl = select_look_back(st, tile, j.epoch, lane);
                if (lane == 0u) { st_status(st + tile, tag | (2ull << 30) | (unsigned long long)(excl + total)); s_base[q] = excl; }
            }
        }
        __syncthreads();
        // ---- scatter ----
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            unsigned before = 0u;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; ++w) before += (unsigned)w < warp ? s_warp[q][w] : 0u;
            unsigned pos = s_base[q] + before + inc[q] - cnt[q];
            uint32_t* out = j.out[q];
#pragma unroll
            for (int k = 0; k < kSelItems; ++k)
                if ((bits[q] >> k) & 1u) out[pos++] = slot[k];
        }
        if (tile == n_tiles - 1u && threadIdx.x == kSelThreads - 1) {
            // (the last thread of the last tile: its inclusive position is the total)
            unsigned long long tot[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                unsigned before = 0u;
#pragma unroll
                for (int w = 0; w < kSelThreads / 32; ++w) before += (unsigned)w < warp ? s_warp[q][w] : 0u;
                tot[q] = (unsigned long long)s_base[q] + before + inc[q];
            }
            select_finish(j, tot[0], tot[1], tot[2]);
        }
        __syncthreads();                                 // s_tile, s_warp, s_base are reused by the next tile
    }
}

// ---- slot <-> (pixel, sample) -----------------------------------------------------------------------------------
struct SlotInfo {
    int x, y;                 // pixel in image coordinates (may lie outside the image for wide filters)
    uint32_t sample;          // sample index of the pixel
    unsigned long long seq;   // sampler stream: ((y - sb_y0) * sb_w + (x - sb_x0)) * spp + sample
    uint32_t pix;             // (y - sb_y0) * sb_w + (x - sb_x0)
};
// (32-bit arithmetic: a wavefront holds at most 2^28 slots, and a 64-bit modulo costs ~100 instructions in every kernel
// that resumes a sampler)
__device__ __forceinline__ SlotInfo slot_info(const PathMap& m, const FilmView& f, uint64_t slot64) {
    const uint32_t slot = (uint32_t)slot64;
    SlotInfo s;
    if (m.explicit_xy) {
        s.x = m.explicit_xy[2ull * slot];
        s.y = m.explicit_xy[2ull * slot + 1];
        s.sample = m.explicit_s[slot];
        s.pix = (uint32_t)(s.y - f.sb_y0) * (uint32_t)f.sb_w + (uint32_t)(s.x - f.sb_x0);
    } else {
        const uint32_t s_local = slot / m.n_pix;
        s.pix = slot - s_local * m.n_pix;
        s.sample = (uint32_t)m.sample0 + s_local;
        const uint32_t row = s.pix / (uint32_t)f.sb_w;
        s.x = f.sb_x0 + (int)(s.pix - row * (uint32_t)f.sb_w);
        s.y = f.sb_y0 + (int)row;
    }
    s.seq = (unsigned long long)s.pix * (unsigned long long)m.spp + s.sample;
    return s;
}

// ---- samplers --------------------------------------------------------------------------------------------------------------
// HaltonSampler (samplers/halton.rs, core/lowdiscrepancy.rs:293-390; pbrt-v3 semantics where the port is broken).
__device__ __forceinline__ unsigned long long inverse_radical_inverse(unsigned base, unsigned long long inverse, int n_digits) {
    unsigned long long index = 0;
    for (int i = 0; i < n_digits; ++i) {
        const unsigned long long digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}
__device__ __forceinline__ long long halton_index(const SamplerView& h, int px, int py, unsigned long long sample_num) {    // halton.rs:117-141
    long long offset = 0;
    if (h.sample_stride > 1ull) {
        const int pm0 = ((px % 128) + 128) % 128, pm1 = ((py % 128) + 128) % 128;
        offset += (long long)(inverse_radical_inverse(2u, (unsigned long long)pm0, h.base_exponents[0]) *
                              (h.sample_stride / (unsigned long long)h.base_scales[0]) * h.mult_inverse[0]);
        offset += (long long)(inverse_radical_inverse(3u, (unsigned long long)pm1, h.base_exponents[1]) *
                              (h.sample_stride / (unsigned long long)h.base_scales[1]) * h.mult_inverse[1]);
        offset %= (long long)h.sample_stride;
    }
    return offset + (long long)(sample_num * h.sample_stride);
}
// Digit loop of radical_inverse_specialized / scramble_radical_inverse_specialized (lowdiscrepancy.rs:293-320): integer
// digits are exact, so a 32-bit index (every practical frame: index < stride * spp) takes 32-bit divisions.
template <class UInt, bool SCRAMBLED>
__device__ __forceinline__ void halton_digits(UInt a, UInt base, const uint16_t* perm, float inv_base, unsigned long long* reversed, float* inv_base_n) {
    while (a != 0) {
        const UInt next = a / base, digit = a - next * base;
        *reversed = *reversed * base + (SCRAMBLED ? (unsigned long long)perm[digit] : (unsigned long long)digit);
        *inv_base_n = *inv_base_n * inv_base;
        a = next;
    }
}
__device__ __forceinline__ float halton_dimension(const SamplerView& h, unsigned long long index, unsigned dim) {            // halton.rs:143-155
    if (dim == 0u) return __ull2float_rn(__brevll(index >> h.base_exponents[0])) * 5.4210108624275222e-20f;
    const unsigned long long a = dim == 1u ? index / (unsigned long long)h.base_scales[1] : index;
    const unsigned base = h.primes[dim];
    const float inv_base = 1.0f / (float)base;
    unsigned long long reversed = 0;
    float inv_base_n = 1.0f;
    if (dim == 1u) {                                                                       // radical_inverse
        if (a >> 32) halton_digits<unsigned long long, false>(a, base, nullptr, inv_base, &reversed, &inv_base_n);
        else halton_digits<unsigned, false>((unsigned)a, base, nullptr, inv_base, &reversed, &inv_base_n);
        return fminf(__ull2float_rn(reversed) * inv_base_n, PB2_ONE_MINUS_EPS);
    }
    const uint16_t* perm = h.perms + h.prime_sums[dim];                                    // scramble_radical_inverse
    if (a >> 32) halton_digits<unsigned long long, true>(a, base, perm, inv_base, &reversed, &inv_base_n);
    else halton_digits<unsigned, true>((unsigned)a, base, perm, inv_base, &reversed, &inv_base_n);
    return fminf(inv_base_n * (__ull2float_rn(reversed) + inv_base * (float)perm[0] / (1.0f - inv_base)), PB2_ONE_MINUS_EPS);
}
// SobolSampler (samplers/sobol.rs:48-58, lowdiscrepancy.rs:507-560; pbrt-v3 semantics where the port cannot run: DESIGN.md).
constexpr unsigned kSobolMatrixSize = 52u;                                 // sobolmatrices.rs:2
__device__ __forceinline__ unsigned long long sobol_index(const SamplerView& h, int px, int py, unsigned long long frame) {     // sobol_interval_to_index
    const unsigned m = (unsigned)h.sobol_log2_resolution;
    if (m == 0u) return 0ull;
    unsigned long long index = frame << (m << 1), delta = 0ull;
    const unsigned long long* vdc = h.sobol_vdc + (size_t)(m - 1u) * kSobolMatrixSize;
    const unsigned long long* inv = h.sobol_vdc_inv + (size_t)(m - 1u) * kSobolMatrixSize;
    while (frame) {                                                        // XOR over the set bits, in any order
        delta ^= __ldg(vdc + (__ffsll((long long)frame) - 1));
        frame &= frame - 1ull;
    }
    unsigned long long b = (unsigned long long)((((unsigned)px) << m) | (unsigned)py) ^ delta;
    while (b) {
        index ^= __ldg(inv + (__ffsll((long long)b) - 1));
        b &= b - 1ull;
    }
    return index;
}
__device__ __forceinline__ float sobol_raw(const SamplerView& h, unsigned long long a, unsigned dim) {       // sobol_sample, scramble = 0
    const uint32_t* m = h.sobol_m32 + (size_t)dim * kSobolMatrixSize;
    unsigned v = 0u;
    while (a) {
        v ^= __ldg(m + (__ffsll((long long)a) - 1));
        a &= a - 1ull;
    }
    return fminf(PB2_ONE_MINUS_EPS, __uint2float_rn(v) * 2.3283064365386963e-10f);
}
// SobolSampler::sample_dimension for dimensions 0 / 1 (pbrt-v3: the film position inside pixel (x, y))
__device__ __forceinline__ float sobol_pixel_dimension(const SamplerView& h, unsigned long long index, unsigned dim, int pixel) {
    float s = sobol_raw(h, index, dim);
    s = s * (float)h.sobol_resolution + (float)h.sobol_min[dim];
    s = s - (float)pixel;
    return s < 0.0f ? 0.0f : (s > PB2_ONE_MINUS_EPS ? PB2_ONE_MINUS_EPS : s);
}
// One path's sampler.  RandomSampler: the PCG32 stream; HaltonSampler: (index, dimension) of the sequence; PixelSamplers
// (stratified, (0,2)): PixelSampler::get_1d / get_2d (sampler.rs:289-307) — the next tabulated dimension of this pixel's
// sample while one is left, then the PCG32 stream.  TABLES = false compiles the table branch out (k_shade is at its register limit).
struct PathSampler {
    Pcg32 rng;
    unsigned long long index;
    unsigned dim;
    unsigned cur1, cur2;          // current_1d_dimension, current_2d_dimension
    unsigned tab_base;            // sample * tab_n_pix + pixel
    const SamplerView* h;
    __device__ __forceinline__ bool sobol() const { return h->kind == 4; }
    __device__ __forceinline__ bool global() const { return h->kind == 1 || h->kind == 4; }          // GlobalSampler (sampler.rs:324-410)
    __device__ __forceinline__ bool tables() const { return h->kind == 2 || h->kind == 3; }         // PixelSampler (:257-322)
    __device__ __forceinline__ unsigned long long global_index(const SlotInfo& si) const {         // get_index_for_sample
        return sobol() ? sobol_index(*h, si.x - h->sobol_min[0], si.y - h->sobol_min[1], si.sample)
                       : (unsigned long long)halton_index(*h, si.x, si.y, si.sample);
    }
    __device__ __forceinline__ float global_dimension(unsigned d) const {                          // sample_dimension, d >= 2 for Sobol'
        return sobol() ? sobol_raw(*h, index, d) : halton_dimension(*h, index, d);
    }
    __device__ __forceinline__ void start(const SamplerView& view, const SlotInfo& si) {   // start of a pixel sample
        h = &view;
        dim = 0u;
        cur1 = cur2 = 0u;
        tab_base = si.sample * view.tab_n_pix + si.pix;
        if (global()) index = global_index(si);
        else rng.set_sequence(si.seq);
    }
    // The first draw of every pixel sample: CameraSample::p_film's offset inside the pixel (sampler.rs:27-33).  Sobol' remaps
    // dimensions 0 / 1 to the pixel, which needs the pixel's coordinates — known here, not carried in the path's sampler state.
    __device__ __forceinline__ void film_offset(const SlotInfo& si, float* u0, float* u1) {
        if (sobol()) {
            *u0 = sobol_pixel_dimension(*h, index, 0u, si.x);
            *u1 = sobol_pixel_dimension(*h, index, 1u, si.y);
            dim = 2u;
            return;
        }
        next2(u0, u1);
    }
    // `extra` = the PixelSampler dimension counters kept in bits 17-30 of the path's state word
    __device__ __forceinline__ void resume(const SamplerView& view, const SlotInfo& si, unsigned long long saved, unsigned extra) {
        h = &view;
        cur1 = extra & 0x7Fu;
        cur2 = (extra >> 7) & 0x7Fu;
        tab_base = si.sample * view.tab_n_pix + si.pix;
        if (global()) { index = global_index(si); dim = (unsigned)saved; }
        else { rng.state = saved; rng.inc = (si.seq << 1) | 1ull; }
    }
    __device__ __forceinline__ unsigned long long save() const { return global() ? (unsigned long long)dim : rng.state; }
    __device__ __forceinline__ unsigned extra() const { return cur1 | (cur2 << 7); }
    template <bool TABLES = true>
    __device__ __forceinline__ float next1() {                                             // Sampler::get_1d
        if (TABLES && tables() && cur1 < (unsigned)h->n_dims) {
            const float v = __ldg(h->t1 + (size_t)(cur1 * (unsigned)h->spp_tab) * h->tab_n_pix + tab_base);
            ++cur1;
            return v;
        }
        if (global()) return global_dimension(dim++);
        return rng.next_float();
    }
    template <bool TABLES = true>
    __device__ __forceinline__ void next2(float* a, float* b) {                            // Sampler::get_2d, x then y
        if (TABLES && tables() && cur2 < (unsigned)h->n_dims) {
            const float2 v = __ldg(h->t2 + (size_t)(cur2 * (unsigned)h->spp_tab) * h->tab_n_pix + tab_base);
            ++cur2;
            *a = v.x; *b = v.y;
            return;
        }
        if (global()) { *a = global_dimension(dim); *b = global_dimension(dim + 1u); dim += 2u; return; }
        *a = rng.next_float();
        *b = rng.next_float();
    }
};

// ---- PixelSampler tables: Sampler::start_pixel for every pixel of the sample bounds ---------------------------------------
// One thread per pixel walks its own RNG stream exactly as StratifiedSampler::start_pixel (stratified.rs:44-76) /
// ZeroTwoSequenceSampler::start_pixel (zerotwosequence.rs:31-48) do (1D dimensions first, then 2D; generate, then shuffle),
// on columns of the tables (stride tab_n_pix).  pbrt-v3 semantics where the port cannot run: DESIGN.md §9 (P1-P4).
struct TableGen {
    int kind, spp, n_dims, xs, ys, jitter;
    uint32_t n_pix;
    unsigned long long seq0;
    float* t1;
    float2* t2;
};
__device__ __forceinline__ uint32_t pcg_bounded(Pcg32& rng, uint32_t b) {                  // rng.rs:36-44 uniform_u32_u32
    const uint32_t threshold = (~b + 1u) % b;
    for (;;) {
        const uint32_t r = rng.next_u32();
        if (r >= threshold) return r % b;
    }
}
template <class T>
__device__ __forceinline__ void shuffle_column(T* col, uint32_t stride, int count, Pcg32& rng) {      // sampling.rs:280-287
    for (int i = 0; i < count; ++i) {
        const int other = i + (int)pcg_bounded(rng, (uint32_t)(count - i));
        const T a = col[(size_t)i * stride], b = col[(size_t)other * stride];
        col[(size_t)i * stride] = b;
        col[(size_t)other * stride] = a;
    }
}
__global__ void __launch_bounds__(kThreads) k_pixel_tables(TableGen g) {
    const float kScale = 2.3283064365386963e-10f;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < g.n_pix; pix += gridDim.x * blockDim.x) {
        Pcg32 rng;
        rng.set_sequence(g.seq0 + pix);
        const uint32_t stride = g.n_pix;
        if (g.kind == 2) {
            const int n = g.xs * g.ys;
            const float inv_n = 1.0f / (float)n;
            for (int d = 0; d < g.n_dims; ++d) {                                           // stratified_sample_1d, sampling.rs:11-17
                float* col = g.t1 + (size_t)d * g.spp * stride + pix;
                for (int i = 0; i < n; ++i) {
                    const float delta = g.jitter ? rng.next_float() : 0.5f;
                    col[(size_t)i * stride] = fminf(PB2_ONE_MINUS_EPS, ((float)i + delta) * inv_n);
                }
                shuffle_column(col, stride, n, rng);
            }
            const float dx = 1.0f / (float)g.xs, dy = 1.0f / (float)g.ys;
            for (int d = 0; d < g.n_dims; ++d) {                                           // stratified_sample_2d, sampling.rs:19-41
                float2* col = g.t2 + (size_t)d * g.spp * stride + pix;
                int i = 0;
                for (int y = 0; y < g.ys; ++y)
                    for (int x = 0; x < g.xs; ++x) {
                        float jx = 0.5f, jy = 0.5f;
                        if (g.jitter) { jx = rng.next_float(); jy = rng.next_float(); }
                        col[(size_t)i * stride] = make_float2(fminf(PB2_ONE_MINUS_EPS, ((float)x + jx) * dx), fminf(PB2_ONE_MINUS_EPS, ((float)y + jy) * dy));
                        ++i;
                    }
                shuffle_column(col, stride, n, rng);
            }
        } else {
            for (int d = 0; d < g.n_dims; ++d) {                                           // van_der_corput, lowdiscrepancy.rs:436-460
                float* col = g.t1 + (size_t)d * g.spp * stride + pix;
                uint32_t v = rng.next_u32();
                for (int i = 0; i < g.spp; ++i) {                                          // gray_code_sample :416-422
                    col[(size_t)i * stride] = fminf(PB2_ONE_MINUS_EPS, __uint2float_rn(v) * kScale);
                    v ^= 0x80000000u >> (__ffs(i + 1) - 1);
                }
                for (int i = 0; i < g.spp; ++i) (void)rng.next_u32();                      // P1: shuffle of 1 element = uniform_u32_u32(1), one draw
                shuffle_column(col, stride, g.spp, rng);
            }
            for (int d = 0; d < g.n_dims; ++d) {                                           // sobol_2d, lowdiscrepancy.rs:462-505
                float2* col = g.t2 + (size_t)d * g.spp * stride + pix;
                uint32_t v0 = rng.next_u32(), v1 = rng.next_u32();
                for (int i = 0; i < g.spp; ++i) {                                          // gray_code_sample_2d :425-434
                    col[(size_t)i * stride] = make_float2(fminf(PB2_ONE_MINUS_EPS, __uint2float_rn(v0) * kScale), fminf(PB2_ONE_MINUS_EPS, __uint2float_rn(v1) * kScale));
                    const int tz = __ffs(i + 1) - 1;
                    v0 ^= 0x80000000u >> tz;
                    // column tz of the second Sobol' matrix: c[0] = 1 << 31, c[j] = c[j-1] ^ (c[j-1] >> 1) (P4)
                    uint32_t c = 0x80000000u;
                    for (int j = 0; j < tz; ++j) c ^= c >> 1;
                    v1 ^= c;
                }
                for (int i = 0; i < g.spp; ++i) (void)rng.next_u32();
                shuffle_column(col, stride, g.spp, rng);
            }
        }
    }
}

// ---- k_raygen: integrator.rs:431-445 + sampler.rs:27-33 -------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_raygen(uint64_t n, PathMap map, FilmView film, CameraView cam, PathBuffers b) {
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const SlotInfo si = slot_info(map, film, slot);
        PathSampler smp;
        smp.start(map.smp, si);
        float u0, u1, l0 = 0.0f, l1 = 0.0f;
        smp.film_offset(si, &u0, &u1);                                // p_film offset (x then y)
        if (smp.global() && !(cam.lens_radius > 0.0f)) smp.dim += 3u; // time, p_lens: drawn, never used by a pinhole
        else { (void)smp.next1(); smp.next2(&l0, &l1); }
        vec3 o, d;
        float t_max;
        camera_ray(cam, (float)si.x + u0, (float)si.y + u1, l0, l1, &o, &d, &t_max);
        b.ray_o[slot] = make_float4(o.x, o.y, o.z, t_max);
        b.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.0f);
        b.beta[slot] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        b.L[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(smp.extra() << 17));
        b.rng[slot] = smp.save();
        b.q_active[0][slot] = (uint32_t)slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.counters[C_ACTIVE_A] = n;
        b.counters[C_ACTIVE_B] = 0;
        b.counters[C_SHADOW] = b.counters[C_MIS] = 0;          // no NEE record precedes the first bounce
        b.counters[C_WORK_EXTEND] = b.counters[C_WORK_SHADOW] = 0;
        b.counters[T_CAMERA] += n;
        b.counters[T_EXTEND] += n;
    }
}

// ---- extend: closest hit over the active queue; hits are binned by material type (material-sorted shading) ---------------
// l += beta * (estimate_direct / light_pdf) (integrator.rs:178-191,243-261,:133, path.rs:117-120) for the NEE record of `slot`:
// ld = [unoccluded light sample] + [BSDF sample that reached the light], added in that order.  Runs inside the visibility
// kernels the moment the last pending query of the record is answered, so there is no separate resolve pass.
__device__ __forceinline__ void resolve_nee(const PathBuffers& b, uint32_t slot, bool add_t1, float4 t1, bool add_t2, float4 t2) {
    rgb3 ld = gray(0.0f);
    if (add_t1) ld = ld + mkc(t1.x, t1.y, t1.z);
    if (add_t2) ld = ld + mkc(t2.x, t2.y, t2.z);
    const float pick_pdf = b.sh_d[slot].w;
    const float4 bn = b.beta_nee[slot];
    float4 Lf = b.L[slot];
    const rgb3 L = mkc(Lf.x, Lf.y, Lf.z) + mkc(bn.x, bn.y, bn.z) * (ld / pick_pdf);
    b.L[slot] = make_float4(L.r, L.g, L.b, Lf.w);
}

// Work items [0, n_extend) are this bounce's path rays; [n_extend, n_extend + n_mis) are the MIS rays the previous bounce's
// estimate_direct left behind (BSDF-sampled rays towards the chosen light, integrator.rs:243-261): both are closest-hit
// walks, so they share the launch (a launch of their own cost 3-4 % of a frame for 0.1 % of the rays).  The shadow rays of
// that bounce were answered by k_shadow earlier in the stream, so an MIS ray finishing here completes its NEE record.
struct ExtendSink {
    PathBuffers b;
    ShadeView sh;
    const uint32_t* queue;
    uint32_t n_extend;
    PB2_D bool load(uint32_t i, vec3* o, vec3* d, float* t_max) const {
        const bool mis = i >= n_extend;
        const uint32_t slot = mis ? b.q_mis[i - n_extend] : queue[i];
        const float4 ro = mis ? b.mis_o[slot] : b.ray_o[slot], rd = mis ? b.mis_d[slot] : b.ray_d[slot];
        *o = mk(ro.x, ro.y, ro.z);
        *d = mk(rd.x, rd.y, rd.z);
        *t_max = mis ? kInf : ro.w;
        return true;
    }
    PB2_D void accept(uint32_t i, uint32_t prim, float t, float b0, float b1, float b2) const {
        (void)t;
        if (i >= n_extend) b.mis_prim[b.q_mis[i - n_extend]] = prim;
        else b.hit[queue[i]] = make_uint4(prim, __float_as_uint(b0), __float_as_uint(b1), __float_as_uint(b2));
    }
    PB2_D void accept_sphere(uint32_t i, uint32_t prim, float t, float u, float v) const {       // hit.y carries t for a sphere
        if (i >= n_extend) b.mis_prim[b.q_mis[i - n_extend]] = prim;
        else b.hit[queue[i]] = make_uint4(prim, __float_as_uint(t), __float_as_uint(u), __float_as_uint(v));
    }
    PB2_D void finish(uint32_t i, bool found, float) const {
        if (i >= n_extend) {
            const uint32_t slot = b.q_mis[i - n_extend];
            const float4 t1 = b.t1[slot], t2 = b.t2[slot];
            const bool reached_light = found && b.mis_prim[slot] == __float_as_uint(t2.w);    // D56 FIX: the closest hit is the light
            const bool lit = (__float_as_uint(t1.w) & 1u) && !b.occluded[slot];               // k_shadow ran before this kernel
            resolve_nee(b, slot, lit, t1, reached_light, t2);
            return;
        }
        const uint32_t slot = queue[i];
        if (!found) { b.state[slot] = kStateDead; return; }      // escaped: no infinite lights in scope, the path is finished
        const uint32_t prim = b.hit[slot].x;             // written by this thread's last accept()
        // shading class = the k_shade instantiation that handles the hit; compact_queues() bins the slots by it (material-sorted shading)
        b.state[slot] = (uint8_t)sh.mats[sh.tri_material[prim]].cls;
    }
    PB2_D void occluded(uint32_t, bool) const {}
};
__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_extend(SceneView s, ShadeView sh, PathBuffers b, int cur, TraceTuning tune) {
    const uint32_t n_extend = (uint32_t)b.counters[C_ACTIVE_A + cur], n_mis = (uint32_t)b.counters[C_MIS];   // the previous bounce's MIS rays
    const ExtendSink sink{b, sh, b.q_active[cur], n_extend};
    trace_persistent<false, false>(s, n_extend + n_mis, &b.counters[C_WORK_EXTEND], sink, tune);
}
// scenes with analytic spheres (sphere.cuh): separate kernels, as in kernels_traverse.cu
__global__ void __launch_bounds__(128, 4) k_extend_spheres(SceneView s, ShadeView sh, PathBuffers b, int cur, TraceTuning tune) {
    const uint32_t n_extend = (uint32_t)b.counters[C_ACTIVE_A + cur], n_mis = (uint32_t)b.counters[C_MIS];
    const ExtendSink sink{b, sh, b.q_active[cur], n_extend};
    trace_persistent<false, true>(s, n_extend + n_mis, &b.counters[C_WORK_EXTEND], sink, tune);
}

struct ShadowSink {
    PathBuffers b;
    PB2_D bool load(uint32_t i, vec3* o, vec3* d, float* t_max) const {
        const uint32_t slot = b.q_shadow[i];
        const float4 ro = b.sh_o[slot], rd = b.sh_d[slot];
        *o = mk(ro.x, ro.y, ro.z);
        *d = mk(rd.x, rd.y, rd.z);
        *t_max = ro.w;
        return true;
    }
    PB2_D void accept(uint32_t, uint32_t, float, float, float, float) const {}
    PB2_D void accept_sphere(uint32_t, uint32_t, float, float, float) const {}
    PB2_D void finish(uint32_t, bool, float) const {}
    PB2_D void occluded(uint32_t i, bool occ) const {
        const uint32_t slot = b.q_shadow[i];
        const float4 t1 = b.t1[slot];
        if (__float_as_uint(t1.w) == 1u) resolve_nee(b, slot, !occ, t1, false, t1);      // no MIS ray pending: done
        else b.occluded[slot] = occ ? 1 : 0;                                              // k_mis finishes the record
    }
};
__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_shadow(SceneView s, PathBuffers b, TraceTuning tune) {
    const ShadowSink sink{b};
    trace_persistent<true, false>(s, (uint32_t)b.counters[C_SHADOW], &b.counters[C_WORK_SHADOW], sink, tune);
}
__global__ void __launch_bounds__(128, 4) k_shadow_spheres(SceneView s, PathBuffers b, TraceTuning tune) {
    const ShadowSink sink{b};
    trace_persistent<true, true>(s, (uint32_t)b.counters[C_SHADOW], &b.counters[C_WORK_SHADOW], sink, tune);
}

// ---- shade -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ vec3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

struct Vertex {           // SurfaceInteraction subset rebuilt from the hit record (triangle.rs:193-311, D59)
    vec3 p, err, n, dpdu;
    vec3 sn, sdpdu;       // shading.n, shading.dpdu: n and dpdu unless the mesh has vertex normals / tangents
    vec3 wo;              // SurfaceInteraction::wo: -ray.d for a triangle; normalize(o2w * -ray_obj.d) for a sphere (sphere.rs:79)
};
// SG = the mesh carries per-vertex normals, tangents or UVs (compiled out otherwise: k_shade is at its register limit).
// ro / rd: the ray that hit (a sphere's interaction is rebuilt from the ray and the hit distance, which a sphere hit carries in
// place of b0; sphere.rs:38-93).
template <bool SG>
__device__ __forceinline__ Vertex rebuild_vertex(const SceneView& s, const ShadeView& sh, uint32_t prim, float b0, float b1, float b2,
                                                 vec3 ro = mk(0.f, 0.f, 0.f), vec3 rd = mk(0.f, 0.f, 1.f)) {
#ifndef PB2_TRIS_BY_PRIM
#define PB2_TRIS_BY_PRIM 1     /* 0: triangle through slot_of_prim (one more dependent fetch; tuning builds) */
#endif
    const float4* tp = PB2_TRIS_BY_PRIM ? s.tris_prim + 3ull * prim : s.tris + 3ull * __ldg(s.slot_of_prim + prim);
    const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
    const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
    Vertex v;
    v.wo = -rd;
    if (SG && s.spheres && (__float_as_uint(c.w) & 2u)) {
        const SphereVertex sv = sphere_vertex_at(sphere_of(s, a), ro, rd, b0);
        v.p = sv.p; v.err = sv.err; v.n = sv.n; v.dpdu = sv.dpdu; v.sn = sv.sn; v.sdpdu = sv.dpdu; v.wo = sv.wo;
        return v;
    }
    const float xs = (fabsf(b0 * p0.x) + fabsf(b1 * p1.x)) + fabsf(b2 * p2.x);
    const float ys = (fabsf(b0 * p0.y) + fabsf(b1 * p1.y)) + fabsf(b2 * p2.y);
    const float zs = (fabsf(b0 * p0.z) + fabsf(b1 * p1.z)) + fabsf(b2 * p2.z);
    v.err = mk(xs, ys, zs) * gammaf_(7.0f);
    v.p = (p0 * b0 + p1 * b1) + p2 * b2;
    v.n = unit(cross3(p0 - p2, p1 - p2));
    vec3 dv;
    if (!SG || !sh.indices) {                                  // (SG without mesh attributes: a scene that has analytic spheres)
        tri_frame(p0, p1, p2, &v.dpdu, &dv);
        v.sn = v.n;
        v.sdpdu = v.dpdu;
        return v;
    }
    const uint32_t i0 = __ldg(sh.indices + 3ull * prim), i1 = __ldg(sh.indices + 3ull * prim + 1), i2 = __ldg(sh.indices + 3ull * prim + 2);
    if (sh.uvs) tri_frame_uv(p0, p1, p2, __ldg(sh.uvs + i0), __ldg(sh.uvs + i1), __ldg(sh.uvs + i2), &v.dpdu, &dv);   // Triangle::get_uvs
    else tri_frame(p0, p1, p2, &v.dpdu, &dv);
    v.sn = v.n;
    v.sdpdu = v.dpdu;
    if (sh.normals || sh.tangents) {                                     // triangle.rs:251-311
        vec3 ns = v.n;
        if (sh.normals) {
            ns = (ld3(sh.normals + 3ull * i0) * b0 + ld3(sh.normals + 3ull * i1) * b1) + ld3(sh.normals + 3ull * i2) * b2;
            ns = len2(ns) > 0.0f ? unit(ns) : v.n;
        }
        vec3 ss = unit(v.dpdu);
        if (sh.tangents) {
            const vec3 st = (ld3(sh.tangents + 3ull * i0) * b0 + ld3(sh.tangents + 3ull * i1) * b1) + ld3(sh.tangents + 3ull * i2) * b2;
            if (len2(st) > 0.0f) ss = unit(st);
        }
        vec3 ts = cross3(ss, ns);
        if (len2(ts) > 0.0f) { ts = unit(ts); ss = cross3(ts, ns); }
        else coord_system(ns, &ss, &ts);
        // SurfaceInteraction::set_shading_geometry(ss, ts, .., true) (interaction.rs:297-316)
        v.sn = unit(cross3(ss, ts));
        v.n = face_toward(v.n, v.sn);                                    // D6 FIX
        v.sdpdu = ss;
    }
    return v;
}
// estimate_direct (integrator.rs:136-266) up to the two visibility queries: fills the NEE record of `slot`.
// out != nullptr (VolPathIntegrator, k_volpath): nothing is stored in the path buffers; the factors of both terms are handed back
// unmultiplied, because with handle_media the transmittance enters the product before f (integrator.rs:172-173, 259-261).
struct NeeOut {
    rgb3 li, f1;            // light sample: radiance and f (* |cos|)
    float w1, light_pdf;    // MIS weight (unused for a delta light) and the light's pdf
    bool delta;
    vec3 p1, p1_err, p1_n;  // the sampled point on the light (VisibilityTester's p1)
    rgb3 lmis, f2;          // BSDF / phase sample: the light's radiance towards the vertex and f (* |cos|)
    float w2, scattering_pdf;
    vec3 mis_o, mis_d;
    unsigned light_prim;
};
template <bool SG, class BsdfType>
__device__ __forceinline__ unsigned direct_lighting(const SceneView& s, const ShadeView& sh, const PathBuffers& b, uint32_t slot, const Vertex& v, vec3 wo,
                                                const BsdfType& bsdf, const DLight& light, float pick_pdf, float ul0, float ul1, float us0,
                                                float us1, rgb3 beta, NeeOut* out = nullptr) {
    const unsigned flags = kAllLobes & ~kSpecular;                       // D23 FIX
    const rgb3 l_emit = mkc(light.l[0], light.l[1], light.l[2]);
    const bool delta = light.type != 1;                                  // light.rs:28-31, D24 FIX: point, spot, distant
    vec3 wi = mk(0.f, 0.f, 0.f);
    float light_pdf = 0.0f, scattering_pdf = 0.0f;
    rgb3 li = gray(0.0f);
    vec3 sh_o = mk(0.f, 0.f, 0.f), sh_d = mk(0.f, 0.f, 0.f);
    const vec3 lp0 = ld3(light.p0), lp1 = ld3(light.p1), lp2 = ld3(light.p2);
    if (delta) {                                                         // point.rs:47-66, spot.rs:71-85, distant.rs:50-67
        vec3 pl = ld3(light.p);
        light_pdf = 1.0f;
        if (light.type == 3) {                                           // DistantLight: the tester's far end is p_outside
            wi = ld3(light.axis);
            pl = v.p + wi * (2.0f * light.world_radius);
            li = l_emit;
        } else {
            wi = unit(pl - v.p);
            if (light.type == 2) {                                       // SpotLight::falloff(-wi), spot.rs:51-63
                const vec3 w = -wi;
                const float cos_theta = (light.axis[0] * w.x + light.axis[1] * w.y) + light.axis[2] * w.z;
                float fall = 1.0f;
                if (cos_theta < light.cos_total_width) fall = 0.0f;
                else if (!(cos_theta >= light.cos_falloff_start)) {
                    const float dl = (cos_theta - light.cos_total_width) / (light.cos_falloff_start - light.cos_total_width);
                    fall = (dl * dl) * (dl * dl);
                }
                li = l_emit * fall / len2(pl - v.p);
            } else li = l_emit / len2(pl - v.p);
        }
        sh_o = offset_ray_origin(v.p, v.err, v.n, pl - v.p);             // interaction.rs:146-153
        const vec3 target = offset_ray_origin(pl, mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f), sh_o - pl);
        sh_d = target - sh_o;
        if (out) { out->p1 = pl; out->p1_err = mk(0.f, 0.f, 0.f); out->p1_n = mk(0.f, 0.f, 0.f); }
    } else if (SG && light.sphere >= 0) {                                // diffuse.rs:60-81 + sphere.rs:127-193
        vec3 ps, pe, ns;
        float pdf;
        sphere_sample2(reinterpret_cast<const DSphere*>(s.spheres)[light.sphere], v.p, v.err, v.n, ul0, ul1, &ps, &pe, &ns, &pdf);
        if (pdf == 0.0f || len2(ps - v.p) == 0.0f) { light_pdf = 0.0f; }
        else {
            light_pdf = pdf;
            wi = unit(ps - v.p);
            li = (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);
            sh_o = offset_ray_origin(v.p, v.err, v.n, ps - v.p);
            const vec3 target = offset_ray_origin(ps, pe, ns, sh_o - ps);
            sh_d = target - sh_o;
            if (out) { out->p1 = ps; out->p1_err = pe; out->p1_n = ns; }
        }
    } else {                                                             // diffuse.rs:60-81, shape.rs:38-53, triangle.rs:330-348
        const float su0 = sqrtf(ul0);
        const float b0 = 1.0f - su0, b1 = ul1 * su0;                     // sampling.rs:275-278
        const float b2 = (1.0f - b0) - b1;
        const vec3 ps = (lp0 * b0 + lp1 * b1) + lp2 * b2;
        vec3 ns = unit(cross3(lp1 - lp0, lp2 - lp0));
        if (light.has_n) ns = face_toward(ns, (ld3(light.n0) * b0 + ld3(light.n1) * b1) + ld3(light.n2) * b2);     // triangle.rs:338-341, D6 FIX
        const vec3 pe = ((abs3(lp0 * b0) + abs3(lp1 * b1)) + abs3(lp2 * b2)) * gammaf_(6.0f);
        float pdf = 1.0f / light.area;
        vec3 w = ps - v.p;
        if (len2(w) == 0.0f) pdf = 0.0f;
        else {
            w = unit(w);
            pdf = pdf * (len2(v.p - ps) / fabsf(dot3(ns, -w)));
            if (isinf(pdf)) pdf = 0.0f;
        }
        if (pdf == 0.0f || len2(ps - v.p) == 0.0f) { light_pdf = 0.0f; }
        else {
            light_pdf = pdf;
            wi = unit(ps - v.p);
            li = (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);      // D55 FIX
            sh_o = offset_ray_origin(v.p, v.err, v.n, ps - v.p);
            const vec3 target = offset_ray_origin(ps, pe, ns, sh_o - ps);
            sh_d = target - sh_o;
            if (out) { out->p1 = ps; out->p1_err = pe; out->p1_n = ns; }
        }
    }
    unsigned pending = 0u;
    rgb3 t1 = gray(0.0f), t2 = gray(0.0f);
    if (light_pdf > 0.0f && !black(li)) {
        scattering_pdf = bsdf_pdf(bsdf, wo, wi, flags);
        const rgb3 f = bsdf_f(bsdf, wo, wi, flags) * cos_factor(bsdf, wi);
        if (!black(f)) {
            pending |= 1u;                                               // VisibilityTester::un_occluded decides (D25 FIX)
            if (out) { out->li = li; out->f1 = f; out->light_pdf = light_pdf; out->delta = delta; out->w1 = delta ? 1.0f : power_heuristic(light_pdf, scattering_pdf); }
            else t1 = delta ? li * f / light_pdf : li * f * power_heuristic(light_pdf, scattering_pdf) / light_pdf;
        }
    }
    vec3 mis_o = mk(0.f, 0.f, 0.f), mis_d = mk(0.f, 0.f, 1.f);
    if (!delta) {
        unsigned sampled = 0u;
        rgb3 f = bsdf_sample_f(bsdf, wo, &wi, us0, us1, &scattering_pdf, flags, &sampled);
        f = f * cos_factor(bsdf, wi);
        const bool sampled_specular = (sampled & kSpecular) != 0u;
        if (!black(f) && scattering_pdf > 0.0f) {
            float weight = 1.0f;
            bool go = true;
            const vec3 ro = offset_ray_origin(v.p, v.err, v.n, wi);      // it.spawn_ray(wi)
            float lb0 = 0.0f, lb1 = 0.0f, lb2 = 0.0f;
            if (SG && light.sphere >= 0) {
                // Light::pdf_li -> Sphere::pdf2 (sphere.rs:195-207); then the light's normal where this ray meets the sphere (the
                // closest hit of the MIS ray is this sphere or the contribution is dropped in k_extend)
                const DSphere& sp = reinterpret_cast<const DSphere*>(s.spheres)[light.sphere];
                if (!sampled_specular) {
                    const float lp = sphere_pdf2(sp, v.p, v.err, v.n, wi);
                    if (lp == 0.0f) go = false;
                    else weight = power_heuristic(scattering_pdf, lp);
                }
                if (go) {
                    float t_l;
                    SphereVertex lv;
                    if (!sphere_intersect(sp, ro, wi, kInf, &t_l, &lv)) go = false;
                    else {
                        const rgb3 lmis = (light.two_sided || dot3(lv.n, -wi) > 0.0f) ? l_emit : gray(0.0f);
                        if (!black(lmis)) {
                            pending |= 2u;
                            if (out) { out->lmis = lmis; out->f2 = f; out->w2 = weight; out->scattering_pdf = scattering_pdf; }
                            else t2 = lmis * f * gray(1.0f) * weight / scattering_pdf;
                            mis_o = ro;
                            mis_d = wi;
                        }
                    }
                }
                go = false;                                              // handled
            } else
            if (!sampled_specular) {
                // Light::pdf_li -> Shape::pdf2 (shape.rs:54-69): the light's own triangle
                const RayCtx rc = make_ray_ctx(ro, wi);
                float t;
                vec3 du, dv;
                const bool frame_ok = light.has_uv ? tri_frame_uv(lp0, lp1, lp2, make_float2(light.uv[0], light.uv[1]), make_float2(light.uv[2], light.uv[3]),
                                                                   make_float2(light.uv[4], light.uv[5]), &du, &dv)
                                                   : tri_frame(lp0, lp1, lp2, &du, &dv);
                if (!tri_test(rc, kInf, lp0, lp1, lp2, &t, &lb0, &lb1, &lb2) || !frame_ok) go = false;
                else {
                    const vec3 p_l = (lp0 * lb0 + lp1 * lb1) + lp2 * lb2;
                    const vec3 n_l = unit(cross3(lp0 - lp2, lp1 - lp2));
                    float lp = len2(v.p - p_l) / (fabsf(dot3(n_l, -wi)) * light.area);
                    if (isinf(lp)) lp = 0.0f;
                    if (lp == 0.0f) go = false;
                    else weight = power_heuristic(scattering_pdf, lp);
                }
            }
            if (go) {
                // li = light_isect.le(-wi) if the closest hit is this light's triangle (D56 FIX); its normal is known here:
                // the geometric one, or — on a mesh with vertex normals / tangents — the one Triangle::intersect leaves in the
                // interaction at these barycentrics (flipped towards the shading normal, set_shading_geometry)
                vec3 n_l = unit(cross3(lp0 - lp2, lp1 - lp2));
                if (SG && !sampled_specular) n_l = rebuild_vertex<true>(s, sh, light.prim, lb0, lb1, lb2).n;
                const rgb3 lmis = (light.two_sided || dot3(n_l, -wi) > 0.0f) ? l_emit : gray(0.0f);
                if (!black(lmis)) {
                    pending |= 2u;
                    if (out) { out->lmis = lmis; out->f2 = f; out->w2 = weight; out->scattering_pdf = scattering_pdf; }
                    else t2 = lmis * f * gray(1.0f) * weight / scattering_pdf;
                    mis_o = ro;
                    mis_d = wi;
                }
            }
        }
    }
    if (pending == 0u) return 0u;
    if (out) {
        out->mis_o = mis_o;
        out->mis_d = mis_d;
        out->light_prim = light.prim;
        return pending;
    }
    b.sh_o[slot] = make_float4(sh_o.x, sh_o.y, sh_o.z, 1.0f - PB2_SHADOW_EPS);
    b.sh_d[slot] = make_float4(sh_d.x, sh_d.y, sh_d.z, pick_pdf);
    b.t1[slot] = make_float4(t1.r, t1.g, t1.b, __uint_as_float(pending));
    b.beta_nee[slot] = make_float4(beta.r, beta.g, beta.b, 0.0f);
    if (pending & 2u) {
        b.mis_o[slot] = make_float4(mis_o.x, mis_o.y, mis_o.z, 0.0f);
        b.mis_d[slot] = make_float4(mis_d.x, mis_d.y, mis_d.z, 0.0f);
        b.t2[slot] = make_float4(t2.r, t2.g, t2.b, __uint_as_float(light.prim));
    }
    (void)s;
    return pending;                                  // bit 0: shadow ray, bit 1: MIS ray — queued by compact_queues()
}

// One path vertex of PathIntegrator::li (path.rs:79-209) for every hit of material type `mat`.
#ifndef PB2_SHADE_BLOCKS
#define PB2_SHADE_BLOCKS 2
#endif
#ifndef PB2_SHADE_THREADS
#define PB2_SHADE_THREADS 256
#endif
template <int MAT, bool TABLES, bool SG>
__global__ void __launch_bounds__(PB2_SHADE_THREADS, PB2_SHADE_BLOCKS) k_shade(SceneView s, ShadeView sh, PathBuffers b, PathMap map, FilmView film, PathParams pp, int cur) {
    const uint64_t n = b.counters[C_MAT0 + MAT];
    const uint32_t* queue = b.q_mat[MAT];
#ifndef PB2_SHADE_PIPE
#define PB2_SHADE_PIPE 1        /* 0: load the queue entry and the hit record where they are used (the round-1 loop; tuning builds) */
#endif
    // The head of an iteration is a chain of dependent loads — queue entry -> hit record -> leaf slot -> triangle — that held
    // 22 % of the kernel's stall samples on its first two links alone (profiles/r02_tuning.md); they are issued two / one
    // iteration ahead, so their latency runs under the previous vertices' arithmetic.
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_n = 0, slot_nn = 0;
    uint4 h_n = make_uint4(0u, 0u, 0u, 0u);
    if (PB2_SHADE_PIPE) {
        if (i0 < n) { slot_n = queue[i0]; h_n = b.hit[slot_n]; }
        if (i0 + stride < n) slot_nn = queue[i0 + stride];
    }
    for (uint64_t i = i0; i < n; i += stride) {
        uint32_t slot;
        uint4 h;
        if (PB2_SHADE_PIPE) {
            slot = slot_n;
            h = h_n;
            slot_n = slot_nn;
            if (i + stride < n) h_n = b.hit[slot_n];
            if (i + 2 * stride < n) slot_nn = queue[i + 2 * stride];
        } else {
            slot = queue[i];
            h = b.hit[slot];
        }
        const float4 rd = b.ray_d[slot];
        float4 Lf = b.L[slot];
        float4 bt = b.beta[slot];
        rgb3 L = mkc(Lf.x, Lf.y, Lf.z), beta = mkc(bt.x, bt.y, bt.z);
        float eta_scale = bt.w;
        const unsigned state = __float_as_uint(Lf.w);
        unsigned bounces = state & 0xFFFFu;
        const bool specular_bounce = (state >> 16) & 1u;
        vec3 ray_o = mk(0.f, 0.f, 0.f);
        if (SG && s.spheres) { const float4 ro = b.ray_o[slot]; ray_o = mk(ro.x, ro.y, ro.z); }
        const Vertex v = rebuild_vertex<SG>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), ray_o, mk(rd.x, rd.y, rd.z));
        const vec3 wo = -mk(rd.x, rd.y, rd.z);
        if (bounces == 0u || specular_bounce) {                          // path.rs:80-82 + interaction.rs:387-395
            const int li = sh.tri_light[h.x];
            if (li >= 0) {
                const DLight& lt = sh.lights[li];
                const rgb3 le = (lt.two_sided || dot3(v.n, wo) > 0.0f) ? mkc(lt.l[0], lt.l[1], lt.l[2]) : gray(0.0f);
                L = L + beta * le;
            }
        }
        bool alive = bounces < (unsigned)pp.max_depth;                   // path.rs:90-92
        unsigned queued = (unsigned)MAT;                                 // b.state[slot]: class | continues << 2 | NEE rays << 3
        if (alive) {
            const auto bsdf = make_bsdf<MAT>(sh.mats[sh.tri_material[h.x]], v.n, v.sn, v.sdpdu);
            PathSampler rng;
            rng.resume(map.smp, slot_info(map, film, slot), b.rng[slot], TABLES ? (state >> 17) & 0x3FFFu : 0u);
            // (class 2 holds specular lobes only — FresnelSpecular, SpecularReflection — so estimate_direct is never reached there)
            if (MAT != 2 && bsdf_count(bsdf, kAllLobes & ~kSpecular) > 0 && sh.n_lights > 0) {      // path.rs:105-121, integrator.rs:99-134
                float pick_pdf;
                // light_distribution.lookup(&isect.p) (path.rs:100-104)
                const float *l_cdf = sh.light_cdf, *l_func = sh.light_func;
                float l_int = sh.light_func_int;
                if (sh.spatial.func) {
                    const size_t vox = spatial_voxel(sh.spatial, v.p);
                    l_cdf = sh.spatial.cdf + vox * (size_t)(sh.n_lights + 1);
                    l_func = sh.spatial.func + vox * (size_t)sh.n_lights;
                    l_int = __ldg(sh.spatial.func_int + vox);
                }
                const int li = sample_discrete(l_cdf, l_func, sh.n_lights, l_int, rng.next1<TABLES>(), &pick_pdf);
                if (pick_pdf != 0.0f) {
                    float ul0, ul1, us0, us1;
                    rng.next2<TABLES>(&ul0, &ul1);
                    rng.next2<TABLES>(&us0, &us1);
                    queued |= direct_lighting<SG>(s, sh, b, slot, v, SG ? v.wo : wo, bsdf, sh.lights[li], pick_pdf, ul0, ul1, us0, us1, beta) << 3;   // estimate_direct reads it.wo
                }
            }
            float u0, u1;
            rng.next2<TABLES>(&u0, &u1);                                                  // path.rs:123-134
            vec3 wi = mk(0.f, 0.f, 0.f);
            float pdf = 0.0f;
            unsigned sampled = 0u;
            const rgb3 f = bsdf_sample_f(bsdf, wo, &wi, u0, u1, &pdf, kAllLobes, &sampled);
            if (black(f) || pdf == 0.0f) alive = false;
            else {
                beta = beta * (f * (fabsf(dot3(wi, bsdf.ns)) / pdf));
                const bool spec = (sampled & kSpecular) != 0u;
                if (spec && (sampled & kTransmission)) {
                    const float eta = bsdf.eta;
                    eta_scale = eta_scale * ((dot3(wo, v.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta));
                }
                const vec3 o = offset_ray_origin(v.p, v.err, v.n, wi);
                const rgb3 rr_beta = beta * eta_scale;                                   // path.rs:200-207, D27 KEEP
                if (max_channel(rr_beta) < pp.rr_threshold && bounces > 3u) {
                    const float q = fminf(1.0f - max_channel(rr_beta), 0.05f);
                    if (rng.next1<TABLES>() < q) alive = false;
                    else beta = beta / (1.0f - q);
                }
                if (alive) {
                    bounces += 1u;
                    b.ray_o[slot] = make_float4(o.x, o.y, o.z, kInf);
                    b.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, 0.0f);
                    b.beta[slot] = make_float4(beta.r, beta.g, beta.b, eta_scale);
                    b.rng[slot] = rng.save();
                    Lf.w = __uint_as_float(bounces | ((spec ? 1u : 0u) << 16) | (TABLES ? rng.extra() << 17 : 0u));
                    queued |= kStateContinues;
                }
            }
        }
        b.state[slot] = (uint8_t)queued;
        b.L[slot] = make_float4(L.r, L.g, L.b, Lf.w);
    }
}

// ---- VolPathIntegrator (src/integrators/volpath.rs) over HomogeneousMedium (src/media/homogeneous.rs) ----------------------
// One thread carries one camera sample through the whole of VolPathIntegrator::li (volpath.rs:60-244): medium sampling, the
// phase-function or BSDF vertex, next-event estimation with transmittance (VisibilityTester::tr, light.rs:137-160) and the
// MIS ray through Scene::intersect_tr (scene.rs:48-71).  Every ray is a closest-hit walk (transmittance rays pass through
// material-less interface surfaces segment by segment), done here with the literal one-level walk of traverse.cuh; the
// wavefront stages of the PathIntegrator are not involved.  Defect ledger D69-D74: spawned rays take
// GetMedium(d), a material-less surface is not a bounce, tr / sample use min and the exponential.
struct VolHit {
    HitRec h;
    Vertex v;
};
__device__ __forceinline__ bool vol_intersect(const SceneView& s, const ShadeView& sh, vec3 o, vec3 d, float t_max, VolHit* out) {
    if (!traverse<false, true>(s, o, d, t_max, &out->h)) return false;
    out->v = rebuild_vertex<true>(s, sh, out->h.prim, out->h.sphere ? out->h.t : out->h.b0, out->h.b1, out->h.b2, o, d);
    return true;
}
// homogeneous.rs:36-38 (D71)
__device__ __forceinline__ rgb3 medium_tr(const DMedium& m, float t_max, vec3 d) {
    const float sdist = fminf(t_max * len(d), 3.402823466e+38f);
    return mkc(det_exp(-(m.sigma_t[0] * sdist)), det_exp(-(m.sigma_t[1] * sdist)), det_exp(-(m.sigma_t[2] * sdist)));
}
// primitive.rs:72-76
__device__ __forceinline__ void hit_interface(const ShadeView& sh, uint32_t prim, int ray_medium, int* inside, int* outside) {
    const int pi = sh.prim_inside ? sh.prim_inside[prim] : -1, po = sh.prim_outside ? sh.prim_outside[prim] : -1;
    if (pi != po) { *inside = pi; *outside = po; }
    else { *inside = ray_medium; *outside = ray_medium; }
}
// VisibilityTester::tr (light.rs:137-160, D74)
__device__ __forceinline__ rgb3 visibility_tr(const SceneView& s, const ShadeView& sh, vec3 p, vec3 err, vec3 n, int med_in, int med_out, vec3 p1,
                                              vec3 p1_err, vec3 p1_n, unsigned long long* n_rays) {
    rgb3 tr = gray(1.0f);
    for (;;) {
        const vec3 origin = offset_ray_origin(p, err, n, p1 - p);
        const vec3 target = offset_ray_origin(p1, p1_err, p1_n, origin - p1);
        const vec3 d = target - origin;
        const int medium = dot3(d, n) > 0.0f ? med_out : med_in;
        VolHit hit;
        ++*n_rays;
        const bool found = vol_intersect(s, sh, origin, d, 1.0f - PB2_SHADOW_EPS, &hit);
        if (found && sh.tri_material[hit.h.prim] != 0xFFFFFFFFu) return gray(0.0f);
        if (medium >= 0) tr = tr * medium_tr(sh.media[medium], found ? hit.h.t : 1.0f - PB2_SHADOW_EPS, d);
        if (!found) break;
        hit_interface(sh, hit.h.prim, medium, &med_in, &med_out);
        p = hit.v.p; err = hit.v.err; n = hit.v.n;
    }
    return tr;
}
// Scene::intersect_tr (scene.rs:48-71, D74): true when the ray ends on a surface with a material (its primitive in *prim)
__device__ __forceinline__ bool intersect_tr(const SceneView& s, const ShadeView& sh, vec3 o, vec3 d, int medium, uint32_t* prim, rgb3* tr,
                                             unsigned long long* n_rays) {
    *tr = gray(1.0f);
    for (;;) {
        VolHit hit;
        ++*n_rays;
        const bool found = vol_intersect(s, sh, o, d, kInf, &hit);
        if (medium >= 0) *tr = *tr * medium_tr(sh.media[medium], found ? hit.h.t : kInf, d);
        if (!found) return false;
        if (sh.tri_material[hit.h.prim] != 0xFFFFFFFFu) { *prim = hit.h.prim; return true; }
        int in, out;
        hit_interface(sh, hit.h.prim, medium, &in, &out);
        o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, d);
        medium = dot3(d, hit.v.n) > 0.0f ? out : in;
    }
}
// uniform_sample_one_light + estimate_direct with handle_media (integrator.rs:92-266) at a surface or medium vertex
template <class BsdfType>
__device__ __forceinline__ rgb3 vol_sample_one_light(const SceneView& s, const ShadeView& sh, const PathBuffers& b, const Vertex& v, vec3 wo_si,
                                                     const BsdfType& bsdf, int med_in, int med_out, PathSampler& smp, unsigned long long* n_shadow,
                                                     unsigned long long* n_mis) {
    if (sh.n_lights <= 0) return gray(0.0f);
    float pick_pdf;
    const float *l_cdf = sh.light_cdf, *l_func = sh.light_func;
    float l_int = sh.light_func_int;
    if (sh.spatial.func) {
        const size_t vox = spatial_voxel(sh.spatial, v.p);
        l_cdf = sh.spatial.cdf + vox * (size_t)(sh.n_lights + 1);
        l_func = sh.spatial.func + vox * (size_t)sh.n_lights;
        l_int = __ldg(sh.spatial.func_int + vox);
    }
    const int li_idx = sample_discrete(l_cdf, l_func, sh.n_lights, l_int, smp.next1(), &pick_pdf);
    if (pick_pdf == 0.0f) return gray(0.0f);
    float ul0, ul1, us0, us1;
    smp.next2(&ul0, &ul1);
    smp.next2(&us0, &us1);
    NeeOut ne;
    const unsigned pending = direct_lighting<true>(s, sh, b, 0u, v, wo_si, bsdf, sh.lights[li_idx], pick_pdf, ul0, ul1, us0, us1, gray(1.0f), &ne);
    rgb3 ld = gray(0.0f);
    if (pending & 1u) {
        const rgb3 li = ne.li * visibility_tr(s, sh, v.p, v.err, v.n, med_in, med_out, ne.p1, ne.p1_err, ne.p1_n, n_shadow);
        if (!black(li)) ld = ld + (ne.delta ? li * ne.f1 / ne.light_pdf : li * ne.f1 * ne.w1 / ne.light_pdf);
    }
    if (pending & 2u) {
        uint32_t prim;
        rgb3 tr;
        const int medium = dot3(ne.mis_d, v.n) > 0.0f ? med_out : med_in;
        if (intersect_tr(s, sh, ne.mis_o, ne.mis_d, medium, &prim, &tr, n_mis) && prim == ne.light_prim)
            ld = ld + ne.lmis * ne.f2 * tr * ne.w2 / ne.scattering_pdf;
    }
    return ld / pick_pdf;
}
// The surface vertex of volpath.rs:133-187 for shading class CLS: NEE, then the BSDF sample that continues the path.
// Returns false when the path ends (black f or zero pdf).
template <int CLS>
__device__ __forceinline__ bool vol_surface(const SceneView& s, const ShadeView& sh, const PathBuffers& b, const VolHit& hit, vec3 ray_d, int med_in,
                                            int med_out, PathSampler& smp, rgb3* L, rgb3* beta, float* eta_scale, bool* specular, vec3* wi_out,
                                            unsigned long long* n_shadow, unsigned long long* n_mis) {
    const Vertex& v = hit.v;
    const auto bsdf = make_bsdf<CLS>(sh.mats[sh.tri_material[hit.h.prim]], v.n, v.sn, v.sdpdu);
    *L = *L + *beta * vol_sample_one_light(s, sh, b, v, v.wo, bsdf, med_in, med_out, smp, n_shadow, n_mis);      // at every surface vertex (volpath.rs:137-146)
    const vec3 wo = -ray_d;
    float u0, u1, pdf = 0.0f;
    smp.next2(&u0, &u1);
    unsigned sampled = 0u;
    vec3 wi = mk(0.f, 0.f, 0.f);
    const rgb3 f = bsdf_sample_f(bsdf, wo, &wi, u0, u1, &pdf, kAllLobes, &sampled);
    if (black(f) || pdf == 0.0f) return false;
    *beta = *beta * (f * (fabsf(dot3(wi, bsdf.ns)) / pdf));
    *specular = (sampled & kSpecular) != 0u;
    if ((sampled & kSpecular) && (sampled & kTransmission)) {
        const float eta = bsdf.eta;
        *eta_scale = *eta_scale * ((dot3(wo, v.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta));
    }
    *wi_out = wi;
    return true;
}
__global__ void __launch_bounds__(128) k_volpath(uint64_t n, SceneView s, ShadeView sh, PathBuffers b, PathMap map, FilmView film, CameraView cam,
                                                 PathParams pp) {
    unsigned long long n_extend = 0, n_shadow = 0, n_mis = 0;
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const SlotInfo si = slot_info(map, film, slot);
        PathSampler smp;
        smp.start(map.smp, si);
        float u0, u1, l0 = 0.0f, l1 = 0.0f;
        smp.film_offset(si, &u0, &u1);
        if (smp.global() && !(cam.lens_radius > 0.0f)) smp.dim += 3u;
        else { (void)smp.next1(); smp.next2(&l0, &l1); }
        vec3 o, d;
        float t_max;
        camera_ray(cam, (float)si.x + u0, (float)si.y + u1, l0, l1, &o, &d, &t_max);
        int ray_medium = sh.camera_medium;
        rgb3 L = gray(0.0f), beta = gray(1.0f);
        bool specular_bounce = false;
        int bounces = 0;
        float eta_scale = 1.0f;
        for (;;) {
            VolHit hit;
            ++n_extend;
            const bool found = vol_intersect(s, sh, o, d, t_max, &hit);
            bool in_medium = false;
            vec3 mp = mk(0.f, 0.f, 0.f);
            if (ray_medium >= 0) {                                       // HomogeneousMedium::sample (homogeneous.rs:40-74; D72, D73)
                const DMedium& m = sh.media[ray_medium];
                const float ray_t_max = found ? hit.h.t : t_max;
                const float uc = smp.next1() * 3.0f;
                const int channel = min(__float2int_rz(uc), 2);
                const float dist = -det_log(1.0f - smp.next1()) / m.sigma_t[channel < 0 ? 0 : channel];
                const float dl = len(d);
                const float t = fminf(dist / dl, ray_t_max);
                in_medium = t < ray_t_max;
                if (in_medium) mp = o + d * t;
                const float tt = fminf(t, 3.402823466e+38f);
                const rgb3 tr = mkc(det_exp(-m.sigma_t[0] * tt * dl), det_exp(-m.sigma_t[1] * tt * dl), det_exp(-m.sigma_t[2] * tt * dl));
                const rgb3 density = in_medium ? mkc(m.sigma_t[0] * tr.r, m.sigma_t[1] * tr.g, m.sigma_t[2] * tr.b) : tr;
                float pdf = 0.0f;
                pdf += density.r; pdf += density.g; pdf += density.b;
                pdf *= 1.0f / 3.0f;
                if (pdf == 0.0f) pdf = 1.0f;
                beta = beta * (in_medium ? (tr * mkc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2])) / pdf : tr / pdf);
            }
            if (black(beta)) break;
            if (in_medium) {
                if (bounces >= pp.max_depth) break;
                const DMedium& m = sh.media[ray_medium];
                const vec3 wo = -d;
                vec3 wi = mk(0.f, 0.f, 0.f);
                float p0, p1;
                smp.next2(&p0, &p1);
                hg_sample_p(m.g, wo, &wi, p0, p1);                       // volpath.rs:96-103: sampled before the light (KEEP)
                Vertex v;
                v.p = mp; v.err = mk(0.f, 0.f, 0.f); v.n = mk(0.f, 0.f, 0.f); v.dpdu = mk(0.f, 0.f, 0.f); v.sn = v.n; v.sdpdu = v.dpdu; v.wo = wo;
                const PhaseHG ph{m.g, mk(0.f, 0.f, 0.f)};
                o = mp;                                                  // mi.spawn_ray(wi): no normal, no offset; the medium stays
                d = wi;
                t_max = kInf;
                specular_bounce = false;
                L = L + beta * vol_sample_one_light(s, sh, b, v, wo, ph, ray_medium, ray_medium, smp, &n_shadow, &n_mis);
            } else {
                if (bounces == 0 || specular_bounce) {
                    if (found) {
                        const int li = sh.tri_light[hit.h.prim];
                        if (li >= 0) {
                            const DLight& lt = sh.lights[li];
                            const rgb3 le = (lt.two_sided || dot3(hit.v.n, -d) > 0.0f) ? mkc(lt.l[0], lt.l[1], lt.l[2]) : gray(0.0f);
                            L = L + beta * le;
                        }
                    }
                }
                if (!found || bounces >= pp.max_depth) break;
                int in, out;
                hit_interface(sh, hit.h.prim, ray_medium, &in, &out);
                const uint32_t mat = sh.tri_material[hit.h.prim];
                if (mat == 0xFFFFFFFFu) {                                // volpath.rs:127-131 (D70): crosses the interface, not a bounce
                    o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, d);
                    t_max = kInf;
                    ray_medium = dot3(d, hit.v.n) > 0.0f ? out : in;
                    continue;
                }
                vec3 wi;
                bool go;
                const int cls = sh.mats[mat].cls;
                if (cls == 0) go = vol_surface<0>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                else if (cls == 1) go = vol_surface<1>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                else go = vol_surface<2>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                if (!go) break;
                o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, wi);
                d = wi;
                t_max = kInf;
                ray_medium = dot3(wi, hit.v.n) > 0.0f ? out : in;
            }
            const rgb3 rr_beta = beta * eta_scale;
            if (max_channel(rr_beta) < pp.rr_threshold && bounces > 3) {
                const float q = fmaxf(1.0f - max_channel(rr_beta), 0.05f);    // volpath.rs:236
                if (smp.next1() < q) break;
                beta = beta / (1.0f - q);
            }
            bounces += 1;
        }
        b.L[slot] = make_float4(L.r, L.g, L.b, 0.0f);
    }
    // ray totals of the launch (pb2_render_counters): one atomic per warp and counter
    for (int off = 16; off > 0; off >>= 1) {
        n_extend += __shfl_down_sync(0xFFFFFFFFu, n_extend, off);
        n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, off);
        n_mis += __shfl_down_sync(0xFFFFFFFFu, n_mis, off);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&b.counters[T_EXTEND], n_extend);
        atomicAdd(&b.counters[T_SHADOW], n_shadow);
        atomicAdd(&b.counters[T_MIS], n_mis);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&b.counters[T_CAMERA], (unsigned long long)n);
}

// ---- Film ------------------------------------------------------------------------------------------------------------------
// FilmTile::add_sample footprint (film.rs:259-294 with D43/D44): calls fn(px, py, filter_weight).
template <class F>
__device__ __forceinline__ void film_footprint(const FilmView& f, float pfx, float pfy, F&& fn) {
    const float dx = pfx - 0.5f, dy = pfy - 0.5f;
    int x0 = (int)ceilf(dx - f.radius_x), y0 = (int)ceilf(dy - f.radius_y);
    int x1 = (int)floorf(dx + f.radius_x) + 1, y1 = (int)floorf(dy + f.radius_y) + 1;
    x0 = max(x0, f.px0); y0 = max(y0, f.py0);
    x1 = min(x1, f.px1); y1 = min(y1, f.py1);
    const float inv_rx = 1.0f / f.radius_x, inv_ry = 1.0f / f.radius_y;
    for (int y = y0; y < y1; ++y) {
        const int iy = min(15, (int)floorf(fabsf(((float)y - dy) * inv_ry * 16.0f)));
        for (int x = x0; x < x1; ++x) {
            const int ix = min(15, (int)floorf(fabsf(((float)x - dx) * inv_rx * 16.0f)));
            fn(x, y, f.table[iy * 16 + ix]);
        }
    }
}
__device__ __forceinline__ rgb3 guard_radiance(rgb3 L) {                  // integrator.rs:455-457, D22 FIX
    if (any_nan(L) || luminance(L) < -1e-5f || isinf(luminance(L))) return gray(0.0f);
    return L;
}
__device__ __forceinline__ rgb3 clamp_luminance(const FilmView& f, rgb3 L) {   // FilmTile::add_sample, film.rs:259-261
    const float y = luminance(L);
    if (y > f.max_lum) L = L * (f.max_lum / y);
    return L;
}
__device__ __forceinline__ void film_atomic_add(const FilmView& f, int px, int py, rgb3 c, float w) {
    float* a = reinterpret_cast<float*>(f.acc + f.index(px, py));
    atomicAdd(a, c.r); atomicAdd(a + 1, c.g); atomicAdd(a + 2, c.b); atomicAdd(a + 3, w);
}
// A stray = the contribution of a sample of pixel (sx, sy) to another pixel (px, py) (exact mode: box filter, r = 0.5, so the
// target is one of the eight neighbours).  Strays of one target are applied in the order (source pixel, sample index); the
// sort key is {target pixel index : 32 | which neighbour the source is, in row-major order : 4 | sample index : 28}, which
// orders the sources of one target exactly as their sample-bounds pixel indices do.  (A 24-bit pixel field, as this key
// had first, wrapped on films of more than 2^24 pixels.)
__device__ __forceinline__ void film_stray(const FilmView& f, unsigned long long* counters, int sx, int sy, uint32_t sample, int px, int py, rgb3 c, float w) {
    const unsigned long long pos = atomicAdd(&counters[C_STRAYS], 1ull);
    if (pos < f.stray_capacity) {
        const unsigned nb = (unsigned)((sy - py + 1) * 3 + (sx - px + 1));
        f.stray_keys[pos] = ((unsigned long long)f.index(px, py) << 32) | ((unsigned long long)nb << 28) | (unsigned long long)(sample & 0x0FFFFFFFu);
        f.stray_vals[pos] = make_float4(c.r, c.g, c.b, w);
    } else {
        atomicAdd(&counters[C_STRAY_OVERFLOW], 1ull);                     // still accumulated, but in arrival order
        film_atomic_add(f, px, py, c, w);
    }
}

// Exact mode (box filter, r = 0.5): one thread per pixel adds that pixel's samples of the batch in sample order.
__global__ void __launch_bounds__(kThreads) k_film_accumulate_exact(PathMap map, FilmView f, PathBuffers b, int n_samples) {
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < map.n_pix; pix += gridDim.x * blockDim.x) {
        const int x = f.sb_x0 + (int)(pix % (uint32_t)f.sb_w), y = f.sb_y0 + (int)(pix / (uint32_t)f.sb_w);
        const bool inside = x >= f.px0 && y >= f.py0 && x < f.px1 && y < f.py1;
        float4 acc = inside ? f.acc[f.index(x, y)] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < n_samples; ++s) {
            const uint64_t slot = (uint64_t)s * map.n_pix + pix;
            const float4 Lf = b.L[slot];
            const rgb3 L = clamp_luminance(f, guard_radiance(mkc(Lf.x, Lf.y, Lf.z)));
            const SlotInfo si = slot_info(map, f, slot);
            PathSampler rng;
            rng.start(map.smp, si);
            float u0, u1;
            rng.film_offset(si, &u0, &u1);
            const float pfx = (float)x + u0;
            const float pfy = (float)y + u1;
            film_footprint(f, pfx, pfy, [&](int px, int py, float fw) {
                const rgb3 c = L * 1.0f * fw;                            // l * sample_weight * filter_weight
                if (px == x && py == y) { acc.x += c.r; acc.y += c.g; acc.z += c.b; acc.w += fw; }
                else film_stray(f, b.counters, x, y, si.sample, px, py, c, fw);
            });
        }
        if (inside) f.acc[f.index(x, y)] = acc;
    }
}
// General mode: one thread per path, atomics into the call's accumulators.
__global__ void __launch_bounds__(kThreads) k_film_accumulate_atomic(uint64_t n, PathMap map, FilmView f, PathBuffers b) {
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const float4 Lf = b.L[slot];
        const rgb3 L = clamp_luminance(f, guard_radiance(mkc(Lf.x, Lf.y, Lf.z)));
        const SlotInfo si = slot_info(map, f, slot);
        PathSampler rng;
        rng.start(map.smp, si);
        float u0, u1;
        rng.film_offset(si, &u0, &u1);
        const float pfx = (float)si.x + u0;
        const float pfy = (float)si.y + u1;
        film_footprint(f, pfx, pfy, [&](int px, int py, float fw) { film_atomic_add(f, px, py, L * 1.0f * fw, fw); });
    }
}
// Wide filters: one CTA per 32x8 tile of sample-bounds pixels.  A sample of pixel (x, y) only touches pixels within
// halo = ceil(radius + 0.5) of it, so the CTA sums all samples of its tile into a shared-memory copy of tile + halo
// (shared-memory atomics) and then adds that copy to the call's accumulators: ~(40 x 16) global atomics per tile and batch
// instead of 25 per sample (ncu: the per-sample version ran at 0.27 IPC behind 530 M L2 atomics).
constexpr int kTileW = 32, kTileH = 8, kMaxHalo = 12;   // (32 + 24) x (8 + 24) float4 = 28 KB of shared memory at most
__global__ void __launch_bounds__(kTileW * kTileH) k_film_accumulate_tiled(PathMap map, FilmView f, PathBuffers b, int n_samples, int hx, int hy) {
    extern __shared__ float4 tile[];
    const int tw = kTileW + 2 * hx, th = kTileH + 2 * hy;
    for (int i = threadIdx.x; i < tw * th; i += blockDim.x) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int tiles_x = (f.sb_w + kTileW - 1) / kTileW;
    const int tx0 = (int)(blockIdx.x % (unsigned)tiles_x) * kTileW, ty0 = (int)(blockIdx.x / (unsigned)tiles_x) * kTileH;
    const int lx = (int)threadIdx.x % kTileW, ly = (int)threadIdx.x / kTileW;
    const int sx = tx0 + lx, sy = ty0 + ly;                        // position inside the sample bounds
    const int ox = f.sb_x0 + tx0 - hx, oy = f.sb_y0 + ty0 - hy;    // image coordinates of tile[0]
    if (sx < f.sb_w && sy < f.sb_h) {
        const uint32_t pix = (uint32_t)sy * (uint32_t)f.sb_w + (uint32_t)sx;
        for (int s = 0; s < n_samples; ++s) {
            const uint64_t slot = (uint64_t)s * map.n_pix + pix;
            const float4 Lf = b.L[slot];
            const rgb3 L = clamp_luminance(f, guard_radiance(mkc(Lf.x, Lf.y, Lf.z)));
            const SlotInfo si = slot_info(map, f, slot);
            PathSampler rng;
            rng.start(map.smp, si);
            float u0, u1;
            rng.film_offset(si, &u0, &u1);
            film_footprint(f, (float)si.x + u0, (float)si.y + u1, [&](int px, int py, float fw) {
                const int cx = px - ox, cy = py - oy;
                if (cx >= 0 && cy >= 0 && cx < tw && cy < th) {
                    float* a = reinterpret_cast<float*>(tile + cy * tw + cx);
                    const rgb3 c = L * 1.0f * fw;
                    atomicAdd(a, c.r); atomicAdd(a + 1, c.g); atomicAdd(a + 2, c.b); atomicAdd(a + 3, fw);
                } else film_atomic_add(f, px, py, L * 1.0f * fw, fw);          // (cannot happen for halo >= ceil(radius + 0.5))
            });
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
        const float4 v = tile[i];
        const int px = ox + i % tw, py = oy + i / tw;
        if ((v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) && px >= f.px0 && py >= f.py0 && px < f.px1 && py < f.py1)
            film_atomic_add(f, px, py, mkc(v.x, v.y, v.z), v.w);
    }
}
// Sorted strays: the first thread of each run of equal target pixels adds the whole run in key order.
__global__ void __launch_bounds__(kThreads) k_apply_strays(FilmView f, const unsigned long long* keys, const uint32_t* index, const unsigned long long* counters) {
    unsigned long long n = counters[C_STRAYS];
    if (n > f.stray_capacity) n = f.stray_capacity;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long pixel = keys[j] >> 32;
        if (j > 0 && (keys[j - 1] >> 32) == pixel) continue;
        float4 acc = f.acc[pixel];
        for (unsigned long long k = j; k < n && (keys[k] >> 32) == pixel; ++k) {
            const float4 v = f.stray_vals[index[k]];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        f.acc[pixel] = acc;
    }
}
// Up to kSmallStrays strays (every frame of ordinary size: about 1e-5 of the samples stray): one CTA sorts {key, index} in shared
// memory (bitonic network) and applies the runs, instead of a device-wide radix sort over the whole buffer.
constexpr int kSmallStrays = 4096;
__global__ void __launch_bounds__(1024) k_strays_small(FilmView f, const unsigned long long* counters) {
    __shared__ unsigned long long key[kSmallStrays];
    __shared__ uint16_t idx[kSmallStrays];
    unsigned long long n64 = counters[C_STRAYS];
    if (n64 > f.stray_capacity) n64 = f.stray_capacity;
    const int n = (int)n64;
    if (n == 0) return;
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) { key[i] = i < n ? f.stray_keys[i] : ~0ull; idx[i] = (uint16_t)i; }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[l];
                    // (keys of distinct strays can be equal only for one sample reaching one pixel twice, which cannot happen)
                    if ((a > b) == up) { key[i] = b; key[l] = a; const uint16_t t = idx[i]; idx[i] = idx[l]; idx[l] = t; }
                }
            }
            __syncthreads();
        }
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const unsigned long long pixel = key[j] >> 32;
        if (j > 0 && (key[j - 1] >> 32) == pixel) continue;
        float4 acc = f.acc[pixel];
        for (int k = j; k < n && (key[k] >> 32) == pixel; ++k) {
            const float4 v = f.stray_vals[idx[k]];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        f.acc[pixel] = acc;
    }
}
__global__ void __launch_bounds__(kThreads) k_fill_index(uint32_t* idx, unsigned long long* keys, uint32_t cap, const unsigned long long* counters) {
    unsigned long long n = counters[C_STRAYS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        idx[i] = i;
        if (i >= n) keys[i] = ~0ull;
    }
}
// Film::merge_film_tile (film.rs:111-123): XYZ of the call's RGB sums is added to the film; the call accumulators reset.
__global__ void __launch_bounds__(kThreads) k_film_merge(FilmView f, unsigned long long* counters) {
    const size_t n = f.n_pixels();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = f.acc[i];
        float x, y, z;
        to_xyz(mkc(a.x, a.y, a.z), &x, &y, &z);
        float4 p = f.xyzw[i];
        p.x += x; p.y += y; p.z += z; p.w += a.w;
        f.xyzw[i] = p;
        f.acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && counters) counters[C_STRAYS] = 0;
}
// FilmTile::add_sample for explicit samples (pb2_film_add_samples): atomics, any filter.
__global__ void __launch_bounds__(kThreads) k_film_add_samples(FilmView f, const float2* pf, const float* L, const float* w, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const rgb3 c = clamp_luminance(f, mkc(L[3 * i], L[3 * i + 1], L[3 * i + 2]));
        const float sw = w[i];
        film_footprint(f, pf[i].x, pf[i].y, [&](int px, int py, float fw) { film_atomic_add(f, px, py, c * sw * fw, fw); });
    }
}
// Film::write_image (film.rs:153-178); the splat term (:167-172) only when the film has ever been splatted
__global__ void __launch_bounds__(kThreads) k_film_resolve(FilmView f, float scale, float splat_scale, float* rgb) {
    const size_t n = f.n_pixels();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 p = f.xyzw[i];
        rgb3 c = from_xyz(p.x, p.y, p.z);
        if (p.w != 0.0f) {
            const float inv = 1.0f / p.w;
            c = mkc(fmaxf(c.r * inv, 0.0f), fmaxf(c.g * inv, 0.0f), fmaxf(c.b * inv, 0.0f));
        }
        if (f.splat) {
            const float4 sp = f.splat[i];
            const rgb3 sc = from_xyz(sp.x, sp.y, sp.z);
            c = mkc(c.r + splat_scale * sc.r, c.g + splat_scale * sc.g, c.b + splat_scale * sc.b);
        }
        rgb[3 * i] = c.r * scale; rgb[3 * i + 1] = c.g * scale; rgb[3 * i + 2] = c.b * scale;
    }
}
// Film::add_splat (film.rs:137-151; D64 FIX: the port returns when the pixel IS inside the cropped bounds and tests the
// inclusive box — pbrt-v3: skip pixels outside, upper bound exclusive).  Float atomics: the order of splats on one pixel is
// not fixed, as in the reference (AtomicFloat from many threads).
__global__ void __launch_bounds__(kThreads) k_film_add_splats(FilmView f, const float2* pf, const float* v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float fx = floorf(pf[i].x), fy = floorf(pf[i].y);
        if (!(fx >= (float)f.px0 && fx < (float)f.px1 && fy >= (float)f.py0 && fy < (float)f.py1)) continue;
        const rgb3 c = clamp_luminance(f, mkc(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
        float x, y, z;
        to_xyz(c, &x, &y, &z);
        float* dst = reinterpret_cast<float*>(f.splat + f.index((int)fx, (int)fy));
        atomicAdd(dst, x); atomicAdd(dst + 1, y); atomicAdd(dst + 2, z);
    }
}
// Film::set_image (film.rs:125-135)
__global__ void __launch_bounds__(kThreads) k_film_set_image(FilmView f, const float* rgb) {
    const size_t n = f.n_pixels();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float x, y, z;
        to_xyz(mkc(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), &x, &y, &z);
        f.xyzw[i] = make_float4(x, y, z, 1.0f);
        if (f.splat) f.splat[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}
__global__ void k_copy_li(uint64_t n, PathMap map, FilmView f, PathBuffers b, float* L_out, float* pf_out) {
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const float4 Lf = b.L[slot];
        L_out[3 * slot] = Lf.x; L_out[3 * slot + 1] = Lf.y; L_out[3 * slot + 2] = Lf.z;
        const SlotInfo si = slot_info(map, f, slot);
        PathSampler rng;
        rng.start(map.smp, si);
        float u0, u1;
        rng.film_offset(si, &u0, &u1);
        pf_out[2 * slot] = (float)si.x + u0;
        pf_out[2 * slot + 1] = (float)si.y + u1;
    }
}

int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

unsigned grid_for(const Wavefront* wf, uint64_t n, int per_sm = 8) {
    const uint64_t want = (n + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)wf->sm_count * per_sm;
    return (unsigned)std::max<uint64_t>(1, std::min(want, cap));
}

// Stable three-way select of the active queue `in` by the paths' state bytes (see SelectJob); counts stay on the device.
void compact_queues(Wavefront* wf, const uint32_t* in, int n_in, bool by_class, uint32_t* o0, int c0, uint32_t* o1, int c1, uint32_t* o2, int c2,
                    cudaStream_t st) {
    const PathBuffers& b = wf->b;
    SelectJob j;
    j.in = in; j.n_in = b.counters + n_in; j.state = b.state; j.by_class = by_class ? 1 : 0;
    j.out[0] = o0; j.out[1] = o1; j.out[2] = o2;
    j.n_out[0] = b.counters + c0; j.n_out[1] = b.counters + c1; j.n_out[2] = b.counters + c2;
    j.status = b.select_status;
    j.max_tiles = (uint32_t)(wf->capacity / kSelTile + 1);
    const unsigned launch = wf->select_launches++;
    j.ticket = b.select_tickets + (launch & 1u);
    j.ticket_next = b.select_tickets + ((launch + 1u) & 1u);
    j.epoch = launch + 1u;                                 // (2^32 launches = years of rendering)
    j.counters = b.counters;
    k_select3<<<(unsigned)wf->sm_count * 4u, kSelThreads, 0, st>>>(j);
}

// k_shade<material, PixelSampler tables, mesh shading geometry> for the three material queues of one bounce.
template <bool TABLES, bool SG>
void launch_shade_t(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st) {
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + PB2_SHADE_THREADS - 1) / PB2_SHADE_THREADS,
                                                                             (uint64_t)wf->sm_count * 2 * PB2_SHADE_BLOCKS));
    // (a class no material of the scene has leaves its queue empty every bounce: no launch)
    if (sh.class_mask & 1u) k_shade<0, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
    if (sh.class_mask & 2u) k_shade<1, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
    if (sh.class_mask & 4u) k_shade<2, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
}
void launch_shade(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                  const PathParams& pp, int cur, uint64_t n, cudaStream_t st) {
    const bool tables = map.smp.kind == 2 || map.smp.kind == 3, sg = sh.indices != nullptr || sv.spheres != nullptr;   // SG = the general vertex
    if (tables && sg) launch_shade_t<true, true>(wf, sv, sh, b, map, film, pp, cur, n, st);
    else if (tables) launch_shade_t<true, false>(wf, sv, sh, b, map, film, pp, cur, n, st);
    else if (sg) launch_shade_t<false, true>(wf, sv, sh, b, map, film, pp, cur, n, st);
    else launch_shade_t<false, false>(wf, sv, sh, b, map, film, pp, cur, n, st);
}

// All bounces of one batch of `n` path slots.
void trace_batch(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film, const PathMap& map,
                 const PathParams& pp, uint64_t n, cudaStream_t st) {
    PathBuffers& b = wf->b;
    if (pp.integrator == 1) {                                            // VolPathIntegrator: one kernel carries whole paths
        k_volpath<<<grid_for(wf, n), 128, 0, st>>>(n, sv, sh, b, map, film, cam, pp);
        wf->totals[4] += 1;
        return;
    }
    const TraceTuning tune = trace_tuning();
    const unsigned trace_grid = (unsigned)wf->sm_count * (unsigned)PB2_MIN_BLOCKS, trace_grid_sph = (unsigned)wf->sm_count * 4u;
    k_raygen<<<grid_for(wf, n), kThreads, 0, st>>>(n, map, film, cam, b);
    int launches = 1;
    for (int depth = 0; depth <= pp.max_depth; ++depth) {
        const int cur = depth & 1;
        if (sv.spheres) k_extend_spheres<<<trace_grid_sph, 128, 0, st>>>(sv, sh, b, cur, tune);
        else k_extend<<<trace_grid, 128, 0, st>>>(sv, sh, b, cur, tune);
        // hits -> one queue per shading class (material-sorted shading)
        compact_queues(wf, b.q_active[cur], C_ACTIVE_A + cur, true, b.q_mat[0], C_MAT0, b.q_mat[1], C_MAT1, b.q_mat[2], C_MAT2, st);
        launch_shade(wf, sv, sh, b, map, film, pp, cur, n, st);
        launches += 2 + __builtin_popcount(sh.class_mask & 7u);
        if (depth == pp.max_depth) break;                                // path.rs:90-92: nothing continues, no NEE record was written
        compact_queues(wf, b.q_active[cur], C_ACTIVE_A + cur, false, b.q_active[cur ^ 1], C_ACTIVE_A + (cur ^ 1), b.q_shadow, C_SHADOW,
                       b.q_mis, C_MIS, st);
        launches += 1;
        if (sh.n_lights > 0) {
            if (sv.spheres) k_shadow_spheres<<<trace_grid_sph, 128, 0, st>>>(sv, b, tune);
            else k_shadow<<<trace_grid, 128, 0, st>>>(sv, b, tune);           // (the MIS rays ride in the next bounce's k_extend)
            launches += 1;
        }
    }
    wf->totals[4] += (uint64_t)launches;
}

}  // namespace

int wavefront_create(uint64_t capacity, Wavefront** out) {
    Wavefront* wf = new Wavefront();
    wf->capacity = capacity;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&wf->sm_count, cudaDevAttrMultiProcessorCount, dev);
    // one arena: 13 float4/uint4 arrays, rng, occluded, mis_prim, 7 queues, counters
    const size_t f4 = capacity * 16;
    const size_t status_bytes = 3 * (capacity / kSelTile + 1) * 8;
    size_t bytes = 13 * f4 + capacity * 8 + capacity * 4 + 2 * capacity + 7 * capacity * 4 + C_COUNT * 8 + status_bytes + 256 + 8192;
    cudaError_t e = cudaMalloc(&wf->arena, bytes);
    if (e != cudaSuccess) { delete wf; *out = nullptr; return (int)e; }
    char* p = (char*)wf->arena;
    auto take = [&](size_t n) { char* r = p; p += (n + 255) & ~(size_t)255; return r; };
    PathBuffers& b = wf->b;
    b.ray_o = (float4*)take(f4); b.ray_d = (float4*)take(f4); b.beta = (float4*)take(f4); b.L = (float4*)take(f4);
    b.hit = (uint4*)take(f4);
    b.sh_o = (float4*)take(f4); b.sh_d = (float4*)take(f4); b.t1 = (float4*)take(f4);
    b.mis_o = (float4*)take(f4); b.mis_d = (float4*)take(f4); b.t2 = (float4*)take(f4); b.beta_nee = (float4*)take(f4);
    b.rng = (unsigned long long*)take(capacity * 8);
    b.mis_prim = (uint32_t*)take(capacity * 4);
    b.occluded = (uint8_t*)take(capacity);
    b.state = (uint8_t*)take(capacity);
    b.select_status = (unsigned long long*)take(status_bytes);
    b.select_tickets = (unsigned*)take(256);
    cudaMemset(b.select_status, 0, status_bytes);          // epoch 0 = never written
    cudaMemset(b.select_tickets, 0, 256);
    for (int i = 0; i < 2; ++i) b.q_active[i] = (uint32_t*)take(capacity * 4);
    for (int i = 0; i < 3; ++i) b.q_mat[i] = (uint32_t*)take(capacity * 4);
    b.q_shadow = (uint32_t*)take(capacity * 4);
    b.q_mis = (uint32_t*)take(capacity * 4);
    b.counters = (unsigned long long*)take(C_COUNT * 8);
    cudaMemset(b.counters, 0, C_COUNT * 8);
    *out = wf;
    return 0;
}

void wavefront_destroy(Wavefront* wf) {
    if (!wf) return;
    if (wf->arena) cudaFree(wf->arena);
    delete wf;
}

int wavefront_render(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film_in,
                     const PathParams& pp, const SamplerView& smp, int spp, int sample_begin, int sample_end, cudaStream_t st) {
    const uint64_t n_pix = (uint64_t)film_in.sb_w * (uint64_t)film_in.sb_h;
    if (n_pix == 0 || sample_end <= sample_begin) return 0;
    // Strays (exact mode) are sorted once per call.  About 1e-5 of a call's samples stray (p_film = x + u rounds up to x + 1), so
    // the sort covers a slice of the stray buffer sized for 1 / 256 of them — at least the 4096 entries one CTA sorts in shared
    // memory — not the whole 2^20-entry buffer; anything beyond the slice is accumulated in arrival order and counted
    // (C_STRAY_OVERFLOW), as with the full buffer before.
    FilmView film = film_in;
    {
        const uint64_t samples = n_pix * (uint64_t)(sample_end - sample_begin);
        uint64_t want = kSmallStrays;
        while (want < samples / 256) want <<= 1;
        film.stray_capacity = (uint32_t)std::min<uint64_t>(film_in.stray_capacity, want);
    }
    const int per_batch = (int)std::max<uint64_t>(1, wf->capacity / n_pix);
    for (int s0 = sample_begin; s0 < sample_end; s0 += per_batch) {
        const int ns = std::min(per_batch, sample_end - s0);
        PathMap map{(uint32_t)n_pix, spp, s0, nullptr, nullptr, smp};
        const uint64_t n = n_pix * (uint64_t)ns;
        trace_batch(wf, sv, sh, cam, film, map, pp, n, st);
        if (film.exact) k_film_accumulate_exact<<<grid_for(wf, n_pix), kThreads, 0, st>>>(map, film, wf->b, ns);
        else {
            const int hx = (int)std::ceil(film.radius_x + 0.5f), hy = (int)std::ceil(film.radius_y + 0.5f);
            if (hx <= kMaxHalo && hy <= kMaxHalo) {
                const unsigned tiles = (unsigned)(((film.sb_w + kTileW - 1) / kTileW) * ((film.sb_h + kTileH - 1) / kTileH));
                const size_t smem = (size_t)(kTileW + 2 * hx) * (kTileH + 2 * hy) * sizeof(float4);
                k_film_accumulate_tiled<<<tiles, kTileW * kTileH, smem, st>>>(map, film, wf->b, ns, hx, hy);
            } else k_film_accumulate_atomic<<<grid_for(wf, n), kThreads, 0, st>>>(n, map, film, wf->b);
        }
        wf->totals[4] += 1;
    }
    film_finish(film, wf->b.counters, st);
    return 0;
}

int wavefront_li(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film,
                 const PathParams& pp, const SamplerView& smp, int spp, const int32_t* d_xy, const uint32_t* d_s, uint64_t n, float* d_L,
                 float* d_pfilm, cudaStream_t st) {
    if (n == 0) return 0;
    PathMap map{(uint32_t)std::max<uint64_t>(1, n), spp, 0, d_xy, d_s, smp};
    trace_batch(wf, sv, sh, cam, film, map, pp, n, st);
    k_copy_li<<<grid_for(wf, n), kThreads, 0, st>>>(n, map, film, wf->b, d_L, d_pfilm);
    return 0;
}

void pixel_tables_generate(int kind, uint32_t n_pix, int spp, int n_dims, int x_samples, int y_samples, int jitter, uint64_t seq0,
                           float* d_t1, float2* d_t2, cudaStream_t st) {
    const TableGen g{kind, spp, n_dims, x_samples, y_samples, jitter, n_pix, seq0, d_t1, d_t2};
    const unsigned grid = (unsigned)std::max<uint64_t>(1, ((uint64_t)n_pix + kThreads - 1) / kThreads);
    k_pixel_tables<<<grid, kThreads, 0, st>>>(g);
}

// Ordered application of the strays (exact mode), then merge of the call's sums into the film.
void film_finish(const FilmView& film, unsigned long long* counters, cudaStream_t st) {
    const unsigned grid = (unsigned)device_sm_count() * 4;
    if (film.exact && counters && film.stray_capacity && film.stray_capacity <= (uint32_t)kSmallStrays) {
        k_strays_small<<<1, 1024, 0, st>>>(film, counters);
    } else if (film.exact && counters && film.stray_capacity) {
        // scratch: sorted keys + index pairs live behind the primary arrays (allocated 2x by the film)
        unsigned long long* keys_in = film.stray_keys;
        unsigned long long* keys_out = film.stray_keys + film.stray_capacity;
        uint32_t* idx_in = reinterpret_cast<uint32_t*>(film.stray_keys + 2ull * film.stray_capacity);
        uint32_t* idx_out = idx_in + film.stray_capacity;
        void* temp = idx_out + film.stray_capacity;
        k_fill_index<<<grid, kThreads, 0, st>>>(idx_in, keys_in, film.stray_capacity, counters);
        size_t temp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)film.stray_capacity, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)film.stray_capacity, 0, 64, st);
        k_apply_strays<<<grid, kThreads, 0, st>>>(film, keys_out, idx_out, counters);
    }
    k_film_merge<<<grid, kThreads, 0, st>>>(film, counters);
}

size_t film_sort_scratch_bytes(uint32_t capacity) {
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)capacity, 0, 64, (cudaStream_t)0);
    return temp_bytes + 256;
}

void film_add_samples(const FilmView& film, const float* d_pfilm, const float* d_L, const float* d_w, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    k_film_add_samples<<<(unsigned)device_sm_count() * 4, kThreads, 0, st>>>(film, (const float2*)d_pfilm, d_L, d_w, n);
    k_film_merge<<<(unsigned)device_sm_count() * 4, kThreads, 0, st>>>(film, nullptr);
}

void film_resolve(const FilmView& film, float scale, float splat_scale, float* d_rgb, cudaStream_t st) {
    k_film_resolve<<<(unsigned)device_sm_count() * 4, kThreads, 0, st>>>(film, scale, splat_scale, d_rgb);
}
void film_add_splats(const FilmView& film, const float* d_pfilm, const float* d_v, uint64_t n, cudaStream_t st) {
    if (n) k_film_add_splats<<<(unsigned)device_sm_count() * 4, kThreads, 0, st>>>(film, (const float2*)d_pfilm, d_v, n);
}
void film_set_image(const FilmView& film, const float* d_rgb, cudaStream_t st) {
    k_film_set_image<<<(unsigned)device_sm_count() * 4, kThreads, 0, st>>>(film, d_rgb);
}

}  // namespace pb2
