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

#include "wavefront_dev.cuh"

namespace pb2 {

namespace {

// ---- order-preserving queues -------------------------------------------------------------------------------------------------
// The stages talk through queues of path slots.  The visibility / shade kernels do not append to them: each leaves one byte per
// path (PathBuffers::state) and compact_queues() rebuilds the queues from the bounce's active queue with a stable three-way
// select, so every queue stays sorted by slot and neighbouring lanes of the next kernel read neighbouring path state.
// (Appending in completion order of the persistent traversal — one warp-aggregated atomic per group of lanes — scattered a
// warp's slots over a ~150 K-slot window; with the queues sorted k_shade runs 34-39 % faster, k_shadow 11-22 %, k_extend 3-8 %:
// profiles/r01_tuning.md, session 4.)

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
__global__ void __launch_bounds__(kSelThreads, 4) k_select3(SelectJob j) {
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
        // ---- warp-striped tile: warp w owns entries [w * 32 * kSelItems, (w + 1) * 32 * kSelItems) of the tile, lane l its rows
        // k * 32 + l, so every load, every state-byte gather (the queue is sorted by slot) and every store of a warp instruction
        // falls on neighbouring addresses.  (First version: kSelItems consecutive entries per thread — a warp instruction then
        // spanned 2 KB of the queue and 32 separate runs of the outputs; ncu saw 4x the L2 traffic the select needs.) ----
        const unsigned long long w0 = (unsigned long long)tile * kSelTile + (unsigned long long)warp * (32u * kSelItems) + lane;
        uint32_t slot[kSelItems];
#pragma unroll
        for (int k = 0; k < kSelItems; ++k) slot[k] = w0 + 32u * k < n ? j.in[w0 + 32u * k] : 0u;
        unsigned bits[3] = {0u, 0u, 0u};           // bit k of bits[q]: row k of this lane goes to output q
#pragma unroll
        for (int k = 0; k < kSelItems; ++k) {
            const unsigned st = w0 + 32u * k < n ? j.state[slot[k]] : (j.by_class ? kStateDead : 0u);
#pragma unroll
            for (int q = 0; q < 3; ++q) bits[q] |= (select_pred(j.by_class, st, q) ? 1u : 0u) << k;
        }
        // warp totals per output
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const unsigned t = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(bits[q]));
            if (lane == 0u) s_warp[q][warp] = t;
        }
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
        // ---- scatter: rows in order, lanes in order inside a row (stable) ----
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            unsigned pos = s_base[q];
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; ++w) pos += (unsigned)w < warp ? s_warp[q][w] : 0u;
            uint32_t* out = j.out[q];
#pragma unroll
            for (int k = 0; k < kSelItems; ++k) {
                const bool mine = (bits[q] >> k) & 1u;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, mine);
                if (mine) out[pos + __popc(m & lt)] = slot[k];
                pos += __popc(m);
            }
        }
        if (tile == n_tiles - 1u && threadIdx.x == 0) {
            unsigned long long tot[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                unsigned all = 0u;
#pragma unroll
                for (int w = 0; w < kSelThreads / 32; ++w) all += s_warp[q][w];
                tot[q] = (unsigned long long)s_base[q] + all;
            }
            select_finish(j, tot[0], tot[1], tot[2]);
        }
        __syncthreads();                                 // s_tile, s_warp, s_base are reused by the next tile
    }
}

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
// (stray_counters: the frame's stray list is shared by the two wavefronts that alternate its batches)
__global__ void __launch_bounds__(kThreads) k_film_accumulate_exact(PathMap map, FilmView f, PathBuffers b, int n_samples, unsigned long long* stray_counters) {
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < map.n_pix; pix += gridDim.x * blockDim.x) {
        const uint32_t row = fast_div(pix, f.sb_w_magic);
        const int x = f.sb_x0 + (int)(pix - row * (uint32_t)f.sb_w), y = f.sb_y0 + (int)row;
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
                else film_stray(f, stray_counters, x, y, si.sample, px, py, c, fw);
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
                    cudaStream_t st, const uint8_t* state = nullptr) {
    const PathBuffers& b = wf->b;
    SelectJob j;
    j.in = in; j.n_in = b.counters + n_in; j.state = state ? state : b.state; j.by_class = by_class ? 1 : 0;
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
    if (pp.integrator == 1) {                                            // VolPathIntegrator (wavefront_volpath.cu)
        static const bool mega = [] { const char* e = getenv("PB2_VOLPATH_MEGAKERNEL"); return e && atoi(e) != 0; }();
        if (mega) {                                                      // one thread carries a whole path (the first device version; kept for comparison)
            launch_volpath(wf, grid_for(wf, n), n, sv, sh, b, map, film, cam, pp, st);
            wf->totals[4] += 1;
        } else trace_batch_vol(wf, sv, sh, cam, film, map, pp, n, st);
        return;
    }
    const TraceTuning tune = trace_tuning();
    const unsigned trace_grid = (unsigned)wf->sm_count * (unsigned)PB2_MIN_BLOCKS, trace_grid_sph = (unsigned)wf->sm_count * 4u;
    // k_shadow (any hit: no candidate bookkeeping) needs fewer registers than k_extend and fits one CTA per SM more
    static const int shadow_per_sm = [] {
        int v = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, (const void*)k_shadow, 128, 0);
        return std::max(v, PB2_MIN_BLOCKS);
    }();
    const unsigned trace_grid_shadow = (unsigned)wf->sm_count * (unsigned)shadow_per_sm;
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
            else k_shadow<<<trace_grid_shadow, 128, 0, st>>>(sv, b, tune);    // (the MIS rays ride in the next bounce's k_extend)
            launches += 1;
        }
    }
    wf->totals[4] += (uint64_t)launches;
}

}  // namespace

// the queue select and ray generation for the other translation units (wavefront_volpath.cu)
void select_queues(Wavefront* wf, const uint32_t* in, int n_in, bool by_class, uint32_t* o0, int c0, uint32_t* o1, int c1, uint32_t* o2, int c2,
                   cudaStream_t st, const uint8_t* state) {
    compact_queues(wf, in, n_in, by_class, o0, c0, o1, c1, o2, c2, st, state);
}
void launch_raygen(Wavefront* wf, uint64_t n, const PathMap& map, const FilmView& film, const CameraView& cam, cudaStream_t st) {
    k_raygen<<<grid_for(wf, n), kThreads, 0, st>>>(n, map, film, cam, wf->b);
}

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
    cudaMallocHost((void**)&wf->h_counters, C_COUNT * 8);
    *out = wf;
    return 0;
}

void wavefront_destroy(Wavefront* wf) {
    if (!wf) return;
    if (wf->peer) wavefront_destroy(wf->peer);
    if (wf->aux_stream) cudaStreamDestroy(wf->aux_stream);
    for (cudaEvent_t e : {wf->ev_fork, wf->ev_join, wf->ev_acc[0], wf->ev_acc[1]})
        if (e) cudaEventDestroy(e);
    if (wf->arena) cudaFree(wf->arena);
    if (wf->vol_arena) cudaFree(wf->vol_arena);
    if (wf->h_counters) cudaFreeHost(wf->h_counters);
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
    const int n_batches = (sample_end - sample_begin + per_batch - 1) / per_batch;
    // Two batches in flight (see Wavefront::peer): measured -7.3 % on C4 and -5.8 % on C2 with two independent frames on two
    // streams (profiles/r02_tuning.md); PB2_TWO_STREAMS=0 renders the batches one after the other on the caller's stream.
    static const bool two_streams = [] { const char* e = getenv("PB2_TWO_STREAMS"); return !e || atoi(e) != 0; }();
    bool overlap = two_streams && n_batches >= 2;
    if (overlap && !wf->peer) {
        if (wavefront_create(wf->capacity, &wf->peer) != 0) { wf->peer = nullptr; cudaGetLastError(); }
        else if (cudaStreamCreateWithFlags(&wf->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
                 cudaEventCreateWithFlags(&wf->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                 cudaEventCreateWithFlags(&wf->ev_join, cudaEventDisableTiming) != cudaSuccess ||
                 cudaEventCreateWithFlags(&wf->ev_acc[0], cudaEventDisableTiming) != cudaSuccess ||
                 cudaEventCreateWithFlags(&wf->ev_acc[1], cudaEventDisableTiming) != cudaSuccess) {
            wavefront_destroy(wf->peer);
            wf->peer = nullptr;
            cudaGetLastError();
        }
    }
    overlap = overlap && wf->peer;
    if (overlap) {
        cudaEventRecord(wf->ev_fork, st);
        cudaStreamWaitEvent(wf->aux_stream, wf->ev_fork, 0);
    }
    int k = 0;
    for (int s0 = sample_begin; s0 < sample_end; s0 += per_batch, ++k) {
        const int ns = std::min(per_batch, sample_end - s0);
        PathMap map{(uint32_t)n_pix, spp, s0, nullptr, nullptr, smp, fast_div_magic((uint32_t)n_pix)};
        const uint64_t n = n_pix * (uint64_t)ns;
        Wavefront* w = overlap && (k & 1) ? wf->peer : wf;
        cudaStream_t ws = overlap && (k & 1) ? wf->aux_stream : st;
        trace_batch(w, sv, sh, cam, film, map, pp, n, ws);
        if (overlap && k > 0) cudaStreamWaitEvent(ws, wf->ev_acc[(k - 1) & 1], 0);      // the film takes the batches in order
        if (film.exact) k_film_accumulate_exact<<<grid_for(wf, n_pix), kThreads, 0, ws>>>(map, film, w->b, ns, wf->b.counters);
        else {
            const int hx = (int)std::ceil(film.radius_x + 0.5f), hy = (int)std::ceil(film.radius_y + 0.5f);
            if (hx <= kMaxHalo && hy <= kMaxHalo) {
                const unsigned tiles = (unsigned)(((film.sb_w + kTileW - 1) / kTileW) * ((film.sb_h + kTileH - 1) / kTileH));
                const size_t smem = (size_t)(kTileW + 2 * hx) * (kTileH + 2 * hy) * sizeof(float4);
                k_film_accumulate_tiled<<<tiles, kTileW * kTileH, smem, ws>>>(map, film, w->b, ns, hx, hy);
            } else k_film_accumulate_atomic<<<grid_for(wf, n), kThreads, 0, ws>>>(n, map, film, w->b);
        }
        if (overlap) cudaEventRecord(wf->ev_acc[k & 1], ws);
        w->totals[4] += 1;
    }
    if (overlap) {
        cudaEventRecord(wf->ev_join, wf->aux_stream);
        cudaStreamWaitEvent(st, wf->ev_join, 0);
    }
    film_finish(film, wf->b.counters, st);
    return 0;
}

int wavefront_li(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film,
                 const PathParams& pp, const SamplerView& smp, int spp, const int32_t* d_xy, const uint32_t* d_s, uint64_t n, float* d_L,
                 float* d_pfilm, cudaStream_t st) {
    if (n == 0) return 0;
    PathMap map{(uint32_t)std::max<uint64_t>(1, n), spp, 0, d_xy, d_s, smp, fast_div_magic((uint32_t)std::max<uint64_t>(1, n))};
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
