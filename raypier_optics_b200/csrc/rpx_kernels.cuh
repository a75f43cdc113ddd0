// rpx_kernels.cuh -- the wavefront tracer kernels of librpx (sm_100a).
//
// One generation of rays = two kernels:
//   k_intersect : ray-parallel nearest-hit search over the face table (staged in shared
//                 memory).  Reads origin+direction (48 B/ray, coalesced SoA), writes the
//                 parent write-back the reference performs (length f64 + end_face_idx u32).
//   k_shade     : orientation + material evaluation + ORDERED child emission.  Children
//                 must appear in parent order, reflected before transmitted
//                 (ctracer.pyx:2084-2117 appends in loop order), so the kernel does a
//                 block-level exclusive scan of the child counts and a decoupled
//                 look-back across tiles (single pass, no second read of the parents,
//                 no staging copy of the children).
// Replaces trace_segment_c (ctracer.pyx:2062-2118) and trace_gausslet_c +
// trace_parabasal_rays (ctracer.pyx:2214-2281, 2350-2385).
#pragma once
#include "rpx_materials.cuh"

namespace rpx {

// ------------------------------------------------------------------ SoA layout
// f[field * cap + i] (doubles), u[field * cap + i] (u32), p[(j*10 + c) * cap + i] (para)
enum {
    F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_NX, F_NY, F_NZ, F_EX, F_EY, F_EZ,
    F_NR, F_NI, F_E1R, F_E1I, F_E2R, F_E2I, F_LEN, F_PHASE, F_APATH, NF = 21
};
enum { U_WL = 0, U_PARENT, U_ENDFACE, U_IDENT, U_TYPE, NU = 5 };
enum { P_OX = 0, P_DX = 3, P_NX = 6, P_LEN = 9, NPF = 10, NP = 60 };

struct Soa {
    double* f;
    uint32_t* u;
    double* p;  // nullptr for plain rays
    // mesh / UV patch scenes only: facet record (+ 1; 0 = none) of the hit stored in end_face_idx, written by the
    // launch that found it (k_intersect, trace-ahead of k_shade) so that k_shade need not walk the BVH again for
    // intersect_t.piece_idx / .uv.  NULL when the collection's hits were not found on the device (imported rays).
    uint32_t* piece;
    unsigned long long n, cap;
};

// Rays per tile = threads per CTA.  128 is what ships; -DRPX_TILE=64 -DRPX_MIN_BLOCKS=8 (same warps per SM,
// half the barrier domain, twice the look-backs -- cheap since the grouped look-back) is prepared for
// the next round and has not run on a GPU.  The lean staging map is a byte: RPX_TILE <= 256.
#ifndef RPX_TILE
#define RPX_TILE 128
#endif
#define RPX_WORDS_RAY 47        // 188 / 4
#define RPX_WORDS_GAUSSLET 167  // 668 / 4

// ------------------------------------------------------------------ AoS <-> SoA
// The host-visible records are packed (4-byte aligned) 188 / 668-byte structs.  A tile
// of records is moved through shared memory as 32-bit words: global side fully
// coalesced, shared side stride 47 / 167 words (odd -> conflict free).
template <int WORDS, int TILE>
__global__ void __launch_bounds__(TILE) k_aos_to_soa(const uint32_t* __restrict__ aos, Soa out) {
    extern __shared__ uint32_t sm[];
    const unsigned long long base = (unsigned long long)blockIdx.x * TILE;
    const unsigned long long n = out.n;
    const int cnt = (int)min((unsigned long long)TILE, n - base);
    const uint32_t* src = aos + base * WORDS;
    for (int w = threadIdx.x; w < cnt * WORDS; w += TILE) sm[w] = src[w];
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= cnt) return;
    const unsigned long long i = base + t;
    const uint32_t* rec = sm + t * WORDS;
#pragma unroll
    for (int k = 0; k < NF; k++) {
        uint2 v = make_uint2(rec[2 * k], rec[2 * k + 1]);
        out.f[k * out.cap + i] = __hiloint2double((int)v.y, (int)v.x);
    }
#pragma unroll
    for (int k = 0; k < NU; k++) out.u[k * out.cap + i] = rec[2 * NF + k];
    if (WORDS == RPX_WORDS_GAUSSLET) {
#pragma unroll 4
        for (int k = 0; k < NP; k++) {
            uint2 v = make_uint2(rec[RPX_WORDS_RAY + 2 * k], rec[RPX_WORDS_RAY + 2 * k + 1]);
            out.p[k * out.cap + i] = __hiloint2double((int)v.y, (int)v.x);
        }
    }
}

template <int WORDS, int TILE>
__global__ void __launch_bounds__(TILE) k_soa_to_aos(Soa in, uint32_t* __restrict__ aos, uint32_t parent_offset) {
    extern __shared__ uint32_t sm[];
    const unsigned long long base = (unsigned long long)blockIdx.x * TILE;
    const unsigned long long n = in.n;
    const int cnt = (int)min((unsigned long long)TILE, n - base);
    const int t = threadIdx.x;
    if (t < cnt) {
        const unsigned long long i = base + t;
        uint32_t* rec = sm + t * WORDS;
#pragma unroll
        for (int k = 0; k < NF; k++) {
            double v = in.f[k * in.cap + i];
            rec[2 * k] = (uint32_t)__double2loint(v);
            rec[2 * k + 1] = (uint32_t)__double2hiint(v);
        }
#pragma unroll
        for (int k = 0; k < NU; k++) rec[2 * NF + k] = in.u[k * in.cap + i] + (k == U_PARENT ? parent_offset : 0u);
        if (WORDS == RPX_WORDS_GAUSSLET) {
#pragma unroll 4
            for (int k = 0; k < NP; k++) {
                double v = in.p[k * in.cap + i];
                rec[RPX_WORDS_RAY + 2 * k] = (uint32_t)__double2loint(v);
                rec[RPX_WORDS_RAY + 2 * k + 1] = (uint32_t)__double2hiint(v);
            }
        }
    }
    __syncthreads();
    uint32_t* dst = aos + base * WORDS;
    for (int w = threadIdx.x; w < cnt * WORDS; w += TILE) dst[w] = sm[w];
}

// reset_length_c for generation 0 of a gausslet trace (ctracer.pyx:1239-1245)
static __global__ void k_reset_length(Soa rays, double max_length) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rays.n) return;
    rays.f[F_LEN * rays.cap + i] = max_length;
    if (rays.p) {
#pragma unroll
        for (int j = 0; j < RPX_NPARA; j++) rays.p[(j * NPF + P_LEN) * rays.cap + i] = max_length;
    }
}

// ------------------------------------------------------------------ scene staging
// Copy the face table, the face-set transforms and the material table into shared memory and
// re-point the DevScene at them.  SS (scene-in-shared) is a template parameter and the copy is
// unconditional, so the compiler can PROVE that S.faces / S.sets / S.mats point into shared memory
// and emits LDS instead of generic LD (143 static LD -> 0 in k_shade; k_intersect 53 -> 47 us per
// 1e6 rays).  The host guarantees the fit; a scene whose tables exceed the budget runs the SS=false
// instantiations, which read the tables from global memory (compiled for the widest face class /
// material mask only).
template <bool SS>
RPX_DEV void stage_scene(DevScene& S, unsigned char* smem) {
    if (!SS) return;
    const int face_bytes = S.n_faces * (int)sizeof(rpx_face);
    const int set_bytes = S.n_sets * (int)sizeof(rpx_face_set);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
    const uint32_t* srcf = reinterpret_cast<const uint32_t*>(S.faces);
    const uint32_t* srcs = reinterpret_cast<const uint32_t*>(S.sets);
    const uint32_t* srcm = reinterpret_cast<const uint32_t*>(S.mats);
    const int fw = face_bytes / 4, sw = set_bytes / 4, mw = S.n_mats * (int)sizeof(rpx_material) / 4;
    for (int w = threadIdx.x; w < fw; w += blockDim.x) dst[w] = srcf[w];
    for (int w = threadIdx.x; w < sw; w += blockDim.x) dst[fw + w] = srcs[w];
    for (int w = threadIdx.x; w < mw; w += blockDim.x) dst[fw + sw + w] = srcm[w];
    __syncthreads();
    S.faces = reinterpret_cast<const rpx_face*>(smem);
    S.sets = reinterpret_cast<const rpx_face_set*>(smem + face_bytes);
    S.mats = reinterpret_cast<const rpx_material*>(smem + face_bytes + set_bytes);
}

// ------------------------------------------------------------------ nearest hit
// FaceList.intersect_c over every face set (ctracer.pyx:1882-1904, 2093-2104): the
// sequential strict-< update keeps the LOWEST face index on equal distances (quirk Q2);
// the running (distance, face) pair lives in two registers.
// only_face >= 0 restricts the search to that one face: FaceList.intersect_one_face_c
// (ctracer.pyx:1861-1879), the sequential-mode step of trace_one_face_segment_c (:2121-2170).
// Forced inline into its call sites: as a real call it cost 230 B of extra stack / spill traffic per
// thread for the ABI (measured: k_shade 0.132 -> 0.125 ms, gausslets 1.107 -> 0.983 ms).
template <int FC>
__device__ __forceinline__ void nearest_hit(
    const DevScene& S, vec3 o, vec3 d, double max_length, int only_face,
                                         double* out_len, uint32_t* out_face, uint32_t* out_rec = nullptr) {
    vec3 point = o + d * max_length;
    double best = max_length;  // ray.length = max_length (ctracer.pyx:2086)
    uint32_t best_face = RPX_NO_FACE;
    uint32_t best_rec = 0;  // mesh class: facet record + 1 of the best hit
    if (only_face >= 0) {
        const rpx_face* f = &S.faces[only_face];
        const rpx_face_set* fs = &S.sets[f->face_set];
        vec3 p1 = transform_pt(fs->inv_trans.m, o);
        vec3 p2 = transform_pt(fs->inv_trans.m, point);
        HitAux ha;
        ha.rec = -1;
        double dist = face_intersect<FC>(S, f, p1, p2, 1, FC == RPX_FC_MESH ? &ha : nullptr);
        if (f->tolerance < dist && dist < best) {
            best = dist;
            best_face = (uint32_t)only_face;
            if (FC == RPX_FC_MESH) best_rec = (uint32_t)(ha.rec + 1);
        }
        *out_len = best;
        *out_face = best_face;
        if (FC == RPX_FC_MESH && out_rec) *out_rec = best_rec;
        return;
    }
    for (int s = 0; s < S.n_sets; s++) {
        const rpx_face_set* fs = &S.sets[s];
        vec3 p1 = transform_pt(fs->inv_trans.m, o);
        vec3 p2 = transform_pt(fs->inv_trans.m, point);
        for (int fi = fs->face_begin; fi < fs->face_end; fi++) {
            const rpx_face* f = &S.faces[fi];
            HitAux ha;
            ha.rec = -1;
            double dist = face_intersect<FC>(S, f, p1, p2, 1, FC == RPX_FC_MESH ? &ha : nullptr);
            if (f->tolerance < dist && dist < best) {
                best = dist;
                best_face = (uint32_t)fi;
                if (FC == RPX_FC_MESH) best_rec = (uint32_t)(ha.rec + 1);
            }
        }
    }
    *out_len = best;
    *out_face = best_face;
    if (FC == RPX_FC_MESH && out_rec) *out_rec = best_rec;
}

// ------------------------------------------------------------------ k_intersect
// Generation 0 only: later generations get their nearest hit from the k_shade launch that
// creates them (the thread that just built a child still has it in registers).
// Simple face classes fit 80 registers (6 CTAs / SM: 0.047 -> 0.045 ms per 1e6 rays, prisms 0.084 -> 0.078);
// the Newton / quadric code of the full class keeps 128.
// Mesh class: the BVH walk is bound by the latency of its dependent node loads (ncu: issue slots 35 % used, fp64
// pipe 2 %), so resident warps count more than spill-free Newton code: RPX_MESH_IBLOCKS CTAs / SM.
#ifndef RPX_MESH_IBLOCKS
#define RPX_MESH_IBLOCKS 6  /* measured: 4 -> 8.7e8, 6 -> 9.5e8, 8 -> 8.6e8 seg/s on the 71k-facet scene */
#endif
template <int FC, bool SS>
__global__ void __launch_bounds__(RPX_TILE, (FC == RPX_FC_SIMPLE ? 6 : FC == RPX_FC_MESH ? RPX_MESH_IBLOCKS : 4) * 128 / RPX_TILE)
k_intersect(DevScene S, Soa rays, double max_length, int only_face) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_scene<SS>(S, smem);
    const unsigned long long i = (unsigned long long)blockIdx.x * RPX_TILE + threadIdx.x;
    if (i >= rays.n) return;
    const unsigned long long cap = rays.cap;
    vec3 o = v3(rays.f[F_OX * cap + i], rays.f[F_OY * cap + i], rays.f[F_OZ * cap + i]);
    vec3 d = v3(rays.f[F_DX * cap + i], rays.f[F_DY * cap + i], rays.f[F_DZ * cap + i]);
    double best;
    uint32_t best_face;
    uint32_t best_rec = 0;
    nearest_hit<FC>(S, o, d, max_length, only_face, &best, &best_face, &best_rec);
    rays.f[F_LEN * cap + i] = best;
    rays.u[U_ENDFACE * cap + i] = best_face;
    if (FC == RPX_FC_MESH && rays.piece) rays.piece[i] = best_rec;
}

// ------------------------------------------------------------------ ordered emission
// tile_state word: bits 63..62 = flag (1 aggregate, 2 inclusive prefix), low 62 = value
#define RPX_FLAG_AGG (1ull << 62)
#define RPX_FLAG_PREFIX (2ull << 62)
#define RPX_VAL_MASK ((1ull << 62) - 1)

RPX_DEV unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
RPX_DEV void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive scan of per-thread counts over the block; returns this thread's offset and
// the block total.  256 threads = 8 warps.
RPX_DEV uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RPX_TILE / 32; w++) {
        uint32_t s = s_warp[w];
        if (w < warp) woff += s;
        tot += s;
    }
    *total = tot;
    return woff + incl - v;
}

// Decoupled look-back, split in two so the wait can be hidden behind useful work:
//   tile_publish   -- as soon as the tile's child count is known, publish it (AGG; tile 0
//                     publishes its inclusive PREFIX at once);
//   tile_lookback  -- later (after the trace-ahead intersections), one warp walks back over
//                     the predecessors 32 at a time until it meets an inclusive prefix.
// Tiles take their index from an atomic ticket, so every predecessor is resident or done.
RPX_DEV void tile_publish(unsigned long long* state, uint32_t tile, uint32_t total) {
    st_relaxed(&state[tile], (tile == 0 ? RPX_FLAG_PREFIX : RPX_FLAG_AGG) | (unsigned long long)total);
}

RPX_DEV unsigned long long tile_lookback(unsigned long long* state, uint32_t tile, uint32_t total) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    unsigned long long excl = 0;
    long long t0 = (long long)tile - 1;  // predecessor inspected by lane 0
    while (true) {
        const long long t = t0 - lane;
        const bool valid = (t >= 0);
        // only the entries in front of the nearest PREFIX have to be there: wait for exactly those
        unsigned long long s = valid ? ld_relaxed(&state[t]) : RPX_FLAG_AGG;
        unsigned pref;
        for (;;) {
            pref = __ballot_sync(0xffffffffu, valid && ((s >> 62) == 2));
            const unsigned have = __ballot_sync(0xffffffffu, (s >> 62) != 0);
            const unsigned need = pref ? ((1u << (__ffs(pref) - 1)) - 1u) : 0xffffffffu;
            if ((have & need) == need) break;
            if ((s >> 62) == 0) s = ld_relaxed(&state[t]);
        }
        unsigned long long v = valid ? (s & RPX_VAL_MASK) : 0ull;
        if (pref) {
            const int first = __ffs(pref) - 1;  // nearest predecessor holding a full prefix
            if (lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (pref) break;  // tile 0 always publishes a prefix, so this is always reached
        t0 -= 32;
    }
    if (lane == 0) st_relaxed(&state[tile], RPX_FLAG_PREFIX | (excl + total));
    return excl;
}

// Two-level look-back (k_shade).  With a persistent grid ~600 tiles are in flight at once and they
// run almost in lock step, so the nearest predecessor holding an inclusive PREFIX is usually a whole
// wave back: the flat look-back above walks ~19 windows of 32 tiles, one L2 round trip each.  Here
// every tile also adds its aggregate to the AGG word of its GROUP of 32 consecutive tiles
// (bits 61..56 = tiles that have contributed, low 48 bits = their sum; one fire-and-forget atomic),
// and the last tile of a group stores the group's inclusive PREFIX in a second word.  A look-back is
// then three loads in flight together: the predecessors inside the own group, and the AGG / PREFIX
// words of the 32 preceding GROUPS (1024 tiles) -- one round trip unless a predecessor has not
// published yet.
// State layout: [n_tiles tile words][n_groups AGG words][n_groups PREFIX words], all zero at launch.
#ifndef RPX_LOOKBACK_GROUPS
#define RPX_LOOKBACK_GROUPS 1
#endif
#ifndef RPX_GROUPS_GAUSS
#define RPX_GROUPS_GAUSS 0
#endif
__host__ __device__ inline size_t rpx_state_words(size_t n_tiles) {
#if RPX_LOOKBACK_GROUPS
    return n_tiles + 2 * ((n_tiles + 31) / 32);
#else
    return n_tiles;
#endif
}
#define RPX_GROUP_ONE (1ull << 56)
#define RPX_GROUP_SUM_MASK ((1ull << 48) - 1)

RPX_DEV void tile_publish_grouped(unsigned long long* state, unsigned long long* gagg, uint32_t tile, uint32_t total) {
    st_relaxed(&state[tile], (tile == 0 ? RPX_FLAG_PREFIX : RPX_FLAG_AGG) | (unsigned long long)total);
    atomicAdd(&gagg[tile >> 5], RPX_GROUP_ONE | (unsigned long long)total);
}

RPX_DEV unsigned long long tile_lookback_grouped(unsigned long long* state, unsigned long long* gagg,
                                                 unsigned long long* gpre, uint32_t tile, uint32_t total) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = tile >> 5, j = tile & 31u;
    unsigned long long excl = 0;
    if (tile != 0) {
        // (A) predecessors inside the own group: lane l looks at tile - 1 - l
        const bool own = (uint32_t)lane < j;
        const unsigned long long* pa = &state[tile - 1 - (own ? lane : 0)];
        // (B) preceding groups: lane l looks at group g - 1 - l; before group 0 = an empty PREFIX
        long long gi = (long long)g - 1 - lane;
        unsigned long long a = own ? ld_relaxed(pa) : RPX_FLAG_AGG;
        unsigned long long bp = (gi >= 0) ? ld_relaxed(&gpre[gi]) : RPX_FLAG_PREFIX;
        unsigned long long ba = (gi >= 0) ? ld_relaxed(&gagg[gi]) : 0ull;
        // only the entries in front of the nearest PREFIX have to be there: wait for exactly those
        unsigned pref_a;
        for (;;) {
            pref_a = __ballot_sync(0xffffffffu, own && (a >> 62) == 2);
            const unsigned have = __ballot_sync(0xffffffffu, (a >> 62) != 0);
            const unsigned need = pref_a ? ((1u << (__ffs(pref_a) - 1)) - 1u) : 0xffffffffu;
            if ((have & need) == need) break;
            if ((a >> 62) == 0) a = ld_relaxed(pa);
        }
        unsigned long long v = own ? (a & RPX_VAL_MASK) : 0ull;
        if (pref_a) {
            if (lane > __ffs(pref_a) - 1) v = 0;
        }
        bool done = (pref_a != 0);
        while (!done) {
            // a group is usable once its PREFIX is stored or all 32 of its tiles have contributed
            unsigned pref_b;
            for (;;) {
                const bool is_pref = (bp >> 62) == 2;
                const bool usable = is_pref || (ba >> 56) == 32ull;
                pref_b = __ballot_sync(0xffffffffu, is_pref);
                const unsigned have = __ballot_sync(0xffffffffu, usable);
                const unsigned need = pref_b ? ((1u << (__ffs(pref_b) - 1)) - 1u) : 0xffffffffu;
                if ((have & need) == need) break;
                if (!usable) {
                    bp = ld_relaxed(&gpre[gi]);
                    ba = ld_relaxed(&gagg[gi]);
                }
            }
            unsigned long long w = ((bp >> 62) == 2) ? (bp & RPX_VAL_MASK) : (ba & RPX_GROUP_SUM_MASK);
            if (pref_b) {
                if (lane > __ffs(pref_b) - 1) w = 0;
                done = true;
            }
            v += w;
            gi -= 32;
            if (!done) {
                bp = (gi >= 0) ? ld_relaxed(&gpre[gi]) : RPX_FLAG_PREFIX;
                ba = (gi >= 0) ? ld_relaxed(&gagg[gi]) : 0ull;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl = v;
    }
    if (lane == 0) {
        st_relaxed(&state[tile], RPX_FLAG_PREFIX | (excl + total));
        if (j == 31u) st_relaxed(&gpre[g], RPX_FLAG_PREFIX | (excl + total));
    }
    return excl;
}

// ------------------------------------------------------------------ ordered compaction of a filter
// k_capture / k_select keep a subset of a collection IN INPUT ORDER.  The per-ray work of such a filter
// is tiny (one plane test, or one compare), so the tile must be large or the kernel is bound by CTA
// scheduling and look-backs (measured on B200 with 128-ray CTAs: 0.47 ms per launch of 2.25e6 gausslets,
// 6x the HBM time): a CTA of RPX_TILE threads takes RPX_FILTER_R consecutive sub-tiles of RPX_TILE rays
// (coalesced loads), counts the kept rays per (sub-tile, warp) with ballots, scans those RPX_FILTER_R x
// warps counts once, and does ONE publish + grouped look-back per RPX_FILTER_R * RPX_TILE rays.
#define RPX_FILTER_R 8
#define RPX_FILTER_TILE (RPX_FILTER_R * RPX_TILE)
struct FilterSlots {
    unsigned long long pos[RPX_FILTER_R];  // output position of the thread's ray of sub-tile r (if kept)
    unsigned long long begin, end;         // the tile's records are out[begin, end)
};
// keep[r]: this thread's ray of sub-tile r is kept.  state: rpx_state_words(n_tiles) zeroed words.
// Returns false for the threads of a CTA that has nothing to copy.
RPX_DEV void filter_positions(const bool (&keep)[RPX_FILTER_R], unsigned long long* tile_state, uint32_t tile,
                              uint32_t n_tiles, unsigned long long base0, FilterSlots* out) {
    constexpr int W = RPX_TILE / 32;
    __shared__ uint32_t s_cnt[RPX_FILTER_R * W];
    __shared__ uint32_t s_off[RPX_FILTER_R * W];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t rank[RPX_FILTER_R];
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++) {
        const unsigned b = __ballot_sync(0xffffffffu, keep[r]);
        rank[r] = (uint32_t)__popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[r * W + warp] = (uint32_t)__popc(b);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the RPX_FILTER_R * W counts in (sub-tile, warp) order = ray order
        static_assert(RPX_FILTER_R * W <= 64, "two entries per lane");
        const int k0 = 2 * lane, k1 = 2 * lane + 1;
        const uint32_t c0 = k0 < RPX_FILTER_R * W ? s_cnt[k0] : 0u, c1 = k1 < RPX_FILTER_R * W ? s_cnt[k1] : 0u;
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t excl = incl - (c0 + c1);
        if (k0 < RPX_FILTER_R * W) s_off[k0] = excl;
        if (k1 < RPX_FILTER_R * W) s_off[k1] = excl + c0;
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long* gagg = tile_state + n_tiles;
        unsigned long long* gpre = gagg + (n_tiles + 31) / 32;
        if (lane == 0) tile_publish_grouped(tile_state, gagg, tile, total);
        const unsigned long long before = tile_lookback_grouped(tile_state, gagg, gpre, tile, total);
        if (lane == 0) {
            s_prefix = before;
            s_total = total;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++) out->pos[r] = base0 + s_prefix + s_off[r * W + warp] + rank[r];
    out->begin = base0 + s_prefix;
    out->end = base0 + s_prefix + s_total;
}

// Copy-out of a filter tile: the kept rays' source indices (relative to the tile) are scattered to shared
// memory by output slot, then ALL threads of the CTA copy row by row with consecutive threads on
// consecutive output slots -- coalesced stores, every lane busy whatever the hit pattern, and each thread
// has a dozen independent loads in flight.  (The first version let every kept ray's own thread copy its 86
// values: 2.4 ms for the 2e6 captured gausslets of a 4e6-gausslet launch, 1.1 TB/s; see profiles/r02_notes.md.)
struct FilterStage {
    uint16_t src[RPX_FILTER_TILE];  // slot -> ray index inside the tile
};

// ------------------------------------------------------------------ k_capture
// select_ray_intersections / select_gausslet_intersections (ctracer.pyx:1981-2058) for ONE
// collection: every ray is re-intersected between its origin and origin + direction * length
// with the capture FaceList (S holds only that face list); rays that hit are appended to `out`
// in input order after the *d_base records already captured from earlier collections.  The copy
// carries length = hit distance, end_face_idx = the capture face's idx and the re-based wavelength
// index (:2004-2006, 2011-2014).  Grid = ceil(n / RPX_FILTER_TILE) CTAs.
template <bool GAUSS, int FC, bool SS>
__global__ void __launch_bounds__(RPX_TILE, 4 * 128 / RPX_TILE)
k_capture(DevScene S, Soa in, Soa out, unsigned long long* tile_state, uint32_t* tile_counter,
          const unsigned long long* d_base, unsigned long long* d_next, uint32_t wl_offset, const uint32_t* wl_map,
          const uint32_t* face_ids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_tile;
    stage_scene<SS>(S, smem);
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);  // ticket order = start order: look-back cannot deadlock
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n_tiles = (uint32_t)((in.n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);
    const unsigned long long first = (unsigned long long)tile * RPX_FILTER_TILE + threadIdx.x;
    const unsigned long long cap = in.cap, ocap = out.cap;
    double len[RPX_FILTER_R];
    uint32_t face[RPX_FILTER_R];
    bool hit[RPX_FILTER_R];
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++) {
        const unsigned long long i = first + (unsigned long long)r * RPX_TILE;
        len[r] = 0.0;
        face[r] = RPX_NO_FACE;
        if (i < in.n) {
            vec3 o = v3(in.f[F_OX * cap + i], in.f[F_OY * cap + i], in.f[F_OZ * cap + i]);
            vec3 d = v3(in.f[F_DX * cap + i], in.f[F_DY * cap + i], in.f[F_DZ * cap + i]);
            nearest_hit<FC>(S, o, d, in.f[F_LEN * cap + i], -1, &len[r], &face[r]);
        }
        hit[r] = (face[r] != RPX_NO_FACE);
    }
    FilterSlots slots;
    filter_positions(hit, tile_state, tile, n_tiles, *d_base, &slots);
    if (threadIdx.x == 0 && tile == n_tiles - 1) *d_next = slots.end;
    const uint32_t total = (uint32_t)(slots.end - slots.begin);
    if (total == 0) return;  // uniform per CTA
    __shared__ FilterStage st;
    __shared__ double s_len[RPX_FILTER_TILE];
    __shared__ uint32_t s_face[RPX_FILTER_TILE];
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++)
        if (hit[r]) {
            const uint32_t ls = (uint32_t)(slots.pos[r] - slots.begin);
            st.src[ls] = (uint16_t)(r * RPX_TILE + threadIdx.x);
            s_len[ls] = len[r];
            s_face[ls] = face_ids ? face_ids[face[r]] : face[r];
        }
    __syncthreads();
    const unsigned long long tile0 = (unsigned long long)tile * RPX_FILTER_TILE;
    for (uint32_t sl = threadIdx.x; sl < total; sl += RPX_TILE) {
        const unsigned long long i = tile0 + st.src[sl];
        const unsigned long long pos = slots.begin + sl;
#pragma unroll
        for (int fld = 0; fld < NF; fld++)
            out.f[(unsigned long long)fld * ocap + pos] = (fld == F_LEN) ? s_len[sl] : in.f[(unsigned long long)fld * cap + i];
#pragma unroll
        for (int fld = 0; fld < NU; fld++) {
            uint32_t v = (fld == U_ENDFACE) ? s_face[sl] : in.u[(unsigned long long)fld * cap + i];
            if (fld == U_WL) {
                v += wl_offset;
                if (wl_map) v = wl_map[v];
            }
            out.u[(unsigned long long)fld * ocap + pos] = v;
        }
        if (GAUSS) {
#pragma unroll 12
            for (int fld = 0; fld < NP; fld++)
                out.p[(unsigned long long)fld * ocap + pos] = in.p[(unsigned long long)fld * cap + i];
        }
    }
}

// ------------------------------------------------------------------ child staging
// The children of one tile (<= 2 * RPX_TILE) are assembled in shared memory, field-major
// like the SoA generation buffer: cs[field * RPX_SLOTS + slot], cu[field * RPX_SLOTS + slot].
// This (a) frees the registers holding the children while their nearest hits are searched,
// (b) lets ANY thread of the block intersect ANY child (parents with 0 children help those
// with 2), and (c) makes the final global stores fully coalesced (slot == consecutive
// addresses) instead of stride-2.
// (Tried and measured slower, see profiles/r01_notes.md: thread-fixed staging slots with a
// compaction map; the look-back before the trace-ahead instead of after it.)
#define RPX_SLOTS (2 * RPX_TILE)
// Lean staging (RPX_LEAN_STAGE=1, what ships): the two children of a parent share origin, normal, E_vector, phase,
// accumulated path, wavelength index and ray ident -- stored ONCE per parent (11 doubles + 2 words x RPX_TILE), only
// direction, n, E1, E2, length, type, end face per child (10 doubles + 2 words x RPX_SLOTS), plus a byte map
// slot -> parent.  35 KB instead of 47 KB per CTA: at the same 4 CTAs / SM the shared-memory carve-out drops from 228
// to 164 KB and the L1 -- where the spill slots live -- triples.  Measured on B200 (profiles/r02_notes.md section 9):
// +0.9 .. +3.1 % on every plain-ray workload, +4.3 % on gausslets (2.776e9 -> 2.895e9 seg/s) once the parabasal loops
// were rolled; on the fully unrolled, instruction-cache-bound gausslet kernel the same change had measured -13 %, and
// at 5 CTAs / SM (96 registers, what round 1 prepared it for) it loses to the spills.  RPX_LEAN_STAGE=0 builds the
// full staging (every field per child slot).
#ifndef RPX_LEAN_STAGE
#define RPX_LEAN_STAGE 1
#endif
#if RPX_LEAN_STAGE
enum { LP_OX = 0, LP_OY, LP_OZ, LP_NX, LP_NY, LP_NZ, LP_EX, LP_EY, LP_EZ, LP_PHASE, LP_APATH, LP_NF = 11 };
enum { LC_DX = 0, LC_DY, LC_DZ, LC_NR, LC_NI, LC_E1R, LC_E1I, LC_E2R, LC_E2I, LC_LEN, LC_NF = 10 };
enum { LPU_WL = 0, LPU_IDENT, LPU_NU = 2 };
enum { LCU_TYPE = 0, LCU_ENDFACE, LCU_NU = 2 };
#define RPX_LEAN_PAR_F 0                                                   /* doubles [LP_NF][RPX_TILE]   */
#define RPX_LEAN_CH_F (RPX_LEAN_PAR_F + LP_NF * RPX_TILE * 8)              /* doubles [LC_NF][RPX_SLOTS]  */
#define RPX_LEAN_PAR_U (RPX_LEAN_CH_F + LC_NF * RPX_SLOTS * 8)             /* words   [LPU_NU][RPX_TILE]  */
#define RPX_LEAN_CH_U (RPX_LEAN_PAR_U + LPU_NU * RPX_TILE * 4)             /* words   [LCU_NU][RPX_SLOTS] */
#define RPX_LEAN_MAP (RPX_LEAN_CH_U + LCU_NU * RPX_SLOTS * 4)              /* bytes   [RPX_SLOTS]         */
#define RPX_STAGE_BYTES (((RPX_LEAN_MAP + RPX_SLOTS) + 15) / 16 * 16)
struct LeanStage {
    double* pf;         // per-parent doubles
    double* cf;         // per-child doubles
    uint32_t* pu;       // per-parent words
    uint32_t* cu;       // per-child words
    unsigned char* map; // slot -> parent (thread) index
};
RPX_DEV LeanStage lean_stage(unsigned char* smem) {
    LeanStage L;
    L.pf = reinterpret_cast<double*>(smem + RPX_LEAN_PAR_F);
    L.cf = reinterpret_cast<double*>(smem + RPX_LEAN_CH_F);
    L.pu = reinterpret_cast<uint32_t*>(smem + RPX_LEAN_PAR_U);
    L.cu = reinterpret_cast<uint32_t*>(smem + RPX_LEAN_CH_U);
    L.map = smem + RPX_LEAN_MAP;
    return L;
}
RPX_DEV void lean_stage_parent(const LeanStage& L, uint32_t p, const Kids& k, uint32_t wl, uint32_t ident) {
    double* f = L.pf + p;
    f[LP_OX * RPX_TILE] = k.origin.x;
    f[LP_OY * RPX_TILE] = k.origin.y;
    f[LP_OZ * RPX_TILE] = k.origin.z;
    f[LP_NX * RPX_TILE] = k.normal.x;
    f[LP_NY * RPX_TILE] = k.normal.y;
    f[LP_NZ * RPX_TILE] = k.normal.z;
    f[LP_EX * RPX_TILE] = k.evec.x;
    f[LP_EY * RPX_TILE] = k.evec.y;
    f[LP_EZ * RPX_TILE] = k.evec.z;
    f[LP_PHASE * RPX_TILE] = k.phase;
    f[LP_APATH * RPX_TILE] = k.apath;
    L.pu[LPU_WL * RPX_TILE + p] = wl;
    L.pu[LPU_IDENT * RPX_TILE + p] = ident;
}
RPX_DEV void lean_stage_child(const LeanStage& L, uint32_t slot, uint32_t p, const Kid& c) {
    double* f = L.cf + slot;
    f[LC_DX * RPX_SLOTS] = c.dir.x;
    f[LC_DY * RPX_SLOTS] = c.dir.y;
    f[LC_DZ * RPX_SLOTS] = c.dir.z;
    f[LC_NR * RPX_SLOTS] = c.n.re;
    f[LC_NI * RPX_SLOTS] = c.n.im;
    f[LC_E1R * RPX_SLOTS] = c.e1.re;
    f[LC_E1I * RPX_SLOTS] = c.e1.im;
    f[LC_E2R * RPX_SLOTS] = c.e2.re;
    f[LC_E2I * RPX_SLOTS] = c.e2.im;
    L.cu[LCU_TYPE * RPX_SLOTS + slot] = c.type;
    L.map[slot] = (unsigned char)p;
}
#else
#define RPX_STAGE_BYTES (RPX_SLOTS * (NF * 8 + NU * 4))
#endif

RPX_DEV void stage_child(double* cs, uint32_t* cu, uint32_t slot, const Kids& k, const Kid& c, uint32_t wl,
                         uint32_t parent, uint32_t ident) {
    double* f = cs + slot;
    f[F_OX * RPX_SLOTS] = k.origin.x;
    f[F_OY * RPX_SLOTS] = k.origin.y;
    f[F_OZ * RPX_SLOTS] = k.origin.z;
    f[F_DX * RPX_SLOTS] = c.dir.x;
    f[F_DY * RPX_SLOTS] = c.dir.y;
    f[F_DZ * RPX_SLOTS] = c.dir.z;
    f[F_NX * RPX_SLOTS] = k.normal.x;
    f[F_NY * RPX_SLOTS] = k.normal.y;
    f[F_NZ * RPX_SLOTS] = k.normal.z;
    f[F_EX * RPX_SLOTS] = k.evec.x;
    f[F_EY * RPX_SLOTS] = k.evec.y;
    f[F_EZ * RPX_SLOTS] = k.evec.z;
    f[F_NR * RPX_SLOTS] = c.n.re;
    f[F_NI * RPX_SLOTS] = c.n.im;
    f[F_E1R * RPX_SLOTS] = c.e1.re;
    f[F_E1I * RPX_SLOTS] = c.e1.im;
    f[F_E2R * RPX_SLOTS] = c.e2.re;
    f[F_E2I * RPX_SLOTS] = c.e2.im;
    f[F_PHASE * RPX_SLOTS] = k.phase;
    f[F_APATH * RPX_SLOTS] = k.apath;
    uint32_t* u = cu + slot;
    u[U_WL * RPX_SLOTS] = wl;
    u[U_PARENT * RPX_SLOTS] = parent;
    u[U_IDENT * RPX_SLOTS] = ident;
    u[U_TYPE * RPX_SLOTS] = c.type;
}

// ------------------------------------------------------------------ k_shade
// One launch per generation.  Per tile of RPX_TILE parents (tile index from an atomic ticket):
//   1. load the parent records (coalesced SoA), orientation + material -> <= 2 children each
//   2. block scan of the child counts; publish the tile aggregate (no waiting)
//   3. stage the children in shared memory in emission order
//   4. trace ahead: nearest hit of every staged child (balanced over the whole block) -- what
//      trace_segment_c computes at the top of the NEXT generation (ctracer.pyx:2084-2104),
//      so the next generation needs no intersect pass
//   5. decoupled look-back for the tile's global offset (the predecessors published while
//      step 4 ran, so it rarely spins)
//   6. coalesced copy of the staged children to the next generation's SoA buffer
//   7. gausslets only: parabasal children (ctracer.pyx:2375-2385), written directly
// Design notes (measured on B200, profiles/r01_notes.md): persistent CTAs fed by per-field TMA
// bulk copies and warp-granular tiles were both tried and were slower -- 22 sub-KB bulk copies
// per tile serialise in the TMA unit, and 4x more tiles mean 4x more look-backs.
#ifndef RPX_MIN_BLOCKS
#define RPX_MIN_BLOCKS 4
#endif
#ifndef RPX_BULK_PREFETCH
#define RPX_BULK_PREFETCH 1
#endif
#ifndef RPX_MIN_BLOCKS_G
#define RPX_MIN_BLOCKS_G RPX_MIN_BLOCKS
#endif
// Gausslet tiles take their next ticket at the END of the current tile (plain rays: at its start, so that the
// next tile's rows can be L2-prefetched).  A gausslet tile is ~100 us long: a CTA that already holds ticket T
// while it still works on its previous tile keeps every tile > T polling in the look-back for T's aggregate
// (ncu: 20 % of the executed instructions of k_shade<GAUSS> were that poll, and launches fell into a slow mode
// -- 592 / 1161 / 2334 us instead of 487 / 924 / 1821 us for the 1e6 / 2e6 / 4e6-gausslet generations of the
// Michelson trace).  With the late ticket a tile publishes its aggregate one material phase after taking its
// ticket, before any later tile can reach its look-back.  Measured (profiles/r02_notes.md): +9 % on one box,
// neutral on another, never the slow mode under ncu.
#ifndef RPX_TILE_COMPACT
#define RPX_TILE_COMPACT 0  // tile-local compaction of the hit rays: measured (profiles/r02_notes.md section 7), +1.4 % on
                            // prisms, -0.6 % on the achromat / Michelson / grating, -4.6 % on gausslets -> off; -DRPX_TILE_COMPACT=1 builds it
#endif
#ifndef RPX_PARA_FIRST
#define RPX_PARA_FIRST 1
#endif
#ifndef RPX_TICKET_END_G
#define RPX_TICKET_END_G 1
#endif
// Gausslets: unroll factor of the two six-ray parabasal loops.  Fully unrolled (6) they were 110 KB of the kernel's
// 205 KB of SASS (six inlined copies of the face intersection, twelve of the parabasal material code); instructions
// beyond the 32 KB L1.5 stream from L2, and with the 16 warps of an SM each in another phase of its tile ncu showed
// 1.5 - 10 % of the stall samples as "no instruction" -- how many depended on the code LAYOUT: two changes that
// removed work from the second loop made the kernel 16 % slower.  Rolled, the loop bodies are re-used from the
// instruction cache.  Measured on B200 (Michelson, 1e6 gausslets, profiles/r02_notes.md section 9):
// unroll 6: 2.616e9 seg/s, 1: 2.648e9, 2: 2.770e9, 3: 2.729e9 (both loops alike); software-pipelining the loads of ray j + 1
// on top: +1.4 % at unroll 1, -4 % at unroll 2 (registers).
// Chosen per loop (first = intersections, second = parabasal children), same workload: 1/2: 2.681e9, 2/1: 2.740e9,
// 2/2: 2.773e9, 2/3: 2.696e9, **3/2: 2.800e9 (shipped)**.
#ifndef RPX_PARA_UNROLL
#define RPX_PARA_UNROLL 3
#endif
#ifndef RPX_PARA_UNROLL2
#define RPX_PARA_UNROLL2 2
#endif
static constexpr int kParaUnroll = RPX_PARA_UNROLL;    // first loop (intersections)
static constexpr int kParaUnroll2 = RPX_PARA_UNROLL2;  // second loop (parabasal children)
// One-shot parent reads through L2 only (ld.global.cg): the 22 / 82 rows of a tile are used once, while the kernel's
// spill slots (150 - 900 B / thread) want to stay in what is left of the L1 beside the shared-memory carve-out.
#ifndef RPX_LDCG
#define RPX_LDCG 0
#endif
#if RPX_LDCG
#define RPX_LD(p) __ldcg(p)
#else
#define RPX_LD(p) (*(p))
#endif
template <bool GAUSS, int FC, uint32_t MM, bool SS>
__global__ void __launch_bounds__(RPX_TILE, GAUSS ? RPX_MIN_BLOCKS_G : RPX_MIN_BLOCKS)
k_shade(DevScene S, Soa in, Soa out, double max_length, unsigned long long* tile_state,
        uint32_t* tile_counter, unsigned long long* d_count, uint32_t* face_counts, uint32_t n_tiles,
        int ahead_face, const unsigned long long* n_dev, unsigned long long* h_count, const uint32_t* hits_in,
        uint32_t* hits_out, const uint32_t* miss_in, uint32_t* miss_out) {
    // hits_in != NULL: set by the launch that built this generation iff its trace-ahead found ANY hit.  A
    // generation nobody hit anything in (the last one of every finite trace: 15 % of the achromat and
    // Michelson steps went into reading it, ncu r02_div_*.csv) has no child, no count and no write-back
    // left to produce: the whole grid leaves at once.  hits_out: the same flag for the generation built here.
    if (hits_in != nullptr && *hits_in == 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            *d_count = 0ull;
            if (h_count) {
                *h_count = 0ull;
                __threadfence_system();
            }
        }
        return;
    }
    // n_dev != NULL: the parent count is not known on the host yet (the launch was enqueued
    // before the previous generation's kernel finished): the grid covers an upper bound and the
    // real count is read here; surplus CTAs leave at once.
    // ahead_face: -1 trace the children ahead against every face (non-sequential mode);
    //             >= 0 only against that face (next step of a face sequence);
    //             -2 leave them untraced (last step of a sequence: the reference appends that
    //                generation as the material left it -- length INF, the parent's end_face_idx)
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[RPX_TILE / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_piece[FC == RPX_FC_MESH ? RPX_SLOTS : 1];  // facet records of the staged children's hits
    // dynamic shared memory: [child staging][scene copy]
#if RPX_LEAN_STAGE
    const LeanStage L = lean_stage(smem);
#else
    double* cs = reinterpret_cast<double*>(smem);
    uint32_t* cu = reinterpret_cast<uint32_t*>(smem + RPX_SLOTS * NF * 8);
#endif
    stage_scene<SS>(S, smem + RPX_STAGE_BYTES);  // once per (persistent) CTA
    const unsigned long long n_in = n_dev ? *n_dev : in.n;
    const uint32_t n_tiles_real = (uint32_t)((n_in + RPX_TILE - 1) / RPX_TILE);
    constexpr bool kGrouped = RPX_LOOKBACK_GROUPS && (!GAUSS || RPX_GROUPS_GAUSS);
    const unsigned long long cap = in.cap;
    // TILE-LOCAL COMPACTION OF THE HIT RAYS (the north star's "binning / compaction pass", done inside the
    // tile so that it costs no HBM traffic and no second permutation).  In branching scenes half of a
    // generation can be rays that left the system (prisms: every reflected partner of a transmitted ray);
    // they sit interleaved with the rays that hit something, so every warp runs the material code half
    // empty (ncu: 23.4 of 32 threads per instruction on the prism scene).  When the launch that built this
    // generation counted >= 1/8 rays without a hit (*miss_in), the tile first reads end_face_idx alone,
    // ballot-compacts the indices of the hit rays IN ORDER into shared memory, and thread t then takes the
    // t-th hit ray: the material code runs on full warps, the rest of the CTA's warps skip it, and because
    // the compaction keeps the order, the parent-ordered child slots still come out of the same thread-order
    // scan.  Uniform per launch, so scenes without misses (achromat, Michelson) keep the direct path.
    // MEASURED on B200 and switched off (RPX_TILE_COMPACT=0): the kernel is latency-bound, not issue-bound, so
    // full warps in the material code buy +1.4 % on prisms while the bookkeeping costs 0.6 - 4.6 % elsewhere.
#if RPX_TILE_COMPACT
    __shared__ uint32_t s_hcnt[RPX_TILE / 32];
    __shared__ unsigned char s_src[RPX_TILE];
    __shared__ uint32_t s_nmiss;
    const bool compact = miss_in != nullptr && (unsigned long long)(*miss_in) * 8ull >= n_in && n_in > 0;
    if (threadIdx.x == 0) s_nmiss = 0u;
#endif
  // PERSISTENT CTA: the grid is one wave of resident CTAs; each pulls tiles from the ticket
  // counter until the (device-resident) tile count is exhausted.
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (;;) {
    __syncthreads();  // s_tile published; previous tile's staging buffer fully consumed
    const uint32_t tile = s_tile;
    if (tile >= n_tiles_real) break;  // uniform per CTA; tickets are dense, so tiles [0, real) all run
    unsigned long long i = (unsigned long long)tile * RPX_TILE + threadIdx.x;  // the parent this thread shades
    bool have = i < n_in;
#if RPX_TILE_COMPACT
    if (compact) {
        const uint32_t f0 = have ? in.u[U_ENDFACE * cap + i] : RPX_NO_FACE;
        const bool h0 = (f0 != RPX_NO_FACE);
        const unsigned b = __ballot_sync(0xffffffffu, h0);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_hcnt[warp] = (uint32_t)__popc(b);
        __syncthreads();
        uint32_t off = 0, n_hit = 0;
#pragma unroll
        for (int w = 0; w < RPX_TILE / 32; w++) {
            const uint32_t c = s_hcnt[w];
            if (w < warp) off += c;
            n_hit += c;
        }
        if (h0) s_src[off + (uint32_t)__popc(b & ((1u << lane) - 1u))] = (unsigned char)threadIdx.x;
        __syncthreads();
        have = threadIdx.x < n_hit;
        i = (unsigned long long)tile * RPX_TILE + (have ? (uint32_t)s_src[threadIdx.x] : 0u);
    }
#endif
    // (s_tile is next written by thread 0 after the barrier inside the block scan: no barrier needed here)
    uint32_t next_tile = 0;
    // take the NEXT ticket now (its latency hides behind this tile's work) ...
    // (gausslets: at the END of the tile, see RPX_TICKET_END_G)
    constexpr bool kLateTicket = GAUSS && RPX_TICKET_END_G;
    if (threadIdx.x == 0 && !kLateTicket) next_tile = atomicAdd(tile_counter, 1u);

    Kids k;
    k.has_a = false;
    k.has_b = false;
    uint32_t wl = 0, ident = 0;
    uint32_t face_idx = RPX_NO_FACE;
    double plen[RPX_NPARA];
    HitAux paux[FC == RPX_FC_MESH ? RPX_NPARA : 1];  // piece_idx / uv of the parabasal hits (mesh, UV patch faces)
    bool hit = false;
    RayIn r;
    // gausslets (RPX_PARA_FIRST, shipped): the parabasal rays are intersected BEFORE orientation + material -- their six
    // serial (origin, direction) round trips then overlap the base ray's own loads, the 31 doubles of the children
    // are not live across the six intersections, and a dropped gausslet (Q16) skips the material.  Measured on three
    // B200s: +3.5 / +4.5 / +6.3 % on the Michelson gausslets (profiles/r02_notes.md section 8).
    constexpr int kParaFirst = GAUSS ? RPX_PARA_FIRST : 0;
    auto load_r = [&]() {
        const double* fi = in.f + i;
        const uint32_t* ui = in.u + i;
        r.o = v3(RPX_LD(fi + F_OX * cap), RPX_LD(fi + F_OY * cap), RPX_LD(fi + F_OZ * cap));
        r.d = v3(RPX_LD(fi + F_DX * cap), RPX_LD(fi + F_DY * cap), RPX_LD(fi + F_DZ * cap));
        r.e = v3(RPX_LD(fi + F_EX * cap), RPX_LD(fi + F_EY * cap), RPX_LD(fi + F_EZ * cap));
        r.n = cx(RPX_LD(fi + F_NR * cap), RPX_LD(fi + F_NI * cap));
        r.e1 = cx(RPX_LD(fi + F_E1R * cap), RPX_LD(fi + F_E1I * cap));
        r.e2 = cx(RPX_LD(fi + F_E2R * cap), RPX_LD(fi + F_E2I * cap));
        r.len = RPX_LD(fi + F_LEN * cap);
        r.phase = RPX_LD(fi + F_PHASE * cap);
        r.apath = RPX_LD(fi + F_APATH * cap);
        r.wl = wl = RPX_LD(ui + U_WL * cap);
        r.ident = ident = RPX_LD(ui + U_IDENT * cap);
        r.type = RPX_LD(ui + U_TYPE * cap);
    };
    if (have) {
        // every load is issued before the first use: one DRAM round trip per tile, not two
        face_idx = RPX_LD(in.u + U_ENDFACE * cap + i);
        load_r();
        hit = (face_idx != RPX_NO_FACE);
    }
    if (hit) {
        const rpx_face* face = &S.faces[face_idx];
        {  // face.count += 1 (ctracer.pyx:2108), aggregated per warp and face
            const unsigned peers = __match_any_sync(__activemask(), face_idx);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1)
                atomicAdd(&face_counts[face_idx], (uint32_t)__popc(peers));
        }
        // trace_parabasal_rays, first loop (ctracer.pyx:2363-2373): every parabasal ray must hit the SAME face
        // (is_base_ray = 0); any miss drops the children (Q16).  A local function so that it can run before
        // (RPX_PARA_FIRST) or after the material.
        auto para_hits = [&]() -> bool {

            bool ok = true;
            const rpx_face_set* fs = &S.sets[face->face_set];
#pragma unroll kParaUnroll
            for (int j = 0; j < RPX_NPARA; j++) {
                plen[j] = max_length;
                if (ok) {
                    const double* pp = in.p + (unsigned long long)(j * NPF) * cap + i;
                    vec3 po = v3(RPX_LD(pp + (P_OX + 0) * cap), RPX_LD(pp + (P_OX + 1) * cap), RPX_LD(pp + (P_OX + 2) * cap));
                    vec3 pd = v3(RPX_LD(pp + (P_DX + 0) * cap), RPX_LD(pp + (P_DX + 1) * cap), RPX_LD(pp + (P_DX + 2) * cap));
                    vec3 ray_end = po + pd * max_length;
                    vec3 p1 = transform_pt(fs->inv_trans.m, po);
                    vec3 p2 = transform_pt(fs->inv_trans.m, ray_end);
                    HitAux pa;
                    pa.rec = -1;
                    pa.piece = 0;
                    pa.u = pa.v = 0.0;
                    double dist = face_intersect<FC>(S, face, p1, p2, 0, FC == RPX_FC_MESH ? &pa : nullptr);
                    if (FC == RPX_FC_MESH) paux[j] = pa;
                    if (face->tolerance < dist && dist < max_length) {
                        plen[j] = dist;
                        in.p[(unsigned long long)(j * NPF + P_LEN) * cap + i] = dist;  // parent write-back
                    } else {
                        ok = false;
                    }
                }
            }
            return ok;
        };
        bool ok = true;
        if (kParaFirst) ok = para_hits();
        if (ok) {
            vec3 point = r.o + r.d * r.len;
            vec3 onormal, otangent;
            HitAux aux;
            aux.rec = -1;
            aux.piece = 0;
            aux.u = aux.v = 0.0;
            if (FC == RPX_FC_MESH && (face->type == RPX_FACE_MESH || face->type == RPX_FACE_UVPATCH)) {
                // intersect_t.piece_idx (which triangle) / .uv (patch parameters) are not part of the ray
                // record: the hit is found again from the same inputs the trace-ahead / k_intersect pass
                // used, so it is the same hit
                // -- unless the launch that found it left the facet record in the side array
                const rpx_face_set* mfs = &S.sets[face->face_set];
                const vec3 q1 = transform_pt(mfs->inv_trans.m, r.o);
                const vec3 q2 = transform_pt(mfs->inv_trans.m, r.o + r.d * max_length);
                const uint32_t rec1 = in.piece ? in.piece[i] : 0u;
                if (rec1 != 0u)
                    face_aux_from_rec(S, face, q1, q2, (int)rec1 - 1, &aux);
                else
                    face_intersect<FC>(S, face, q1, q2, 1, &aux);
                if (aux.piece < 0) aux.piece = 0;
            }
            compute_orientation<FC>(S, face, point, &onormal, &otangent, &aux);
            material_eval<MM>(S, &S.mats[face->material], r, point, onormal, otangent, k);
            if (GAUSS && !kParaFirst && !para_hits()) {
                k.has_a = false;
                k.has_b = false;
            }
        }
    }

    const uint32_t parent = (uint32_t)i;
    // ---- 2. counts -> offsets inside the tile; publish the tile aggregate
    const uint32_t cnt = (k.has_a ? 1u : 0u) + (k.has_b ? 1u : 0u);
    uint32_t total;
    const uint32_t local = block_exclusive_scan(cnt, &total, s_warp);
    if (threadIdx.x == 0) {
        if (kGrouped)
            tile_publish_grouped(tile_state, tile_state + n_tiles, tile, total);
        else
            tile_publish(tile_state, tile, total);
        if (!kLateTicket) s_tile = next_tile;  // ... and hand it to the CTA (read after the next barrier)
    }
    // ---- 3. stage children in emission order (reflected, then transmitted)
    const uint32_t slot_a = local, slot_b = local + (k.has_a ? 1u : 0u);
#if RPX_LEAN_STAGE
    if (cnt) lean_stage_parent(L, threadIdx.x, k, wl, ident);
    if (k.has_a) lean_stage_child(L, slot_a, threadIdx.x, k.a);
    if (k.has_b) lean_stage_child(L, slot_b, threadIdx.x, k.b);
#else
    if (k.has_a) stage_child(cs, cu, slot_a, k, k.a, wl, parent, ident);
    if (k.has_b) stage_child(cs, cu, slot_b, k, k.b, wl, parent, ident);
#endif
    const uint32_t own_a = slot_a, own_b = slot_b;
    __syncthreads();
    {   // pull the next tile's parent records towards L2 while this tile computes
        const uint32_t nt = kLateTicket ? n_tiles_real : s_tile;  // late ticket: the next tile is not known yet
#if RPX_BULK_PREFETCH
        // one bulk L2 prefetch per field row (1 KB of doubles / 512 B of u32), issued by 22 threads of
        // the LAST warp (warp 0 runs the look-back): 22 instructions per tile instead of ~90
        if (nt < n_tiles_real && threadIdx.x >= RPX_TILE - 32 && threadIdx.x < RPX_TILE - 32 + 22) {
            const uint32_t q = threadIdx.x - (RPX_TILE - 32);
            const unsigned long long nb = (unsigned long long)nt * RPX_TILE;
            if (q < 18) {
                const uint32_t fld = q < 6 ? q : q + 3;  // every double field but the parent normal
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.f + (unsigned long long)fld * cap + nb),
                             "r"(RPX_TILE * 8));
            } else {
                const uint32_t w = q - 18, fld = w + (w ? 1u : 0u);  // WL, ENDFACE, IDENT, TYPE
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in.u + (unsigned long long)fld * cap + nb),
                             "r"(RPX_TILE * 4));
            }
        }
        if (false) {
#else
        if (nt < n_tiles_real) {
#endif
            const unsigned long long ni = (unsigned long long)nt * RPX_TILE + threadIdx.x;
            if ((threadIdx.x & 3) == 0) {  // one prefetch per 32-byte sector
                const int pf[18] = {F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_EX, F_EY, F_EZ, F_NR, F_NI,
                                    F_E1R, F_E1I, F_E2R, F_E2I, F_LEN, F_PHASE, F_APATH};
#pragma unroll
                for (int q = 0; q < 18; q++)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(in.f + (unsigned long long)pf[q] * cap + ni));
            }
            if ((threadIdx.x & 7) == 0) {
                const int pu[4] = {U_WL, U_ENDFACE, U_IDENT, U_TYPE};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(in.u + (unsigned long long)pu[q] * cap + ni));
            }
        }
    }
    // ---- 4. trace ahead
    bool any_hit = false;
#if RPX_TILE_COMPACT
    uint32_t n_miss = 0;
#endif
    if (ahead_face != -2) {
        for (uint32_t slot = threadIdx.x; slot < total; slot += RPX_TILE) {
#if RPX_LEAN_STAGE
            const double* fp = L.pf + L.map[slot];
            const double* fc = L.cf + slot;
            vec3 o = v3(fp[LP_OX * RPX_TILE], fp[LP_OY * RPX_TILE], fp[LP_OZ * RPX_TILE]);
            vec3 d = v3(fc[LC_DX * RPX_SLOTS], fc[LC_DY * RPX_SLOTS], fc[LC_DZ * RPX_SLOTS]);
#else
            const uint32_t src = slot;
            const double* f = cs + src;
            vec3 o = v3(f[F_OX * RPX_SLOTS], f[F_OY * RPX_SLOTS], f[F_OZ * RPX_SLOTS]);
            vec3 d = v3(f[F_DX * RPX_SLOTS], f[F_DY * RPX_SLOTS], f[F_DZ * RPX_SLOTS]);
#endif
            double len;
            uint32_t face;
            uint32_t hit_rec = 0;
            nearest_hit<FC>(S, o, d, max_length, ahead_face, &len, &face, &hit_rec);
            if (FC == RPX_FC_MESH) s_piece[slot] = hit_rec;
            any_hit = any_hit || (face != RPX_NO_FACE);
#if RPX_TILE_COMPACT
            if (face == RPX_NO_FACE) n_miss++;
#endif
#if RPX_LEAN_STAGE
            L.cf[LC_LEN * RPX_SLOTS + slot] = len;
            L.cu[LCU_ENDFACE * RPX_SLOTS + slot] = face;
#else
            cs[F_LEN * RPX_SLOTS + src] = len;
            cu[U_ENDFACE * RPX_SLOTS + src] = face;
#endif
        }
    } else {
        // untraced: sp_ray.length = INF (every material; gausslets: reset_length_c -> max_length),
        // end_face_idx still the copy of the parent's (the face that was just hit)
        const double untraced_len = GAUSS ? max_length : RPX_INF;
#if RPX_LEAN_STAGE
        if (k.has_a) {
            L.cf[LC_LEN * RPX_SLOTS + own_a] = untraced_len;
            L.cu[LCU_ENDFACE * RPX_SLOTS + own_a] = face_idx;
        }
        if (k.has_b) {
            L.cf[LC_LEN * RPX_SLOTS + own_b] = untraced_len;
            L.cu[LCU_ENDFACE * RPX_SLOTS + own_b] = face_idx;
        }
#else
        if (k.has_a) {
            cs[F_LEN * RPX_SLOTS + own_a] = untraced_len;
            cu[U_ENDFACE * RPX_SLOTS + own_a] = face_idx;
        }
        if (k.has_b) {
            cs[F_LEN * RPX_SLOTS + own_b] = untraced_len;
            cu[U_ENDFACE * RPX_SLOTS + own_b] = face_idx;
        }
#endif
    }
    // ---- 5. global offset of the tile
    if (threadIdx.x < 32) {
        const unsigned long long excl =
            kGrouped ? tile_lookback_grouped(tile_state, tile_state + n_tiles,
                                             tile_state + n_tiles + (n_tiles + 31) / 32, tile, total)
                     : tile_lookback(tile_state, tile, total);
        if (threadIdx.x == 0) {
            s_prefix = excl;
            if (tile == n_tiles_real - 1) {  // len(new_rays)
                *d_count = excl + total;
                if (h_count) {  // pipelined loop: straight into mapped pinned host memory, no copy op
                    *h_count = excl + total;
                    __threadfence_system();
                }
            }
        }
    }
#if RPX_TILE_COMPACT
    if (n_miss && miss_out != nullptr) atomicAdd(&s_nmiss, n_miss);
#endif
    const int tile_hit = __syncthreads_or(any_hit ? 1 : 0);
    if (threadIdx.x == 0) {
        if (tile_hit && hits_out != nullptr) *hits_out = 1u;  // idempotent, one store per tile
#if RPX_TILE_COMPACT
        if (miss_out != nullptr && s_nmiss) {                  // rays of the new generation that hit nothing
            atomicAdd(miss_out, s_nmiss);
            s_nmiss = 0u;  // next added to after the next tile's barriers
        }
#endif
    }
    const unsigned long long base = s_prefix;
    // ---- 6. coalesced copy-out: slot == consecutive addresses.  Two explicit passes (a tile has
    // at most 2 * RPX_TILE children), each a straight line of 26 independent LDS -> STG pairs.
    // (Letting the compiler unroll a generic slot loop bloated the kernel by ~90 KB of SASS; a
    // fully rolled loop cost 14 % of the stall samples in branch / index overhead.)
    {
        const unsigned long long ocap = out.cap;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const uint32_t slot = threadIdx.x + pass * RPX_TILE;
            if (slot < total) {
                double* dst = out.f + base + slot;
                uint32_t* dstu = out.u + base + slot;
#if RPX_LEAN_STAGE
                const uint32_t p = L.map[slot];
                const double* sp = L.pf + p;
                const double* sc = L.cf + slot;
                // SoA field <- (per-parent | per-child) staging row
                constexpr int kParRow[NF] = {LP_OX, LP_OY, LP_OZ, -1, -1, -1, LP_NX, LP_NY, LP_NZ, LP_EX, LP_EY, LP_EZ,
                                             -1, -1, -1, -1, -1, -1, -1, LP_PHASE, LP_APATH};
                constexpr int kChRow[NF] = {-1, -1, -1, LC_DX, LC_DY, LC_DZ, -1, -1, -1, -1, -1, -1,
                                            LC_NR, LC_NI, LC_E1R, LC_E1I, LC_E2R, LC_E2I, LC_LEN, -1, -1};
#pragma unroll
                for (int fld = 0; fld < NF; fld++)
                    dst[(unsigned long long)fld * ocap] =
                        kParRow[fld] >= 0 ? sp[kParRow[fld] * RPX_TILE] : sc[kChRow[fld] * RPX_SLOTS];
                dstu[(unsigned long long)U_WL * ocap] = L.pu[LPU_WL * RPX_TILE + p];
                dstu[(unsigned long long)U_PARENT * ocap] = tile * RPX_TILE + p;
                dstu[(unsigned long long)U_ENDFACE * ocap] = L.cu[LCU_ENDFACE * RPX_SLOTS + slot];
                dstu[(unsigned long long)U_IDENT * ocap] = L.pu[LPU_IDENT * RPX_TILE + p];
                dstu[(unsigned long long)U_TYPE * ocap] = L.cu[LCU_TYPE * RPX_SLOTS + slot];
#else
                const uint32_t from = slot;
                const double* src = cs + from;
#pragma unroll
                for (int fld = 0; fld < NF; fld++) dst[(unsigned long long)fld * ocap] = src[fld * RPX_SLOTS];
                const uint32_t* srcu = cu + from;
#pragma unroll
                for (int fld = 0; fld < NU; fld++) dstu[(unsigned long long)fld * ocap] = srcu[fld * RPX_SLOTS];
#endif
                if (FC == RPX_FC_MESH && out.piece && ahead_face != -2) out.piece[base + slot] = s_piece[slot];
            }
        }
    }

    if (GAUSS && cnt) {
        // trace_parabasal_rays, second loop (ctracer.pyx:2375-2385) + reset_length_c (:2280):
        // parabasal lengths of a new gausslet are max_length; the base ray's length slot already
        // holds its nearest-hit distance from the trace-ahead step, which is what the next
        // generation's intersect would have written over max_length anyway.
        const rpx_face* face = &S.faces[face_idx];
        const rpx_material* M = &S.mats[face->material];
        const unsigned long long ocap = out.cap;
        const unsigned long long pos_a = base + slot_a, pos_b = base + slot_b;
        const ParaSnell ps = para_snell_setup(S, M, wl);
#pragma unroll kParaUnroll2
        for (int j = 0; j < RPX_NPARA; j++) {
            const double* pp = in.p + (unsigned long long)(j * NPF) * cap + i;
            vec3 po = v3(RPX_LD(pp + (P_OX + 0) * cap), RPX_LD(pp + (P_OX + 1) * cap), RPX_LD(pp + (P_OX + 2) * cap));
            vec3 pd = v3(RPX_LD(pp + (P_DX + 0) * cap), RPX_LD(pp + (P_DX + 1) * cap), RPX_LD(pp + (P_DX + 2) * cap));
            vec3 ppoint = po + pd * plen[j];
            vec3 pn, pt;
            compute_orientation<FC>(S, face, ppoint, &pn, &pt, FC == RPX_FC_MESH ? &paux[j] : nullptr);
            vec3 nn = norm(pn);
            if (k.has_a) {
                vec3 dir = material_eval_para(S, M, wl, k.a.n.re, pd, ppoint, pn, pt, k.a.type, ps);
                double* q = out.p + (unsigned long long)(j * NPF) * ocap + pos_a;
                q[0 * ocap] = ppoint.x; q[1 * ocap] = ppoint.y; q[2 * ocap] = ppoint.z;
                q[3 * ocap] = dir.x; q[4 * ocap] = dir.y; q[5 * ocap] = dir.z;
                q[6 * ocap] = nn.x; q[7 * ocap] = nn.y; q[8 * ocap] = nn.z;
                q[9 * ocap] = max_length;
            }
            if (k.has_b) {
                vec3 dir = material_eval_para(S, M, wl, k.b.n.re, pd, ppoint, pn, pt, k.b.type, ps);
                double* q = out.p + (unsigned long long)(j * NPF) * ocap + pos_b;
                q[0 * ocap] = ppoint.x; q[1 * ocap] = ppoint.y; q[2 * ocap] = ppoint.z;
                q[3 * ocap] = dir.x; q[4 * ocap] = dir.y; q[5 * ocap] = dir.z;
                q[6 * ocap] = nn.x; q[7 * ocap] = nn.y; q[8 * ocap] = nn.z;
                q[9 * ocap] = max_length;
            }
        }
    }
    if (kLateTicket && threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  }  // persistent tile loop
}

}  // namespace rpx
