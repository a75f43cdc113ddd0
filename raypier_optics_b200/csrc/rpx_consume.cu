// rpx_consume.cu -- what happens to a trace AFTER the generation loop, on the device:
//   * terminal-ray selection (the rays a multi-GPU trace gathers at the end),
//   * device-resident AoS export / import (the send / receive buffers of that gather),
//   * rpx_trace_consume: the chunked trace whose generations never leave the GPU -- every chunk is
//     traced, filtered (terminal rays, capture plane), summed into a detector field and freed.  This is
//     how the BASELINE configs at 1e8 - 1e9 rays run: one generation of 1e9 gausslets is 668 GB, so
//     "return every generation" (core/tracer.py:39-45) cannot be the interface at that size; what the
//     reference's own post-trace consumers keep (probes.py:119-143 capture planes, fields.py:206-277
//     E-field planes, Face.count) can.
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <vector>

#include "../../include/rpx.h"
#include "rpx_internal.h"
#include "rpx_launch.h"

using namespace rpx;

namespace {

// ------------------------------------------------------------------ k_select
// Ordered compaction of the rays of one collection whose end_face_idx marks them terminal:
// RPX_NO_FACE (nothing was hit: ctracer.pyx:2086-2087 leaves (unsigned)-1) when `unterminated`, or a
// face with face_select[idx] != 0.  Same ordered compaction as k_capture (filter_positions); records are
// copied unchanged except for parent_idx, which gets `parent_offset` added (global numbering across the
// chunks of rpx_trace_consume).  *d_base = records already in `out`; the last tile writes *d_next.
template <bool GAUSS>
__global__ void __launch_bounds__(RPX_TILE, 4 * 128 / RPX_TILE)
k_select(Soa in, Soa out, unsigned long long* tile_state, uint32_t* tile_counter, const unsigned long long* d_base,
         unsigned long long* d_next, int unterminated, const unsigned char* face_select, uint32_t parent_offset,
         int copy) {
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);  // ticket order = start order: look-back cannot deadlock
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n_tiles = (uint32_t)((in.n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);
    const unsigned long long first = (unsigned long long)tile * RPX_FILTER_TILE + threadIdx.x;
    const unsigned long long cap = in.cap, ocap = out.cap;
    bool sel[RPX_FILTER_R];
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++) {
        const unsigned long long i = first + (unsigned long long)r * RPX_TILE;
        sel[r] = false;
        if (i < in.n) {
            const uint32_t face = in.u[U_ENDFACE * cap + i];
            sel[r] = (face == RPX_NO_FACE) ? (unterminated != 0) : (face_select != nullptr && face_select[face] != 0);
        }
    }
    FilterSlots slots;
    filter_positions(sel, tile_state, tile, n_tiles, *d_base, &slots);
    if (threadIdx.x == 0 && tile == n_tiles - 1) *d_next = slots.end;
    const uint32_t total = (uint32_t)(slots.end - slots.begin);
    if (!copy || total == 0) return;  // uniform per CTA
    __shared__ FilterStage st;
#pragma unroll
    for (int r = 0; r < RPX_FILTER_R; r++)
        if (sel[r]) st.src[(uint32_t)(slots.pos[r] - slots.begin)] = (uint16_t)(r * RPX_TILE + threadIdx.x);
    __syncthreads();
    const unsigned long long tile0 = (unsigned long long)tile * RPX_FILTER_TILE;
    for (uint32_t sl = threadIdx.x; sl < total; sl += RPX_TILE) {
        const unsigned long long i = tile0 + st.src[sl];
        const unsigned long long pos = slots.begin + sl;
        if (pos >= ocap) break;  // capacity overrun: reported by the host from *d_next, never written
#pragma unroll
        for (int fld = 0; fld < NF; fld++) out.f[(unsigned long long)fld * ocap + pos] = in.f[(unsigned long long)fld * cap + i];
#pragma unroll
        for (int fld = 0; fld < NU; fld++) {
            uint32_t v = in.u[(unsigned long long)fld * cap + i];
            if (fld == U_PARENT) v += parent_offset;
            out.u[(unsigned long long)fld * ocap + pos] = v;
        }
        if (GAUSS) {
#pragma unroll 12
            for (int fld = 0; fld < NP; fld++) out.p[(unsigned long long)fld * ocap + pos] = in.p[(unsigned long long)fld * cap + i];
        }
    }
}

// Append collection `in` to `out` behind the `off` records it already holds (row-wise copy, coalesced).
__global__ void k_append(Soa in, Soa out, unsigned long long off, uint32_t parent_offset) {
    const unsigned long long n = in.n, cap = in.cap, ocap = out.cap;
    const int rows = NF + NU + (in.p ? NP : 0);
    for (int row = blockIdx.y; row < rows; row += gridDim.y) {
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (unsigned long long)gridDim.x * blockDim.x) {
            if (row < NF) {
                out.f[(unsigned long long)row * ocap + off + i] = in.f[(unsigned long long)row * cap + i];
            } else if (row < NF + NU) {
                const int r = row - NF;
                uint32_t v = in.u[(unsigned long long)r * cap + i];
                if (r == U_PARENT) v += parent_offset;
                out.u[(unsigned long long)r * ocap + off + i] = v;
            } else {
                const int r = row - NF - NU;
                out.p[(unsigned long long)r * ocap + off + i] = in.p[(unsigned long long)r * cap + i];
            }
        }
    }
}

void launch_aos_to_soa(cudaStream_t st, const void* d_aos, const rpx_rays* r) {
    const uint64_t n = r->soa.n;
    if (!n) return;
    if (r->is_gausslet) {
        const int T = 64;
        k_aos_to_soa<RPX_WORDS_GAUSSLET, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, st>>>((const uint32_t*)d_aos, r->soa);
    } else {
        const int T = 256;
        k_aos_to_soa<RPX_WORDS_RAY, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_RAY_BYTES, st>>>((const uint32_t*)d_aos, r->soa);
    }
}

void launch_soa_to_aos(cudaStream_t st, const rpx_rays* r, void* d_aos) {
    const uint64_t n = r->soa.n;
    if (!n) return;
    if (r->is_gausslet) {
        const int T = 64;
        k_soa_to_aos<RPX_WORDS_GAUSSLET, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, st>>>(r->soa, (uint32_t*)d_aos, 0u);
    } else {
        const int T = 256;
        k_soa_to_aos<RPX_WORDS_RAY, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_RAY_BYTES, st>>>(r->soa, (uint32_t*)d_aos, 0u);
    }
}

// Selection of terminal rays from a list of device collections into `dst` behind the *d_total records
// already there.  d_total is a device counter chained through the launches (like rpx_capture's totals).
// Enqueues only; the caller synchronises and reads the counters back.  scratch must hold
// tiles_total u64 + n_gens u32 (zeroed here).  Returns kernels launched or a negative status.
struct SelectPlan {
    size_t state_bytes, cnt_bytes, total_bytes;
};
SelectPlan select_plan(const rpx_rays* const* gens, int n_gens) {
    unsigned long long tiles = 0;
    for (int j = 0; j < n_gens; j++)
        if (gens[j]) tiles += rpx_state_words((gens[j]->soa.n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);
    SelectPlan p;
    p.state_bytes = sizeof(unsigned long long) * (size_t)(tiles ? tiles : 1);
    p.cnt_bytes = sizeof(uint32_t) * (size_t)(n_gens + 2);
    p.total_bytes = sizeof(unsigned long long) * (size_t)(n_gens + 1);
    return p;
}

int enqueue_select(rpx_ctx* ctx, const rpx_rays* const* gens, int n_gens, int unterminated, const unsigned char* d_face_select,
                   const uint32_t* parent_offsets, rpx_rays* dst, int copy, unsigned char* scratch, const SelectPlan& plan,
                   unsigned long long first_total_from /* device ptr or null */, unsigned long long** d_totals_out) {
    // scratch layout: [totals (n_gens + 1) u64][tile state][ticket counters]
    (void)first_total_from;
    cudaStream_t st = ctx->stream;
    unsigned long long* d_totals = (unsigned long long*)scratch;
    unsigned long long* d_state = (unsigned long long*)(scratch + plan.total_bytes);
    uint32_t* d_cnt = (uint32_t*)(scratch + plan.total_bytes + plan.state_bytes);
    *d_totals_out = d_totals;
    int launches = 0;
    unsigned long long tile_off = 0;
    for (int j = 0; j < n_gens; j++) {
        const unsigned long long n = gens[j] ? gens[j]->soa.n : 0;
        if (n == 0) {
            if (cudaMemcpyAsync(d_totals + j + 1, d_totals + j, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                return RPX_ERR_CUDA;
            continue;
        }
        const unsigned n_tiles = (unsigned)((n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);
        const uint32_t poff = parent_offsets ? parent_offsets[j] : 0u;
        if (gens[j]->is_gausslet)
            k_select<true><<<n_tiles, RPX_TILE, 0, st>>>(gens[j]->soa, dst->soa, d_state + tile_off, d_cnt + j, d_totals + j,
                                                         d_totals + j + 1, unterminated, d_face_select, poff, copy);
        else
            k_select<false><<<n_tiles, RPX_TILE, 0, st>>>(gens[j]->soa, dst->soa, d_state + tile_off, d_cnt + j, d_totals + j,
                                                          d_totals + j + 1, unterminated, d_face_select, poff, copy);
        if (cudaGetLastError() != cudaSuccess) return RPX_ERR_CUDA;
        tile_off += rpx_state_words(n_tiles);
        launches++;
    }
    return launches;
}

}  // namespace

// ------------------------------------------------------------------ rpx_select_terminal
extern "C" int rpx_select_terminal(rpx_ctx* ctx, const rpx_rays* const* gens, int n_gens, int select_unterminated,
                                   const uint8_t* face_select, rpx_rays** out, uint64_t* counts) {
    if (!ctx || !gens || !out || n_gens <= 0) return fail(ctx, RPX_ERR_INVALID, "NULL / empty argument");
    *out = nullptr;
    if (face_select && !ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before selecting by face");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long total = 0;
    int is_g = -1;
    for (int j = 0; j < n_gens; j++) {
        if (!gens[j]) return fail(ctx, RPX_ERR_INVALID, "collection %d is NULL (dropped generation?)", j);
        if (is_g < 0) is_g = gens[j]->is_gausslet;
        if (gens[j]->is_gausslet != is_g) return fail(ctx, RPX_ERR_INVALID, "collections mix rays and gausslets");
        total += gens[j]->soa.n;
    }
    cudaStream_t st = ctx->stream;
    const SelectPlan plan = select_plan(gens, n_gens);
    const size_t nsel = face_select ? (size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1) : 0;
    const size_t scratch_bytes = plan.total_bytes + plan.state_bytes + plan.cnt_bytes + 8;
    unsigned char* scratch = nullptr;
    unsigned char* d_sel = nullptr;
    cudaError_t e;
    auto cleanup = [&]() {
        if (scratch) cudaFreeAsync(scratch, st);
        if (d_sel) cudaFreeAsync(d_sel, st);
    };
    if ((e = cudaMallocAsync((void**)&scratch, scratch_bytes, st)) != cudaSuccess ||
        (e = cudaMemsetAsync(scratch, 0, scratch_bytes, st)) != cudaSuccess ||
        (nsel && ((e = cudaMallocAsync((void**)&d_sel, nsel, st)) != cudaSuccess ||
                  (e = cudaMemcpyAsync(d_sel, face_select, (size_t)ctx->n_traced, cudaMemcpyHostToDevice, st)) != cudaSuccess))) {
        cleanup();
        return fail(ctx, RPX_ERR_NOMEM, "terminal selection scratch: %s", cudaGetErrorString(e));
    }
    // pass 1 counts, pass 2 copies into an exactly sized collection (terminal rays are a small part of a
    // trace: sizing the output like the input would double the footprint of the largest generation)
    std::vector<unsigned long long> h_totals((size_t)n_gens + 1, 0);
    unsigned long long* d_totals = nullptr;
    rpx_rays dummy;
    dummy.soa = gens[0]->soa;
    int rc = enqueue_select(ctx, gens, n_gens, select_unterminated, d_sel, nullptr, &dummy, 0, scratch, plan, 0, &d_totals);
    if (rc < 0 || (e = cudaMemcpyAsync(h_totals.data(), d_totals, plan.total_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
        cleanup();
        return fail(ctx, RPX_ERR_CUDA, "terminal count pass failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    const unsigned long long n_sel = h_totals[(size_t)n_gens];
    rpx_rays* dst = nullptr;
    rc = rpx_rays_alloc(ctx, n_sel, is_g, &dst);
    if (rc != RPX_OK) {
        cleanup();
        return rc;
    }
    if (n_sel) {
        if ((e = cudaMemsetAsync(scratch, 0, scratch_bytes, st)) != cudaSuccess ||
            enqueue_select(ctx, gens, n_gens, select_unterminated, d_sel, nullptr, dst, 1, scratch, plan, 0, &d_totals) < 0 ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess) {
            cleanup();
            rpx_rays_free(ctx, dst);
            return fail(ctx, RPX_ERR_CUDA, "terminal copy pass failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    cleanup();
    dst->soa.n = n_sel;
    if (counts)
        for (int j = 0; j < n_gens; j++) counts[j] = h_totals[(size_t)j + 1] - h_totals[(size_t)j];
    *out = dst;
    (void)total;
    return RPX_OK;
}

// ------------------------------------------------------------------ device-resident AoS export / import
extern "C" int rpx_rays_export_device(rpx_ctx* ctx, const rpx_rays* rays, void* d_aos, uint64_t capacity) {
    if (!ctx || !rays || (!d_aos && rays->soa.n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (capacity < rays->soa.n)
        return fail(ctx, RPX_ERR_INVALID, "output holds %llu records, collection has %llu", (unsigned long long)capacity,
                    (unsigned long long)rays->soa.n);
    if (((uintptr_t)d_aos & 3u) != 0) return fail(ctx, RPX_ERR_INVALID, "device buffer must be 4-byte aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    launch_soa_to_aos(ctx->stream, rays, d_aos);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return RPX_OK;
}

extern "C" int rpx_rays_import_device(rpx_ctx* ctx, const void* d_aos, uint64_t n, int is_gausslet, rpx_rays** out_rays) {
    if (!ctx || !out_rays || (!d_aos && n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    *out_rays = nullptr;
    if (((uintptr_t)d_aos & 3u) != 0) return fail(ctx, RPX_ERR_INVALID, "device buffer must be 4-byte aligned");
    if (n >= 0xFFFFFFFFull) return fail(ctx, RPX_ERR_INVALID, "%llu rays exceed the 32-bit parent_idx of ray_t", (unsigned long long)n);
    CU(ctx, cudaSetDevice(ctx->device));
    rpx_rays* r = nullptr;
    int rc = rpx_rays_alloc(ctx, n, is_gausslet, &r);
    if (rc != RPX_OK) return rc;
    r->soa.n = n;
    launch_aos_to_soa(ctx->stream, d_aos, r);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        rpx_rays_free(ctx, r);
        return fail(ctx, RPX_ERR_CUDA, "import failed: %s", cudaGetErrorString(e));
    }
    *out_rays = r;
    return RPX_OK;
}

// ------------------------------------------------------------------ rpx_trace_consume
extern "C" int rpx_trace_consume(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                                 int recursion_limit, const rpx_consume_opts* opts, uint32_t* face_counts,
                                 rpx_consume_result* result) {
    if (!ctx || !opts || !result || (!rays_aos && n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    memset(result, 0, sizeof *result);
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    const uint32_t flags = opts->flags;
    const bool on_device = (flags & RPX_CONSUME_SOURCE_ON_DEVICE) != 0;
    const bool want_term = (flags & RPX_CONSUME_TERMINAL) != 0 || opts->terminal_faces != nullptr;
    const bool want_cap = (flags & RPX_CONSUME_CAPTURE) != 0;
    const bool want_field = (flags & RPX_CONSUME_FIELD) != 0;
    if (want_cap && !ctx->have_capture) return fail(ctx, RPX_ERR_STATE, "RPX_CONSUME_CAPTURE needs rpx_capture_scene_set");
    if (want_field && (!want_cap || !opts->detector || !is_gausslet))
        return fail(ctx, RPX_ERR_INVALID, "RPX_CONSUME_FIELD needs gausslets, RPX_CONSUME_CAPTURE and a detector");
    if ((opts->per_chunk_terminal || opts->per_chunk_captured) && (opts->max_gens <= 0 || opts->max_gens > RPX_CONSUME_MAX_GENS))
        return fail(ctx, RPX_ERR_INVALID, "max_gens must be 1..%d when per-chunk counts are requested", RPX_CONSUME_MAX_GENS);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!on_device && !ctx->have_copy_streams) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
        CU(ctx, cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) CU(ctx, cudaEventCreateWithFlags(&ctx->st_out_done[k], cudaEventDisableTiming));
        ctx->have_copy_streams = true;
    }
    const size_t rec = is_gausslet ? RPX_GAUSSLET_BYTES : RPX_RAY_BYTES;
    uint64_t chunk_rays = opts->chunk_rays ? opts->chunk_rays : (is_gausslet ? (1ull << 20) : (1ull << 22));
    const uint64_t n_chunks = n ? (n + chunk_rays - 1) / chunk_rays : 0;
    if (n_chunks > 0x7fffffffull) return fail(ctx, RPX_ERR_INVALID, "too many chunks");
    result->n_chunks = (int32_t)n_chunks;
    std::vector<uint64_t> totals((size_t)RPX_CONSUME_MAX_GENS, 0);
    std::vector<uint32_t> fc((size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1), 0);
    int rc = RPX_OK;
    cudaError_t e = cudaSuccess;

    // kept collections + the device counters that chain the appends
    rpx_rays* kept_term = nullptr;
    rpx_rays* kept_cap = nullptr;
    unsigned char* d_sel = nullptr;
    if (want_term && opts->terminal_capacity) rc = rpx_rays_alloc(ctx, opts->terminal_capacity, is_gausslet, &kept_term);
    if (rc == RPX_OK && want_cap && opts->captured_capacity) rc = rpx_rays_alloc(ctx, opts->captured_capacity, is_gausslet, &kept_cap);
    if (rc == RPX_OK && opts->terminal_faces) {
        const size_t nsel = (size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1);
        if ((e = cudaMallocAsync((void**)&d_sel, nsel, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_sel, opts->terminal_faces, (size_t)ctx->n_traced, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            rc = fail(ctx, RPX_ERR_NOMEM, "terminal face table: %s", cudaGetErrorString(e));
    }
    uint64_t n_term = 0, n_cap = 0;

    // host source: double-buffered upload ring on stream_in (as rpx_trace_streamed)
    void* d_in[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    bool used_once[2] = {false, false};
    if (!on_device && n_chunks && rc == RPX_OK) {
        const uint64_t in_bytes = (n < chunk_rays ? n : chunk_rays) * rec;
        for (int k = 0; k < (n_chunks > 1 ? 2 : 1) && e == cudaSuccess; k++) {
            if (ctx->st_in_bytes[k] < in_bytes) {
                if (ctx->st_in[k]) cudaFree(ctx->st_in[k]);
                ctx->st_in[k] = nullptr;
                ctx->st_in_bytes[k] = 0;
                if ((e = cudaMalloc(&ctx->st_in[k], in_bytes)) == cudaSuccess) ctx->st_in_bytes[k] = in_bytes;
            }
            d_in[k] = ctx->st_in[k];
        }
        if (e != cudaSuccess) rc = fail(ctx, RPX_ERR_NOMEM, "chunk staging (%llu bytes): %s", (unsigned long long)in_bytes, cudaGetErrorString(e));
        for (int k = 0; k < 2; k++) {
            cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev_used[k], cudaEventDisableTiming);
        }
    }
    auto issue_upload = [&](uint64_t c) -> cudaError_t {
        const int b = (int)(c & 1);
        const uint64_t lo = c * chunk_rays, cnt = (n - lo < chunk_rays) ? n - lo : chunk_rays;
        cudaError_t err = cudaSuccess;
        if (used_once[b]) err = cudaStreamWaitEvent(ctx->stream_in, ev_used[b], 0);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(d_in[b], (const unsigned char*)rays_aos + lo * rec, cnt * rec, cudaMemcpyHostToDevice, ctx->stream_in);
        if (err == cudaSuccess) err = cudaEventRecord(ev_in[b], ctx->stream_in);
        return err;
    };
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    cudaEventCreate(&ev_begin);
    cudaEventCreate(&ev_end);
    cudaEventRecord(ev_begin, st);
    if (rc == RPX_OK && !on_device && n_chunks && (e = issue_upload(0)) != cudaSuccess)
        rc = fail(ctx, RPX_ERR_CUDA, "chunk upload: %s", cudaGetErrorString(e));

    for (uint64_t c = 0; c < n_chunks && rc == RPX_OK; c++) {
        const int b = (int)(c & 1);
        const uint64_t lo = c * chunk_rays, cnt = (n - lo < chunk_rays) ? n - lo : chunk_rays;
        if (!on_device && c + 1 < n_chunks && (e = issue_upload(c + 1)) != cudaSuccess) {
            rc = fail(ctx, RPX_ERR_CUDA, "chunk upload: %s", cudaGetErrorString(e));
            break;
        }
        rpx_rays* r = nullptr;
        if ((rc = rpx_rays_alloc(ctx, cnt, is_gausslet, &r)) != RPX_OK) break;
        r->soa.n = cnt;
        const void* src = on_device ? (const void*)((const unsigned char*)rays_aos + lo * rec) : d_in[b];
        if (!on_device) cudaStreamWaitEvent(st, ev_in[b], 0);
        launch_aos_to_soa(st, src, r);
        result->launches++;
        if (!on_device) {
            cudaEventRecord(ev_used[b], st);
            used_once[b] = true;
        }
        rpx_result* res = nullptr;
        rc = rpx_trace_device(ctx, r, max_length, recursion_limit, RPX_TRACE_DEFAULT, &res);  // owns r
        if (rc != RPX_OK) break;
        result->trace_ms += res->device_ms;
        result->launches += res->launches;
        result->intersect_ms += res->k_ms[0];
        result->shade_ms += res->k_ms[1];
        result->intersect_launches += res->k_launches[0];
        result->shade_launches += res->k_launches[1];
        for (size_t i = 0; i < res->face_counts.size() && i < fc.size(); i++) fc[i] += res->face_counts[i];
        const int ng = (int)res->gens.size();
        if (ng > RPX_CONSUME_MAX_GENS) {
            rc = fail(ctx, RPX_ERR_INVALID, "trace produced %d generations (limit %d)", ng, RPX_CONSUME_MAX_GENS);
            rpx_result_free(ctx, res);
            break;
        }
        // global parent numbering of this chunk's generation g: rays of generation g-1 in earlier chunks
        std::vector<uint32_t> poff((size_t)(ng > 0 ? ng : 1), 0u);
        for (int g = 0; g < ng && rc == RPX_OK; g++) {
            const uint64_t m = res->gens[(size_t)g] ? res->gens[(size_t)g]->soa.n : 0;
            if (totals[(size_t)g] + m >= 0xFFFFFFFFull)
                rc = fail(ctx, RPX_ERR_INVALID, "generation %d would exceed the 32-bit parent_idx of ray_t (%llu rays)", g,
                          (unsigned long long)(totals[(size_t)g] + m));
            poff[(size_t)g] = g > 0 ? (uint32_t)totals[(size_t)g - 1] : 0u;
        }
        // ---- consumer 1: terminal rays
        if (rc == RPX_OK && want_term && ng > 0) {
            const SelectPlan plan = select_plan(res->gens.data(), ng);
            const size_t scratch_bytes = plan.total_bytes + plan.state_bytes + plan.cnt_bytes + 8;
            unsigned char* scratch = nullptr;
            std::vector<unsigned long long> h_tot((size_t)ng + 1, 0);
            if ((e = cudaMallocAsync((void**)&scratch, scratch_bytes, st)) != cudaSuccess ||
                (e = cudaMemsetAsync(scratch, 0, scratch_bytes, st)) != cudaSuccess) {
                rc = fail(ctx, RPX_ERR_NOMEM, "terminal selection scratch: %s", cudaGetErrorString(e));
            } else {
                // the chain starts at the records already kept
                unsigned long long start = kept_term ? n_term : 0;
                cudaMemcpyAsync(scratch, &start, sizeof start, cudaMemcpyHostToDevice, st);
                unsigned long long* d_totals = nullptr;
                rpx_rays dummy;
                dummy.soa = res->gens[0]->soa;
                const int l = enqueue_select(ctx, res->gens.data(), ng, (flags & RPX_CONSUME_TERMINAL) ? 1 : 0, d_sel, poff.data(),
                                             kept_term ? kept_term : &dummy, kept_term ? 1 : 0, scratch, plan, 0, &d_totals);
                if (l < 0 || (e = cudaMemcpyAsync(h_tot.data(), d_totals, plan.total_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                    (e = cudaStreamSynchronize(st)) != cudaSuccess)
                    rc = fail(ctx, RPX_ERR_CUDA, "terminal selection failed: %s", cudaGetErrorString(cudaGetLastError()));
                else
                    result->launches += (uint64_t)l;
            }
            if (scratch) cudaFreeAsync(scratch, st);
            if (rc == RPX_OK) {
                const unsigned long long got = h_tot[(size_t)ng] - h_tot[0];
                if (kept_term && n_term + got > opts->terminal_capacity)
                    rc = fail(ctx, RPX_ERR_NOMEM, "terminal_capacity %llu is too small (chunk %llu needs %llu)",
                              (unsigned long long)opts->terminal_capacity, (unsigned long long)c, (unsigned long long)(n_term + got));
                if (opts->per_chunk_terminal)
                    for (int g = 0; g < ng && g < opts->max_gens; g++)
                        opts->per_chunk_terminal[c * (uint64_t)opts->max_gens + (uint64_t)g] = h_tot[(size_t)g + 1] - h_tot[(size_t)g];
                n_term += got;
            }
        }
        // ---- consumer 2: capture plane (+ detector field)
        if (rc == RPX_OK && want_cap && ng > 0) {
            rpx_rays* cap = nullptr;
            std::vector<uint64_t> ccounts((size_t)ng, 0);
            rc = rpx_capture(ctx, res->gens.data(), ng, nullptr, nullptr, 0, &cap, ccounts.data());
            if (rc == RPX_OK) {
                result->launches += (uint64_t)ng;
                const uint64_t got = cap->soa.n;
                if (opts->per_chunk_captured)
                    for (int g = 0; g < ng && g < opts->max_gens; g++)
                        opts->per_chunk_captured[c * (uint64_t)opts->max_gens + (uint64_t)g] = ccounts[(size_t)g];
                if (want_field && got) {
                    rc = rpx_detector_accumulate(ctx, opts->detector, cap);
                    result->launches += 2;
                }
                if (rc == RPX_OK && kept_cap && got) {
                    if (n_cap + got > opts->captured_capacity) {
                        rc = fail(ctx, RPX_ERR_NOMEM, "captured_capacity %llu is too small (chunk %llu needs %llu)",
                                  (unsigned long long)opts->captured_capacity, (unsigned long long)c, (unsigned long long)(n_cap + got));
                    } else {
                        // captured records keep the parent_idx of their generation: renumber per piece
                        unsigned long long off_in = 0;
                        for (int g = 0; g < ng; g++) {
                            const uint64_t m = ccounts[(size_t)g];
                            if (!m) continue;
                            Soa piece = cap->soa;
                            piece.f += off_in;
                            piece.u += off_in;
                            if (piece.p) piece.p += off_in;
                            piece.n = m;
                            dim3 grid((unsigned)((m + 255) / 256 < 1024 ? (m + 255) / 256 : 1024), (unsigned)(NF + NU + (is_gausslet ? NP : 0)));
                            k_append<<<grid, 256, 0, st>>>(piece, kept_cap->soa, n_cap + off_in, poff[(size_t)g]);
                            result->launches++;
                            off_in += m;
                        }
                        if ((e = cudaGetLastError()) != cudaSuccess) rc = fail(ctx, RPX_ERR_CUDA, "k_append: %s", cudaGetErrorString(e));
                    }
                }
                n_cap += got;
                rpx_rays_free(ctx, cap);
            }
        }
        if (rc == RPX_OK) {
            for (int g = 0; g < ng; g++) totals[(size_t)g] += res->gens[(size_t)g] ? res->gens[(size_t)g]->soa.n : 0;
            if (ng > result->n_gens) result->n_gens = ng;
        }
        rpx_result_free(ctx, res);  // generation buffers go back to the pool in stream order
    }
    cudaEventRecord(ev_end, st);
    if (!on_device && ctx->have_copy_streams) cudaStreamSynchronize(ctx->stream_in);
    cudaStreamSynchronize(st);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_begin, ev_end) == cudaSuccess) result->device_ms = ms;
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
    for (int k = 0; k < 2; k++) {
        if (ev_in[k]) cudaEventDestroy(ev_in[k]);
        if (ev_used[k]) cudaEventDestroy(ev_used[k]);
    }
    if (d_sel) cudaFreeAsync(d_sel, st);
    if (rc != RPX_OK) {
        if (kept_term) rpx_rays_free(ctx, kept_term);
        if (kept_cap) rpx_rays_free(ctx, kept_cap);
        return rc;
    }
    for (int g = 0; g < result->n_gens; g++) result->counts[g] = totals[(size_t)g];
    result->n_terminal = n_term;
    result->n_captured = n_cap;
    if (kept_term) kept_term->soa.n = n_term;
    if (kept_cap) kept_cap->soa.n = n_cap;
    result->terminal = kept_term;
    result->captured = kept_cap;
    if (face_counts)
        for (int i = 0; i < ctx->n_traced; i++) face_counts[i] = fc[(size_t)i];
    return RPX_OK;
}
