// rpx_api.cu -- the C ABI of librpx (include/rpx.h): context, scene upload, ray
// upload/download (AoS <-> SoA on the device), the generation loop and result handles.
//
// Host runtime design (B200): one context per process per GPU; one non-blocking CUDA
// stream; generation buffers come from the stream-ordered memory pool (cudaMallocAsync
// with an unlimited release threshold, so after the first trace no allocation ever
// reaches the driver); every kernel launch is bracketed by CUDA events on that stream so
// bench.py can report per-kernel device time without a profiler attached.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <cstddef>
#include <cstring>
#include <vector>

#include "../../include/rpx.h"
#include "rpx_internal.h"
#include "rpx_launch.h"

using namespace rpx;

static_assert(sizeof(rpx_ray) == 188, "ray_t layout");
static_assert(sizeof(rpx_para) == 80, "para_t layout");
static_assert(sizeof(rpx_gausslet) == 668, "gausslet_t layout");
static_assert(sizeof(rpx_face) == 176, "rpx_face layout");
static_assert(sizeof(rpx_material) == 112, "rpx_material layout");
static_assert(sizeof(rpx_face_set) == 200, "rpx_face_set layout");
static_assert(sizeof(rpx_shape_op) == 48, "rpx_shape_op layout");
static_assert(sizeof(rpx_implicit_op) == 64, "rpx_implicit_op layout");
static_assert(sizeof(rpx_ztape_op) == 32, "rpx_ztape_op layout");
static_assert(sizeof(rpx_zcoef) == 40, "rpx_zcoef layout");
static_assert(sizeof(rpx_distortion) == 64, "rpx_distortion layout");

static thread_local std::string g_init_error;

int rpx_fail(rpx_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_init_error = buf;
    return code;
}

extern "C" int rpx_abi_version(void) { return RPX_ABI_VERSION; }

extern "C" const char* rpx_last_error(const rpx_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int rpx_init(int device, rpx_ctx** out_ctx) {
    if (!out_ctx) return fail(nullptr, RPX_ERR_INVALID, "out_ctx is NULL");
    *out_ctx = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RPX_ERR_NODEVICE,
                    "no CUDA device available (%s); librpx has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev)
        return fail(nullptr, RPX_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    rpx_ctx* ctx = new (std::nothrow) rpx_ctx();
    if (!ctx) return fail(nullptr, RPX_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->have_scene = false;
    ctx->scene_block = nullptr;
    ctx->have_copy_streams = false;
    for (int k = 0; k < 2; k++) { ctx->st_in[k] = nullptr; ctx->st_in_bytes[k] = 0; }
    for (int k = 0; k < 4; k++) { ctx->st_out[k] = nullptr; ctx->st_out_bytes[k] = 0; ctx->st_out_busy[k] = false; }
    ctx->have_capture = false;
    ctx->cap_block = nullptr;
    ctx->cap_face_ids = nullptr;
    ctx->tile_state = nullptr;
    ctx->tile_state_cap = 0;
    ctx->ev_used = 0;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        fail(nullptr, RPX_ERR_CUDA, "cannot open device %d: %s", device, cudaGetErrorString(e));
        delete ctx;
        return RPX_ERR_CUDA;
    }
    // keep freed generation buffers in the pool: no driver allocation after warm-up
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if ((e = cudaMalloc(&ctx->tile_counter, sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_count, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMallocHost(&ctx->h_count, sizeof(unsigned long long))) != cudaSuccess) {
        fail(nullptr, RPX_ERR_CUDA, "context scratch allocation failed: %s", cudaGetErrorString(e));
        delete ctx;
        return RPX_ERR_CUDA;
    }
    ctx->d_face_counts = nullptr;
    if ((e = cudaMalloc(&ctx->d_counts, sizeof(unsigned long long) * RPX_MAX_PIPE_GENS)) != cudaSuccess ||
        (e = cudaHostAlloc(&ctx->h_counts, sizeof(unsigned long long) * RPX_MAX_PIPE_GENS, cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(&ctx->h_counts_dev, ctx->h_counts, 0)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->pipe_counters, sizeof(uint32_t) * RPX_MAX_PIPE_GENS)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->pipe_hits, sizeof(uint32_t) * RPX_MAX_PIPE_GENS)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->pipe_miss, sizeof(uint32_t) * RPX_MAX_PIPE_GENS)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->pipe_state, sizeof(unsigned long long) * RPX_PIPE_STATE_TILES)) != cudaSuccess) {
        fail(nullptr, RPX_ERR_CUDA, "context scratch allocation failed: %s", cudaGetErrorString(e));
        delete ctx;
        return RPX_ERR_CUDA;
    }
    ctx->pipe_state_cap = RPX_PIPE_STATE_TILES;
    *out_ctx = ctx;
    return RPX_OK;
}

extern "C" void rpx_shutdown(rpx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t ev : ctx->ev_pool) cudaEventDestroy(ev);
    if (ctx->scene_block) cudaFree(ctx->scene_block);
    if (ctx->cap_block) cudaFree(ctx->cap_block);
    if (ctx->cap_face_ids) cudaFree(ctx->cap_face_ids);
    if (ctx->tile_state) cudaFree(ctx->tile_state);
    if (ctx->d_face_counts) cudaFree(ctx->d_face_counts);
    cudaFree(ctx->tile_counter);
    cudaFree(ctx->d_counts);
    cudaFree(ctx->pipe_counters);
    cudaFree(ctx->pipe_hits);
    cudaFree(ctx->pipe_miss);
    cudaFree(ctx->pipe_state);
    cudaFreeHost(ctx->h_counts);
    cudaFree(ctx->d_count);
    cudaFreeHost(ctx->h_count);
    if (ctx->have_copy_streams) {
        for (int k = 0; k < 2; k++) if (ctx->st_in[k]) cudaFree(ctx->st_in[k]);
        for (int k = 0; k < 4; k++) {
            if (ctx->st_out[k]) cudaFree(ctx->st_out[k]);
            cudaEventDestroy(ctx->st_out_done[k]);
        }
        cudaStreamDestroy(ctx->stream_in);
        cudaStreamDestroy(ctx->stream_out);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void* rpx_stream(rpx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" void* rpx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void rpx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ scene
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int validate_scene(rpx_ctx* ctx, const rpx_scene* s) {
    if (s->abi_version != RPX_ABI_VERSION)
        return fail(ctx, RPX_ERR_INVALID, "scene abi_version %d != %d", s->abi_version, RPX_ABI_VERSION);
    if (s->n_traced_faces < 0 || s->n_faces < s->n_traced_faces || s->n_face_sets < 0)
        return fail(ctx, RPX_ERR_INVALID, "bad face counts");
    for (int i = 0; i < s->n_faces; i++) {
        const rpx_face& f = s->faces[i];
        if (f.type < RPX_FACE_CIRCULAR || f.type > RPX_FACE_UVPATCH)
            return fail(ctx, RPX_ERR_UNSUPPORTED, "face %d: unsupported face type %d", i, f.type);
        if (f.face_set < 0 || f.face_set >= s->n_face_sets)
            return fail(ctx, RPX_ERR_INVALID, "face %d: face_set %d out of range", i, f.face_set);
        if (i < s->n_traced_faces && (f.material < 0 || f.material >= s->n_materials))
            return fail(ctx, RPX_ERR_INVALID, "face %d: material %d out of range", i, f.material);
        if (f.shape_off >= 0 && f.shape_off + f.shape_len > s->n_shape_ops)
            return fail(ctx, RPX_ERR_INVALID, "face %d: shape program out of range", i);
        if (f.shape_len > 32) return fail(ctx, RPX_ERR_UNSUPPORTED, "face %d: shape tree too deep", i);
        if (f.type == RPX_FACE_DISTORTION) {
            if (f.base_face < 0 || f.base_face >= s->n_faces || s->faces[f.base_face].type == RPX_FACE_DISTORTION)
                return fail(ctx, RPX_ERR_INVALID, "face %d: bad base_face", i);
            if (f.aux_off < 0 || f.aux_off >= s->n_distortions)
                return fail(ctx, RPX_ERR_INVALID, "face %d: distortion index out of range", i);
        }
        if (f.type == RPX_FACE_IMPLICIT_PLANAR && (f.aux_off < 0 || f.aux_off + f.aux_n > s->n_implicit_ops))
            return fail(ctx, RPX_ERR_INVALID, "face %d: implicit program out of range", i);
        if ((f.type == RPX_FACE_POLYGON || f.type == RPX_FACE_ORIENTED_POLYGON) &&
            (f.aux_off < 0 || f.aux_off + 2 * f.aux_n > s->n_pool || f.aux_n < 1))
            return fail(ctx, RPX_ERR_INVALID, "face %d: polygon points out of range", i);
        if (f.type == RPX_FACE_EXT_POLY && (f.aux_off < 0 || f.aux_off + f.aux_n * f.aux_m > s->n_pool))
            return fail(ctx, RPX_ERR_INVALID, "face %d: coefficient table out of range", i);
        if (f.type == RPX_FACE_MESH || f.type == RPX_FACE_UVPATCH) {
            if (f.aux_off < 0 || f.aux_n < 1 || f.aux_m < 1 || f.aux_off + 8 > s->n_pool)
                return fail(ctx, RPX_ERR_INVALID, "face %d: mesh block out of range", i);
            const double* H = s->pool + f.aux_off;
            const long long n_pts = (long long)H[0], n_cells = (long long)H[1], n_nodes = (long long)H[2];
            const long long end = f.aux_off + 8 + 3 * n_pts + 3 * n_cells + 16 * n_cells + 8 * n_nodes;
            if (n_cells != f.aux_n || n_nodes != f.aux_m || n_pts < 3 || end > s->n_pool || (long long)H[3] != 8 ||
                (long long)H[4] != 8 + 3 * n_pts || (long long)H[5] != 8 + 3 * n_pts + 3 * n_cells ||
                (long long)H[6] != 8 + 3 * n_pts + 19 * n_cells)
                return fail(ctx, RPX_ERR_INVALID, "face %d: inconsistent mesh block header", i);
            // the device traversal trusts the node links: check them here, once
            const double* cells = H + (long long)H[4];
            for (long long c = 0; c < 3 * n_cells; c++)
                if (!(cells[c] >= 0 && cells[c] < (double)n_pts))
                    return fail(ctx, RPX_ERR_INVALID, "face %d: mesh cell refers to a missing point", i);
            const double* nodes = H + (long long)H[6];
            for (long long k = 0; k < n_nodes; k++) {
                const double a = nodes[8 * k + 6], b = nodes[8 * k + 7];
                const bool ok = a >= 0 ? (a > (double)k && a < (double)n_nodes && b > (double)k && b < (double)n_nodes)
                                       : (-a - 1 >= 0 && b >= 1 && (-a - 1) + b <= (double)n_cells);
                if (!ok) return fail(ctx, RPX_ERR_INVALID, "face %d: bad BVH node %lld", i, k);
            }
            // the device walk keeps an explicit stack of 64 entries (one per level + 1): children have
            // larger ids than their parent (checked above), so one backward pass gives every subtree height
            {
                std::vector<int> height((size_t)n_nodes, 1);
                for (long long k = n_nodes - 1; k >= 0; k--) {
                    const double a = nodes[8 * k + 6], b = nodes[8 * k + 7];
                    if (a >= 0) height[(size_t)k] = 1 + (height[(size_t)a] > height[(size_t)b] ? height[(size_t)a] : height[(size_t)b]);
                }
                if (height[0] > 60)
                    return fail(ctx, RPX_ERR_UNSUPPORTED, "face %d: BVH depth %d exceeds the device traversal stack (60)", i, height[0]);
            }
        }
        if (f.type == RPX_FACE_UVPATCH) {
            const long long N = (long long)f.p[3], M = (long long)f.p[4], off = (long long)f.p[7];
            const long long n_pts = (long long)s->pool[f.aux_off];
            const long long na = (long long)f.p[8], nb = (long long)f.p[9];
            const int kind = (int)f.p[2];
            if (N < 0 || M < 0 || N > 64 || M > 64 || off < 0 || na < 0 || nb < 0 || (kind != 0 && kind != 1) ||
                off + 2 * n_pts + 3 * (N + 1) * (M + 1) + na + nb > s->n_pool)
                return fail(ctx, RPX_ERR_INVALID, "face %d: patch block out of range", i);
            if (kind == 0 ? (na != N + 1 || nb != M + 1)
                          : (f.p[5] < 0 || f.p[6] < 0 || f.p[5] > 8 || f.p[6] > 8 || na < N + (long long)f.p[5] + 2 ||
                             nb < M + (long long)f.p[6] + 2))
                return fail(ctx, RPX_ERR_INVALID, "face %d: patch tables inconsistent with its orders / degrees", i);
        }
        if (f.type == RPX_FACE_EXTRUDED_BEZIER && (f.aux_off < 0 || f.aux_n < 1 || f.aux_off + 8 * f.aux_n > s->n_pool))
            return fail(ctx, RPX_ERR_INVALID, "face %d: Bezier control points out of range", i);
        if (f.type == RPX_FACE_ELLIPSOIDAL && (f.aux_off < 0 || f.aux_off + 24 > s->n_pool))
            return fail(ctx, RPX_ERR_INVALID, "face %d: transforms out of range", i);
    }
    for (int i = 0; i < s->n_face_sets; i++) {
        const rpx_face_set& fs = s->face_sets[i];
        if (fs.face_begin < 0 || fs.face_end < fs.face_begin || fs.face_end > s->n_traced_faces)
            return fail(ctx, RPX_ERR_INVALID, "face set %d: bad face range", i);
    }
    for (int i = 0; i < s->n_materials; i++) {
        const rpx_material& m = s->materials[i];
        if (m.type < RPX_MAT_OPAQUE || m.type > RPX_MAT_RECT_APERTURE)
            return fail(ctx, RPX_ERR_UNSUPPORTED, "material %d: unsupported material type %d", i, m.type);
        bool needs_tab = m.type == RPX_MAT_DIELECTRIC || m.type == RPX_MAT_FULL_DIELECTRIC || m.type == RPX_MAT_COATED;
        if (needs_tab && (m.ntab_off < 0 || m.ntab_off + 3 * s->n_wavelengths > s->n_ntab))
            return fail(ctx, RPX_ERR_INVALID, "material %d: n-table out of range", i);
    }
    for (int i = 0; i < s->n_distortions; i++) {
        const rpx_distortion& d = s->distortions[i];
        if (d.type == RPX_DIST_ZERNIKE) {
            if (d.k_max > RPX_ZERNIKE_MAX_K)
                return fail(ctx, RPX_ERR_UNSUPPORTED, "distortion %d: k_max %d > %d", i, d.k_max, RPX_ZERNIKE_MAX_K);
            if (d.coef_off < 0 || d.coef_off + d.n_coefs > s->n_zcoefs || d.tape_z_off + d.tape_z_len > s->n_ztape ||
                d.tape_g_off + d.tape_g_len > s->n_ztape)
                return fail(ctx, RPX_ERR_INVALID, "distortion %d: tables out of range", i);
        } else if (d.type != RPX_DIST_ZERNIKE_J7) {
            return fail(ctx, RPX_ERR_UNSUPPORTED, "distortion %d: unsupported type %d", i, d.type);
        }
    }
    for (int i = 0; i < s->n_ztape; i++) {
        const rpx_ztape_op& t = s->ztape[i];
        const int lim = 2 + 3 * RPX_ZERNIKE_MAX_K;
        if (t.dst < 0 || t.dst >= RPX_ZERNIKE_MAX_K || t.a < 0 || t.a >= lim || t.b < 0 || t.b >= lim ||
            t.c < 0 || t.c >= lim || t.d < 0 || t.d >= lim || t.e < 0 || t.e >= lim)
            return fail(ctx, RPX_ERR_INVALID, "zernike tape op %d out of range", i);
    }
    return RPX_OK;
}

// The BVH of every mesh / UV patch face repacked for the device walk (rpx_faces.cuh::mesh_intersect): one 64-byte
// node per INNER node of the tree, holding both children's boxes as fp32 rounded outwards and two child references
// (>= 0: packed node, < 0: leaf -(first * 8 + count - 1) - 1, RPX_BVH32_NONE: no child).  A tree the packed format
// cannot hold (a leaf of more than 8 triangles, 2^27 triangles, depth over 46) keeps off[face] = -1 and is walked
// through its fp64 nodes.  RPX_MESH_F64=1 in the environment forces that for every face (A/B measurements).
static float f32_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}
static float f32_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}
static void pack_mesh_bvh(const rpx_scene* s, std::vector<float>& packed, std::vector<int>& off) {
    off.assign((size_t)(s->n_faces > 0 ? s->n_faces : 1), -1);
    const char* env = getenv("RPX_MESH_F64");
    if (env && env[0] == '1') return;
    for (int i = 0; i < s->n_faces; i++) {
        const rpx_face& f = s->faces[i];
        if (f.type != RPX_FACE_MESH && f.type != RPX_FACE_UVPATCH) continue;
        const double* H = s->pool + f.aux_off;
        const long long n_cells = (long long)H[1], n_nodes = (long long)H[2];
        const double* nodes = H + (long long)H[6];
        if (n_cells >= (1ll << 27)) continue;
        bool ok = true;
        std::vector<int> height((size_t)n_nodes, 1);
        for (long long k = n_nodes - 1; k >= 0 && ok; k--) {
            const double a = nodes[8 * k + 6], b = nodes[8 * k + 7];
            if (a >= 0)
                height[(size_t)k] = 1 + std::max(height[(size_t)a], height[(size_t)b]);
            else if (b > 8)
                ok = false;
        }
        if (!ok || height[0] > 46) continue;
        // inner nodes -> packed ids in index order (children have larger ids than their parent)
        std::vector<int> pid((size_t)n_nodes, -1);
        int n_inner = 0;
        for (long long k = 0; k < n_nodes; k++)
            if (nodes[8 * k + 6] >= 0) pid[(size_t)k] = n_inner++;
        const size_t base = packed.size() / 16;
        const bool root_leaf = nodes[6] < 0;
        packed.resize(packed.size() + 16 * (size_t)(root_leaf ? 1 : n_inner), 0.0f);
        auto ref_of = [&](long long k) -> int {
            const double a = nodes[8 * k + 6], b = nodes[8 * k + 7];
            if (a >= 0) return pid[(size_t)k];
            return -(int)((long long)(-a - 1) * 8 + ((long long)b - 1)) - 1;
        };
        auto put = [&](float* n, int slot, long long k) {  // child `slot` of packed node n := tree node k
            float* lo = slot == 0 ? n + 0 : n + 6;
            float* hi = lo + 3;
            for (int c = 0; c < 3; c++) {
                lo[c] = f32_down(nodes[8 * k + c]);
                hi[c] = f32_up(nodes[8 * k + 3 + c]);
            }
            const int r = ref_of(k);
            memcpy(n + 12 + slot, &r, 4);
        };
        auto put_none = [&](float* n, int slot) {
            float* lo = slot == 0 ? n + 0 : n + 6;
            for (int c = 0; c < 6; c++) lo[c] = 3.0e38f;
            const int r = RPX_BVH32_NONE;
            memcpy(n + 12 + slot, &r, 4);
        };
        float* P = packed.data() + 16 * base;
        if (root_leaf) {
            put(P, 0, 0);
            put_none(P, 1);
        } else {
            for (long long k = 0; k < n_nodes; k++) {
                if (pid[(size_t)k] < 0) continue;
                float* n = P + 16 * (size_t)pid[(size_t)k];
                put(n, 0, (long long)nodes[8 * k + 6]);
                put(n, 1, (long long)nodes[8 * k + 7]);
            }
        }
        double m = 0.0;  // largest |coordinate| of the root box: scales the padding of the fp32 slab test
        for (int c = 0; c < 6; c++) m = std::max(m, fabs(nodes[c]));
        P[14] = f32_up(m);
        off[(size_t)i] = (int)base;
    }
}

// Copy the flat scene tables into ONE device block and point a DevScene at them.
// Layout: faces | sets | materials | shape ops | implicit ops | distortions | zcoefs | ztape |
//         wavelengths | ntab | pool
static int upload_scene(rpx_ctx* ctx, const rpx_scene* s, DevScene* out, void** block) {
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<float> bvh32;
    std::vector<int> mesh32_off;
    pack_mesh_bvh(s, bvh32, mesh32_off);
    Part parts[13] = {
        {s->faces, (size_t)s->n_faces * sizeof(rpx_face), 0},
        {s->face_sets, (size_t)s->n_face_sets * sizeof(rpx_face_set), 0},
        {s->materials, (size_t)s->n_materials * sizeof(rpx_material), 0},
        {s->shape_ops, (size_t)s->n_shape_ops * sizeof(rpx_shape_op), 0},
        {s->implicit_ops, (size_t)s->n_implicit_ops * sizeof(rpx_implicit_op), 0},
        {s->distortions, (size_t)s->n_distortions * sizeof(rpx_distortion), 0},
        {s->zcoefs, (size_t)s->n_zcoefs * sizeof(rpx_zcoef), 0},
        {s->ztape, (size_t)s->n_ztape * sizeof(rpx_ztape_op), 0},
        {s->wavelengths, (size_t)s->n_wavelengths * sizeof(double), 0},
        {s->ntab, (size_t)s->n_ntab * 2 * sizeof(double), 0},
        {s->pool, (size_t)s->n_pool * sizeof(double), 0},
        {bvh32.data(), bvh32.size() * sizeof(float), 0},
        {mesh32_off.data(), mesh32_off.size() * sizeof(int), 0},
    };
    size_t total = 0;
    for (Part& p : parts) {
        p.off = total;
        total += align_up(p.bytes ? p.bytes : 8, 256);
    }
    std::vector<unsigned char> host(total, 0);
    for (Part& p : parts)
        if (p.bytes) {
            if (!p.src) return fail(ctx, RPX_ERR_INVALID, "scene table pointer is NULL");
            memcpy(host.data() + p.off, p.src, p.bytes);
        }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*block) {
        CU(ctx, cudaFree(*block));
        *block = nullptr;
    }
    CU(ctx, cudaMalloc(block, total));
    CU(ctx, cudaMemcpy(*block, host.data(), total, cudaMemcpyHostToDevice));
    unsigned char* b = (unsigned char*)*block;
    DevScene& d = *out;
    d.faces = (const rpx_face*)(b + parts[0].off);
    d.sets = (const rpx_face_set*)(b + parts[1].off);
    d.mats = (const rpx_material*)(b + parts[2].off);
    d.shape_ops = (const rpx_shape_op*)(b + parts[3].off);
    d.impl_ops = (const rpx_implicit_op*)(b + parts[4].off);
    d.dists = (const rpx_distortion*)(b + parts[5].off);
    d.zcoefs = (const rpx_zcoef*)(b + parts[6].off);
    d.ztape = (const rpx_ztape_op*)(b + parts[7].off);
    d.wavelengths = (const double*)(b + parts[8].off);
    d.ntab = (const double*)(b + parts[9].off);
    d.pool = (const double*)(b + parts[10].off);
    d.bvh32 = (const float4*)(b + parts[11].off);
    d.mesh32_off = (const int*)(b + parts[12].off);
    d.n_traced = s->n_traced_faces;
    d.n_faces = s->n_faces;
    d.n_sets = s->n_face_sets;
    d.n_mats = s->n_materials;
    d.n_wl = s->n_wavelengths;
    d.n_dists = s->n_distortions;
    return RPX_OK;
}

static int scene_face_class(const rpx_scene* s) {
    for (int i = 0; i < s->n_faces; i++)
        if (s->faces[i].type == RPX_FACE_MESH || s->faces[i].type == RPX_FACE_UVPATCH) return RPX_FC_MESH;
    for (int i = 0; i < s->n_faces; i++) {
        int t = s->faces[i].type;
        bool simple = t == RPX_FACE_CIRCULAR || t == RPX_FACE_SHAPED_PLANAR || t == RPX_FACE_ELLIPTICAL_PLANE ||
                      t == RPX_FACE_RECTANGULAR || t == RPX_FACE_SPHERICAL || t == RPX_FACE_SHAPED_SPHERICAL ||
                      t == RPX_FACE_EXTRUDED_PLANAR || t == RPX_FACE_POLYGON || t == RPX_FACE_ORIENTED_POLYGON;
        if (!simple) return RPX_FC_FULL;
    }
    return RPX_FC_SIMPLE;
}

// Bytes of scene tables the kernels stage in shared memory (faces, face-set transforms, materials);
// 0 = they do not fit: the SS=false kernel instantiations read them from global memory instead.
static int scene_smem_bytes(const rpx_scene* s) {
    size_t smem = (size_t)s->n_faces * sizeof(rpx_face) + (size_t)s->n_face_sets * sizeof(rpx_face_set) +
                  (size_t)s->n_materials * sizeof(rpx_material);
    return smem <= 40 * 1024 ? (int)smem : 0;
}

extern "C" int rpx_scene_set(rpx_ctx* ctx, const rpx_scene* s) {
    if (!ctx || !s) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = validate_scene(ctx, s);
    if (rc != RPX_OK) return rc;
    rc = upload_scene(ctx, s, &ctx->ds, &ctx->scene_block);
    if (rc != RPX_OK) return rc;
    ctx->n_traced = s->n_traced_faces;
    ctx->max_kids = 0;
    for (int i = 0; i < s->n_traced_faces; i++) {
        int t = s->materials[s->faces[i].material].type;
        int kids = (t == RPX_MAT_OPAQUE) ? 0
                   : (t == RPX_MAT_PARTIALLY_REFLECTIVE || t == RPX_MAT_LINEAR_POLARISING ||
                      t == RPX_MAT_FULL_DIELECTRIC || t == RPX_MAT_COATED) ? 2 : 1;
        if (kids > ctx->max_kids) ctx->max_kids = kids;
    }
    // kernel variant: the smallest compiled face class / material mask covering the scene
    ctx->face_class = scene_face_class(s);
    uint32_t used = 0;
    for (int i = 0; i < s->n_traced_faces; i++) used |= RPX_MBIT(s->materials[s->faces[i].material].type);
    const uint32_t masks[RPX_N_MM] = {RPX_MM_LIGHT, RPX_MM_COATED, RPX_MM_FULLDIEL, RPX_MM_ALL};
    ctx->mm_idx = RPX_N_MM - 1;
    for (int m = 0; m < RPX_N_MM; m++)
        if ((used & ~masks[m]) == 0) {
            ctx->mm_idx = m;
            break;
        }
    ctx->scene_smem = scene_smem_bytes(s);
    if (ctx->d_face_counts) {
        CU(ctx, cudaFree(ctx->d_face_counts));
        ctx->d_face_counts = nullptr;
    }
    CU(ctx, cudaMalloc(&ctx->d_face_counts, sizeof(uint32_t) * (size_t)(s->n_traced_faces > 0 ? s->n_traced_faces : 1)));
    ctx->have_scene = true;
    return RPX_OK;
}

extern "C" int rpx_capture_scene_set(rpx_ctx* ctx, const rpx_scene* s, const uint32_t* face_ids) {
    if (!ctx || !s) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = validate_scene(ctx, s);
    if (rc != RPX_OK) return rc;
    ctx->have_capture = false;
    rc = upload_scene(ctx, s, &ctx->cap_ds, &ctx->cap_block);
    if (rc != RPX_OK) return rc;
    if (ctx->cap_face_ids) {
        CU(ctx, cudaFree(ctx->cap_face_ids));
        ctx->cap_face_ids = nullptr;
    }
    if (face_ids && s->n_faces > 0) {
        CU(ctx, cudaMalloc(&ctx->cap_face_ids, sizeof(uint32_t) * (size_t)s->n_faces));
        CU(ctx, cudaMemcpy(ctx->cap_face_ids, face_ids, sizeof(uint32_t) * (size_t)s->n_faces, cudaMemcpyHostToDevice));
    }
    ctx->cap_face_class = scene_face_class(s);
    ctx->cap_smem = scene_smem_bytes(s);
    ctx->have_capture = true;
    return RPX_OK;
}

// ------------------------------------------------------------------ ray buffers
int rpx_rays_alloc(rpx_ctx* ctx, unsigned long long cap_req, int is_gausslet, rpx_rays** out) {
    rpx_rays* r = new (std::nothrow) rpx_rays();
    if (!r) return fail(ctx, RPX_ERR_NOMEM, "out of host memory");
    // every field array 1-KB aligned and a whole number of tiles (TMA bulk copies fetch full tiles)
    unsigned long long cap = (cap_req + (RPX_TILE - 1ull)) / RPX_TILE * RPX_TILE;
    if (cap == 0) cap = RPX_TILE;
    size_t fb = (size_t)NF * cap * sizeof(double);
    size_t ub = (size_t)NU * cap * sizeof(uint32_t);
    size_t pb = is_gausslet ? (size_t)NP * cap * sizeof(double) : 0;
    size_t qb = (size_t)cap * sizeof(uint32_t);  // facet record of the hit (mesh / UV patch scenes; untouched otherwise)
    r->bytes = fb + ub + pb + qb;
    r->is_gausslet = is_gausslet;
    cudaError_t e = cudaMallocAsync(&r->block, r->bytes, ctx->stream);
    if (e != cudaSuccess) {
        delete r;
        return fail(ctx, RPX_ERR_NOMEM, "cannot allocate %zu bytes for a generation of %llu %s: %s", fb + ub + pb + qb,
                    cap, is_gausslet ? "gausslets" : "rays", cudaGetErrorString(e));
    }
    unsigned char* b = (unsigned char*)r->block;
    r->soa.f = (double*)b;
    r->soa.p = is_gausslet ? (double*)(b + fb) : nullptr;
    r->soa.u = (uint32_t*)(b + fb + pb);
    r->soa.piece = (uint32_t*)(b + fb + pb + ub);
    r->soa.n = 0;
    r->soa.cap = cap;
    *out = r;
    return RPX_OK;
}

extern "C" void rpx_rays_free(rpx_ctx* ctx, rpx_rays* rays) {
    if (!rays) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaFreeAsync(rays->block, ctx->stream);
    }
    delete rays;
}

extern "C" uint64_t rpx_rays_count(const rpx_rays* rays) { return rays ? rays->soa.n : 0; }

extern "C" int rpx_rays_clone(rpx_ctx* ctx, const rpx_rays* rays, rpx_rays** out_rays) {
    if (!ctx || !rays || !out_rays) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    rpx_rays* r = nullptr;
    int rc = rpx_rays_alloc(ctx, rays->soa.cap, rays->is_gausslet, &r);
    if (rc != RPX_OK) return rc;
    r->soa.n = rays->soa.n;
    if (r->bytes != rays->bytes) {
        rpx_rays_free(ctx, r);
        return fail(ctx, RPX_ERR_INVALID, "clone size mismatch");
    }
    cudaError_t e = cudaMemcpyAsync(r->block, rays->block, rays->bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) {
        rpx_rays_free(ctx, r);
        return fail(ctx, RPX_ERR_CUDA, "device copy failed: %s", cudaGetErrorString(e));
    }
    *out_rays = r;
    return RPX_OK;
}

static cudaEvent_t next_event(rpx_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        ctx->ev_pool.push_back(ev);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

extern "C" int rpx_rays_upload(rpx_ctx* ctx, const void* aos, uint64_t n, int is_gausslet, rpx_rays** out_rays) {
    if (!ctx || !out_rays || (!aos && n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n >= 0xFFFFFFFFull)
        return fail(ctx, RPX_ERR_INVALID, "%llu rays per generation exceed the 32-bit parent_idx of ray_t",
                    (unsigned long long)n);
    rpx_rays* r = nullptr;
    int rc = rpx_rays_alloc(ctx, n, is_gausslet, &r);
    if (rc != RPX_OK) return rc;
    r->soa.n = n;
    if (n) {
        const size_t rec = is_gausslet ? RPX_GAUSSLET_BYTES : RPX_RAY_BYTES;
        void* d_aos = nullptr;
        CU(ctx, cudaMallocAsync(&d_aos, n * rec, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_aos, aos, n * rec, cudaMemcpyHostToDevice, ctx->stream));
        if (is_gausslet) {
            const int T = 64;
            k_aos_to_soa<RPX_WORDS_GAUSSLET, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, ctx->stream>>>(
                (const uint32_t*)d_aos, r->soa);
        } else {
            const int T = 256;
            k_aos_to_soa<RPX_WORDS_RAY, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_RAY_BYTES, ctx->stream>>>(
                (const uint32_t*)d_aos, r->soa);
        }
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaFreeAsync(d_aos, ctx->stream));
        // the caller may reuse `aos` as soon as we return
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out_rays = r;
    return RPX_OK;
}

extern "C" int rpx_rays_download(rpx_ctx* ctx, const rpx_rays* rays, void* out_aos, uint64_t capacity) {
    if (!ctx || !rays || (!out_aos && rays->soa.n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = rays->soa.n;
    if (capacity < n) return fail(ctx, RPX_ERR_INVALID, "output holds %llu records, generation has %llu",
                                  (unsigned long long)capacity, (unsigned long long)n);
    if (!n) return RPX_OK;
    const size_t rec = rays->is_gausslet ? RPX_GAUSSLET_BYTES : RPX_RAY_BYTES;
    void* d_aos = nullptr;
    CU(ctx, cudaMallocAsync(&d_aos, n * rec, ctx->stream));
    if (rays->is_gausslet) {
        const int T = 64;
        k_soa_to_aos<RPX_WORDS_GAUSSLET, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, ctx->stream>>>(
            rays->soa, (uint32_t*)d_aos, 0u);
    } else {
        const int T = 256;
        k_soa_to_aos<RPX_WORDS_RAY, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_RAY_BYTES, ctx->stream>>>(
            rays->soa, (uint32_t*)d_aos, 0u);
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out_aos, d_aos, n * rec, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaFreeAsync(d_aos, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return RPX_OK;
}

// ------------------------------------------------------------------ the generation loop
extern "C" void rpx_result_free(rpx_ctx* ctx, rpx_result* res) {
    if (!res) return;
    for (rpx_rays* g : res->gens) rpx_rays_free(ctx, g);
    delete res;
}

// ------------------------------------------------------------------ pipelined generation loop
// The exact loop below learns len(new_rays) with a host round trip per generation, which leaves
// the GPU idle for ~20 us each time (14 % of a 1e6-ray achromat trace).  Here generation g's
// kernel is enqueued as soon as the count of generation g-1 is known: its grid and its output
// buffer are sized from the bound  n_g <= kids * n_{g-1}  and the kernel reads the real n_g from
// device memory.  The host only ever waits for a count that is one generation old, i.e. for a
// kernel that has already finished, so kernels run back to back.  Used for non-sequential,
// keep-everything traces whose buffers stay small enough for the 2x over-allocation.
static int trace_pipelined(rpx_ctx* ctx, rpx_rays* rays, double ml, int recursion_limit, rpx_result* res) {
    cudaStream_t st = ctx->stream;
    const int is_g = rays->is_gausslet;
    const int smem = ctx->scene_smem;
    const unsigned long long kids = (unsigned long long)(ctx->max_kids > 0 ? ctx->max_kids : 1);
    std::vector<rpx_rays*> bufs;       // bufs[g] = generation g (bound-sized for g >= 1)
    std::vector<cudaEvent_t> ev_cnt;   // ev_cnt[g]: h_counts[g] is valid once it completed
    std::vector<cudaEvent_t> ev_i0, ev_i1, ev_s0, ev_s1;
    bufs.push_back(rays);
    auto bail = [&](int code) {
        cudaStreamSynchronize(st);
        for (rpx_rays* b : bufs) rpx_rays_free(ctx, b);
        delete res;
        return code;
    };
#define CUP(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(ctx, e_ == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA,   \
                             "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    const int max_g = recursion_limit < RPX_MAX_PIPE_GENS - 2 ? recursion_limit : RPX_MAX_PIPE_GENS - 2;
    cudaEvent_t ev_begin = next_event(ctx), ev_end = next_event(ctx);
    CUP(cudaMemsetAsync(ctx->d_face_counts, 0, sizeof(uint32_t) * (size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1), st));
    CUP(cudaMemsetAsync(ctx->d_counts, 0, sizeof(unsigned long long) * RPX_MAX_PIPE_GENS, st));
    CUP(cudaMemsetAsync(ctx->pipe_counters, 0, sizeof(uint32_t) * RPX_MAX_PIPE_GENS, st));
    CUP(cudaMemsetAsync(ctx->pipe_hits, 0, sizeof(uint32_t) * RPX_MAX_PIPE_GENS, st));
    CUP(cudaMemsetAsync(ctx->pipe_miss, 0, sizeof(uint32_t) * RPX_MAX_PIPE_GENS, st));
    for (int i = 0; i < RPX_MAX_PIPE_GENS; i++) ctx->h_counts[i] = 0;  // a kernel that never runs leaves 0
    size_t state_off = 0, zero_end = 0;
    ctx->h_counts[0] = rays->soa.n;
    CUP(cudaMemcpyAsync(ctx->d_counts, ctx->h_counts, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    CUP(cudaEventRecord(ev_begin, st));
    ev_cnt.push_back(ev_begin);  // n_0 is known
    if (is_g && rays->soa.n) {
        k_reset_length<<<(unsigned)((rays->soa.n + 255) / 256), 256, 0, st>>>(rays->soa, ml);
        res->launches++;
    }
    {   // generation 0: nearest hit
        const uint32_t n_tiles = (uint32_t)((rays->soa.n + RPX_TILE - 1) / RPX_TILE);
        cudaEvent_t a0 = next_event(ctx), a1 = next_event(ctx);
        CUP(cudaEventRecord(a0, st));
        CUP(launch_intersect(ctx->face_class, st, n_tiles, smem, ctx->ds, rays->soa, ml, -1));
        CUP(cudaEventRecord(a1, st));
        ev_i0.push_back(a0);
        ev_i1.push_back(a1);
        res->launches++;
    }
    // tile state: one region per launch so that no memset has to wait between kernels
    unsigned long long bound = rays->soa.n;  // upper bound on n_g (exact for g == 0)
    int g = 0;
    for (; g < max_g; g++) {
        if (g >= 1) {
            // n_g <= kids * n_{g-1}.  n_{g-1} is ONE generation old: the kernel that produced it
            // finished before the one now running started, so this wait never stalls the GPU.
            CUP(cudaEventSynchronize(ev_cnt[g - 1]));
            const unsigned long long n_prev = ctx->h_counts[g - 1];
            if (n_prev == 0) break;  // generation g-1 is empty: nothing left to trace
            bound = kids * n_prev;
            if (bound > bufs[g]->soa.cap) bound = bufs[g]->soa.cap;
        }
        if (bound == 0) break;
        if (kids * bound >= 0xFFFFFFFFull)
            return bail(fail(ctx, RPX_ERR_INVALID, "generation would exceed the 32-bit parent_idx of ray_t"));
        rpx_rays* child = nullptr;
        {
            int rc = rpx_rays_alloc(ctx, kids * bound, is_g, &child);
            if (rc != RPX_OK) return bail(rc);
        }
        bufs.push_back(child);
        const uint32_t n_tiles = (uint32_t)((bound + RPX_TILE - 1) / RPX_TILE);
        // look-back state: a fresh, already-zeroed slice per generation, so that NO memset sits
        // between two kernels; the zeroed frontier is pushed ahead in large steps
        const size_t state_words = rpx_state_words(n_tiles);
        if (state_off + state_words > ctx->pipe_state_cap) {  // wrap: everything before is finished by then
            state_off = 0;
            zero_end = 0;
        }
        if (state_off + state_words > zero_end) {
            size_t want = state_off + state_words * 6;
            if (want > ctx->pipe_state_cap) want = ctx->pipe_state_cap;
            CUP(cudaMemsetAsync(ctx->pipe_state + zero_end, 0, (want - zero_end) * sizeof(unsigned long long), st));
            zero_end = want;
        }
        unsigned long long* state = ctx->pipe_state + state_off;
        state_off += state_words;
        ShadeArgs sa;
        sa.S = ctx->ds;
        sa.in = bufs[g]->soa;
        sa.in.n = bound;
        sa.out = child->soa;
        sa.max_length = ml;
        sa.tile_state = state;
        sa.tile_counter = ctx->pipe_counters + g;
        sa.d_count = ctx->d_counts + (g + 1);
        sa.face_counts = ctx->d_face_counts;
        sa.n_tiles = n_tiles;
        sa.smem_bytes = smem;
        sa.ahead_face = -1;
        sa.n_dev = ctx->d_counts + g;
        sa.h_count = ctx->h_counts_dev + (g + 1);
        sa.hits_in = g >= 1 ? ctx->pipe_hits + g : nullptr;  // generation 0 was intersected by k_intersect
        sa.hits_out = ctx->pipe_hits + (g + 1);
        sa.miss_in = g >= 1 ? ctx->pipe_miss + g : nullptr;
        sa.miss_out = ctx->pipe_miss + (g + 1);
        cudaEvent_t b0 = next_event(ctx), b1 = next_event(ctx);
        CUP(cudaEventRecord(b0, st));
        CUP(shade_launcher(is_g, ctx->face_class, ctx->mm_idx, smem > 0)(st, sa));
        CUP(cudaEventRecord(b1, st));  // also marks h_counts[g + 1] valid
        ev_s0.push_back(b0);
        ev_s1.push_back(b1);
        ev_cnt.push_back(b1);
        res->launches++;
    }
    CUP(cudaEventRecord(ev_end, st));
    CUP(cudaMemcpyAsync(res->face_counts.data(), ctx->d_face_counts, sizeof(uint32_t) * (size_t)ctx->n_traced,
                        cudaMemcpyDeviceToHost, st));
    CUP(cudaStreamSynchronize(st));
    // traced_rays = generations 0..G-1 with n_g > 0 and g < recursion_limit
    const int launched = (int)bufs.size();  // bufs[0..launched-1]; counts known for all of them
    int G = 0;
    while (G < launched && G < recursion_limit && ctx->h_counts[G] > 0) G++;
    for (int i = 0; i < launched; i++) {
        if (i < G) {
            bufs[i]->soa.n = ctx->h_counts[i];
            res->gens.push_back(bufs[i]);
            res->counts.push_back(ctx->h_counts[i]);
        } else {
            rpx_rays_free(ctx, bufs[i]);
        }
    }
    bufs.clear();
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    res->device_ms = ms;
    for (size_t i = 0; i < ev_i0.size(); i++) {
        cudaEventElapsedTime(&ms, ev_i0[i], ev_i1[i]);
        res->k_ms[0] += ms;
        res->k_launches[0]++;
    }
    // only kernels that had work count towards the per-launch average
    for (size_t i = 0; i < ev_s0.size() && (int)i < G; i++) {
        cudaEventElapsedTime(&ms, ev_s0[i], ev_s1[i]);
        res->k_ms[1] += ms;
        res->k_launches[1]++;
    }
#undef CUP
    return RPX_OK;
}

// The generation loop.  face_seq == NULL: non-sequential trace_rays (core/tracer.py:39-45).
// face_seq != NULL: trace_ray_sequence (core/tracer.py:84-97): step s intersects only face
// face_seq[s]; the generation produced by the last step is appended untraced.
static int trace_loop(rpx_ctx* ctx, rpx_rays* rays, double max_length, int recursion_limit, uint32_t flags,
                      const int32_t* face_seq, int n_seq, rpx_result** out_result) {
    if (!ctx || !rays || !out_result) {
        if (ctx && rays) rpx_rays_free(ctx, rays);
        return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    }
    if (!ctx->have_scene) {
        rpx_rays_free(ctx, rays);
        return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    }
    for (int s = 0; face_seq && s < n_seq; s++)
        if (face_seq[s] < 0 || face_seq[s] >= ctx->n_traced) {
            rpx_rays_free(ctx, rays);
            return fail(ctx, RPX_ERR_INVALID, "face sequence entry %d (= %d) out of range", s, face_seq[s]);
        }
    const bool sequential = face_seq != nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int is_g = rays->is_gausslet;
    // `float max_length` of trace_segment_c (ctracer.pyx:2066); trace_gausslet_c takes a double
    const double ml = is_g ? max_length : (double)(float)max_length;
    rpx_result* res = new (std::nothrow) rpx_result();
    if (!res) return fail(ctx, RPX_ERR_NOMEM, "out of host memory");
    res->device_ms = 0;
    res->launches = 0;
    res->k_ms[0] = res->k_ms[1] = 0;
    res->k_launches[0] = res->k_launches[1] = 0;
    res->face_counts.assign((size_t)ctx->n_traced, 0u);
    ctx->ev_used = 0;
    std::vector<cudaEvent_t> ev_i0, ev_i1, ev_s0, ev_s1;

    // Mesh / UV patch scenes run UNFUSED: k_shade leaves its children untraced and k_intersect finds every
    // generation's nearest hits.  The BVH walk lives on L1 hits (upper tree levels, per-thread stacks); inside
    // k_shade the child staging (47 KB per CTA when this was measured) left it ~60 KB of L1 per SM (ncu: 58 % L1 hit rate against 79 % in
    // k_intersect), and the walk cost 1.21 ms per 1e6 rays there against 0.75 ms in k_intersect -- far more than the
    // 60 bytes per ray the extra pass moves.  RPX_MESH_FUSED=1 keeps the fused trace-ahead (A/B measurements).
    static const bool mesh_fused = [] { const char* e = getenv("RPX_MESH_FUSED"); return e && e[0] == '1'; }();
    const bool unfused = ctx->face_class == RPX_FC_MESH && !mesh_fused;
    // small / medium keep-everything traces: pipelined launches (no per-generation host stall)
    if (!unfused) {
        const size_t rec = is_g ? RPX_GAUSSLET_BYTES : RPX_RAY_BYTES;
        const unsigned long long kids = (unsigned long long)(ctx->max_kids > 0 ? ctx->max_kids : 1);
        const bool fits = (double)rays->soa.n * (double)rec * (double)(kids * kids) <= 4.0e9;
        if (!sequential && !(flags & (RPX_TRACE_KEEP_LAST_ONLY | RPX_TRACE_EXACT_SYNC)) && fits && rays->soa.n > 0 &&
            recursion_limit <= RPX_MAX_PIPE_GENS - 2) {
            int rc = trace_pipelined(ctx, rays, ml, recursion_limit, res);
            if (rc != RPX_OK) return rc;
            *out_result = res;
            return RPX_OK;
        }
    }

    // Ownership: `rays` belongs to this call from here on (success or failure).  `cur` is
    // either the last entry of res->gens or not yet part of it; `child` likewise.
    rpx_rays* cur = rays;
    rpx_rays* child = nullptr;
    auto bail = [&](int code) {
        if (child && child != cur) rpx_rays_free(ctx, child);
        if (cur && (res->gens.empty() || res->gens.back() != cur)) rpx_rays_free(ctx, cur);
        rpx_result_free(ctx, res);
        return code;
    };
#define CUR(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(ctx, e_ == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA,   \
                             "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)

    cudaEvent_t ev_begin = next_event(ctx), ev_end = next_event(ctx);
    CUR(cudaMemsetAsync(ctx->d_face_counts, 0, sizeof(uint32_t) * (size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1), st));
    CUR(cudaMemsetAsync(ctx->pipe_hits, 0, sizeof(uint32_t) * RPX_MAX_PIPE_GENS, st));
    CUR(cudaMemsetAsync(ctx->pipe_miss, 0, sizeof(uint32_t) * RPX_MAX_PIPE_GENS, st));
    CUR(cudaEventRecord(ev_begin, st));
    if (is_g && cur->soa.n) {  // input_rays.reset_length(max_length), core/tracer.py:22
        k_reset_length<<<(unsigned)((cur->soa.n + 255) / 256), 256, 0, st>>>(cur->soa, ml);
        res->launches++;
    }
    int count = 0;
    const int smem = ctx->scene_smem;
    // sequential mode: traced_rays starts as [input_rays] and one step runs per sequence entry
    while (sequential ? (count < n_seq && cur->soa.n > 0) : (cur->soa.n > 0 && count < recursion_limit)) {
        const unsigned long long n = cur->soa.n;
        if (!sequential || count == 0) {
            res->gens.push_back(cur);
            res->counts.push_back(n);
        }
        const uint32_t n_tiles = (uint32_t)((n + RPX_TILE - 1) / RPX_TILE);   // CTA tiles of k_intersect
        const uint32_t n_wtiles = n_tiles;
        // ---- nearest hit: generation 0 only (k_shade traces its children ahead), every generation when unfused
        if (count == 0 || unfused) {
            cudaEvent_t a0 = next_event(ctx), a1 = next_event(ctx);
            CUR(cudaEventRecord(a0, st));
            CUR(launch_intersect(ctx->face_class, st, n_tiles, smem, ctx->ds, cur->soa, ml, sequential ? face_seq[count] : -1));
            CUR(cudaEventRecord(a1, st));
            ev_i0.push_back(a0);
            ev_i1.push_back(a1);
            res->launches++;
        }
        // ---- children
        unsigned long long cap_child = n * (unsigned long long)(ctx->max_kids > 0 ? ctx->max_kids : 1);
        if (cap_child >= 0xFFFFFFFFull)
            return bail(fail(ctx, RPX_ERR_INVALID, "generation would exceed the 32-bit parent_idx of ray_t"));
        {
            int rc = rpx_rays_alloc(ctx, cap_child, is_g, &child);
            if (rc != RPX_OK) return bail(rc);
        }
        const size_t state_words = rpx_state_words(n_wtiles);
        if (state_words > ctx->tile_state_cap) {
            if (ctx->tile_state) CUR(cudaFree(ctx->tile_state));
            ctx->tile_state = nullptr;
            ctx->tile_state_cap = 0;
            size_t cap_t = state_words * 2;
            CUR(cudaMalloc(&ctx->tile_state, cap_t * sizeof(unsigned long long)));
            ctx->tile_state_cap = cap_t;
        }
        CUR(cudaMemsetAsync(ctx->tile_state, 0, state_words * sizeof(unsigned long long), st));
        CUR(cudaMemsetAsync(ctx->tile_counter, 0, sizeof(uint32_t), st));
        cudaEvent_t b0 = next_event(ctx), b1 = next_event(ctx);
        CUR(cudaEventRecord(b0, st));
        {
            ShadeArgs sa;
            sa.S = ctx->ds;
            sa.in = cur->soa;
            sa.out = child->soa;
            sa.max_length = ml;
            sa.tile_state = ctx->tile_state;
            sa.tile_counter = ctx->tile_counter;
            sa.d_count = ctx->d_count;
            sa.face_counts = ctx->d_face_counts;
            sa.n_tiles = n_wtiles;
            sa.smem_bytes = smem;
            sa.ahead_face = unfused ? -2 : !sequential ? -1 : (count + 1 < n_seq ? face_seq[count + 1] : -2);
            sa.n_dev = nullptr;
            sa.h_count = nullptr;
            if (!sequential && !unfused && count + 1 < RPX_MAX_PIPE_GENS) {  // (a sequence's last step leaves its children untraced)
                sa.hits_in = count >= 1 ? ctx->pipe_hits + count : nullptr;
                sa.hits_out = ctx->pipe_hits + (count + 1);
                sa.miss_in = count >= 1 ? ctx->pipe_miss + count : nullptr;
                sa.miss_out = ctx->pipe_miss + (count + 1);
            }
            cudaError_t le = shade_launcher(is_g, ctx->face_class, ctx->mm_idx, smem > 0)(st, sa);
            CUR(le);
        }
        CUR(cudaEventRecord(b1, st));
        ev_s0.push_back(b0);
        ev_s1.push_back(b1);
        res->launches += 1;
        // ---- len(new_rays): the only host round trip of a generation
        CUR(cudaMemcpyAsync(ctx->h_count, ctx->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUR(cudaStreamSynchronize(st));
        child->soa.n = *ctx->h_count;
        if (sequential) {
            // core/tracer.py:91-97: `if (count > recursion_limit) or (rays.n_rays==0): break`
            // comes BEFORE the append
            if (count > recursion_limit || child->soa.n == 0) {
                rpx_rays_free(ctx, child);
                child = nullptr;
                break;
            }
            res->gens.push_back(child);
            res->counts.push_back(child->soa.n);
        }
        if ((flags & RPX_TRACE_KEEP_LAST_ONLY) && !sequential) {
            // streaming mode: the parent generation is complete (write-back done); drop it
            rpx_rays_free(ctx, res->gens.back());
            res->gens.back() = nullptr;
        }
        cur = child;
        child = nullptr;
        count++;
    }
    CUR(cudaEventRecord(ev_end, st));
    // the generation that was built but never traced is not part of traced_rays
    if (res->gens.empty() || res->gens.back() != cur) {
        bool owned = false;
        for (rpx_rays* g : res->gens) owned = owned || (g == cur);
        if (!owned) rpx_rays_free(ctx, cur);
    }
    cur = nullptr;
    CUR(cudaMemcpyAsync(res->face_counts.data(), ctx->d_face_counts, sizeof(uint32_t) * (size_t)ctx->n_traced,
                        cudaMemcpyDeviceToHost, st));
    CUR(cudaStreamSynchronize(st));
    float ms = 0;
    CUR(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    res->device_ms = ms;
    for (size_t g = 0; g < ev_i0.size(); g++) {
        CUR(cudaEventElapsedTime(&ms, ev_i0[g], ev_i1[g]));
        res->k_ms[0] += ms;
        res->k_launches[0]++;
    }
    for (size_t g = 0; g < ev_s0.size(); g++) {
        CUR(cudaEventElapsedTime(&ms, ev_s0[g], ev_s1[g]));
        res->k_ms[1] += ms;
        res->k_launches[1]++;
    }
#undef CUR
    *out_result = res;
    return RPX_OK;
}

extern "C" int rpx_trace_device(rpx_ctx* ctx, rpx_rays* rays, double max_length, int recursion_limit, uint32_t flags,
                                rpx_result** out_result) {
    return trace_loop(ctx, rays, max_length, recursion_limit, flags, nullptr, 0, out_result);
}

extern "C" int rpx_trace_sequence(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                                  int recursion_limit, const int32_t* face_seq, int n_seq, rpx_result** out_result) {
    if (!ctx || !out_result || !face_seq || n_seq < 1) return fail(ctx, RPX_ERR_INVALID, "bad face sequence");
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    rpx_rays* r = nullptr;
    int rc = rpx_rays_upload(ctx, rays_aos, n, is_gausslet, &r);
    if (rc != RPX_OK) return rc;
    return trace_loop(ctx, r, max_length, recursion_limit, 0, face_seq, n_seq, out_result);
}

extern "C" int rpx_trace(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                         int recursion_limit, uint32_t flags, rpx_result** out_result) {
    if (!ctx || !out_result) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    rpx_rays* r = nullptr;
    int rc = rpx_rays_upload(ctx, rays_aos, n, is_gausslet, &r);
    if (rc != RPX_OK) return rc;
    return rpx_trace_device(ctx, r, max_length, recursion_limit, flags, out_result);  // owns r either way
}

// ------------------------------------------------------------------ one generation at a time
// trace_segment_c / trace_gausslet_c as ONE call (ctracer.pyx:2062-2118, 2214-2281) for traces that
// need the host between generations: ResampleGaussletMaterial hands the gausslets that reached it to a
// Python callback and appends what it returns to the new generation (cmaterials.pyx:1766-1831,
// ctracer.pyx:2274-2278).  `rays` is intersected with every face and written back in place (length,
// end_face_idx, parabasal lengths); the children come back un-intersected, as the materials left them
// (length INF, or max_length for gausslets: reset_length_c, :2280) -- the next step intersects them.
extern "C" int rpx_trace_step(rpx_ctx* ctx, rpx_rays* rays, double max_length, rpx_rays** out_children,
                              uint32_t* face_counts) {
    if (!ctx || !rays || !out_children) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    *out_children = nullptr;
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int is_g = rays->is_gausslet;
    const double ml = is_g ? max_length : (double)(float)max_length;  // `float max_length`, ctracer.pyx:2066
    const unsigned long long n = rays->soa.n;
    const unsigned long long kids = (unsigned long long)(ctx->max_kids > 0 ? ctx->max_kids : 1);
    if (n * kids >= 0xFFFFFFFFull) return fail(ctx, RPX_ERR_INVALID, "generation would exceed the 32-bit parent_idx of ray_t");
    rpx_rays* child = nullptr;
    int rc = rpx_rays_alloc(ctx, n * kids, is_g, &child);
    if (rc != RPX_OK) return rc;
    if (n == 0) {
        *out_children = child;
        return RPX_OK;
    }
    auto bail = [&](cudaError_t e, const char* what) {
        rpx_rays_free(ctx, child);
        return fail(ctx, e == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    const uint32_t n_tiles = (uint32_t)((n + RPX_TILE - 1) / RPX_TILE);
    const size_t state_words = rpx_state_words(n_tiles);
    cudaError_t e;
    if (state_words > ctx->tile_state_cap) {
        if (ctx->tile_state) cudaFree(ctx->tile_state);
        ctx->tile_state = nullptr;
        ctx->tile_state_cap = 0;
        if ((e = cudaMalloc(&ctx->tile_state, state_words * 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc(tile state)");
        ctx->tile_state_cap = state_words * 2;
    }
    const size_t nfc = (size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1);
    std::vector<uint32_t> fc(nfc, 0u);
    if ((e = cudaMemsetAsync(ctx->d_face_counts, 0, sizeof(uint32_t) * nfc, st)) != cudaSuccess ||
        (e = cudaMemsetAsync(ctx->tile_state, 0, state_words * sizeof(unsigned long long), st)) != cudaSuccess ||
        (e = cudaMemsetAsync(ctx->tile_counter, 0, sizeof(uint32_t), st)) != cudaSuccess ||
        (e = launch_intersect(ctx->face_class, st, n_tiles, ctx->scene_smem, ctx->ds, rays->soa, ml, -1)) != cudaSuccess)
        return bail(e, "k_intersect");
    ShadeArgs sa;
    sa.S = ctx->ds;
    sa.in = rays->soa;
    sa.out = child->soa;
    sa.max_length = ml;
    sa.tile_state = ctx->tile_state;
    sa.tile_counter = ctx->tile_counter;
    sa.d_count = ctx->d_count;
    sa.face_counts = ctx->d_face_counts;
    sa.n_tiles = n_tiles;
    sa.smem_bytes = ctx->scene_smem;
    sa.ahead_face = -2;  // children stay un-intersected
    sa.n_dev = nullptr;
    sa.h_count = nullptr;
    if ((e = shade_launcher(is_g, ctx->face_class, ctx->mm_idx, ctx->scene_smem > 0)(st, sa)) != cudaSuccess ||
        (e = cudaMemcpyAsync(ctx->h_count, ctx->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(fc.data(), ctx->d_face_counts, sizeof(uint32_t) * (size_t)ctx->n_traced, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return bail(e, "k_shade");
    child->soa.n = *ctx->h_count;
    if (face_counts)
        for (int i = 0; i < ctx->n_traced; i++) face_counts[i] += fc[(size_t)i];
    *out_children = child;
    return RPX_OK;
}

// ------------------------------------------------------------------ capture planes
extern "C" const rpx_rays* rpx_result_rays(const rpx_result* res, int g) {
    if (!res || g < 0 || g >= (int)res->gens.size()) return nullptr;
    return res->gens[(size_t)g];
}

extern "C" int rpx_capture(rpx_ctx* ctx, const rpx_rays* const* gens, int n_gens, const uint32_t* wl_offsets,
                           const uint32_t* wl_map, uint32_t n_wl_map, rpx_rays** out, uint64_t* counts) {
    if (!ctx || !gens || !out || n_gens <= 0) return fail(ctx, RPX_ERR_INVALID, "NULL / empty argument");
    *out = nullptr;
    if (!ctx->have_capture) return fail(ctx, RPX_ERR_STATE, "rpx_capture_scene_set has not been called");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long total = 0, tiles_total = 0;
    for (int j = 0; j < n_gens; j++) {
        if (!gens[j]) return fail(ctx, RPX_ERR_INVALID, "collection %d is NULL (dropped generation?)", j);
        if (gens[j]->is_gausslet != gens[0]->is_gausslet)
            return fail(ctx, RPX_ERR_INVALID, "collections mix rays and gausslets");
        total += gens[j]->soa.n;
        tiles_total += rpx_state_words((gens[j]->soa.n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);  // grouped look-back state
    }
    const int is_g = gens[0]->is_gausslet;
    rpx_rays* dst = nullptr;
    int rc = rpx_rays_alloc(ctx, total, is_g, &dst);
    if (rc != RPX_OK) return rc;
    // scratch: [totals (n_gens + 1) u64][tile state u64 x tiles][ticket counters u32 x n_gens][wl_map u32]
    const size_t off_state = sizeof(unsigned long long) * (size_t)(n_gens + 1);
    const size_t off_cnt = off_state + sizeof(unsigned long long) * (size_t)tiles_total;
    const size_t off_map = align_up(off_cnt + sizeof(uint32_t) * (size_t)n_gens, 8);
    const size_t bytes = off_map + sizeof(uint32_t) * (size_t)(wl_map ? n_wl_map : 0) + 8;
    unsigned char* scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&scratch, bytes, ctx->stream);
    if (e != cudaSuccess) {
        rpx_rays_free(ctx, dst);
        return fail(ctx, RPX_ERR_NOMEM, "capture scratch (%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    auto bail = [&](cudaError_t err, const char* what) {
        cudaFreeAsync(scratch, ctx->stream);
        rpx_rays_free(ctx, dst);
        return fail(ctx, RPX_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
    };
    if ((e = cudaMemsetAsync(scratch, 0, off_map, ctx->stream)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    unsigned long long* d_totals = (unsigned long long*)scratch;
    unsigned long long* d_state = (unsigned long long*)(scratch + off_state);
    uint32_t* d_cnt = (uint32_t*)(scratch + off_cnt);
    uint32_t* d_map = nullptr;
    if (wl_map && n_wl_map) {
        d_map = (uint32_t*)(scratch + off_map);
        if ((e = cudaMemcpyAsync(d_map, wl_map, sizeof(uint32_t) * n_wl_map, cudaMemcpyHostToDevice, ctx->stream)) !=
            cudaSuccess)
            return bail(e, "cudaMemcpyAsync(wl_map)");
    }
    unsigned long long tile_off = 0;
    for (int j = 0; j < n_gens; j++) {
        const unsigned long long n = gens[j]->soa.n;
        if (n == 0) {  // nothing to filter: carry the running total forward
            if ((e = cudaMemcpyAsync(d_totals + j + 1, d_totals + j, sizeof(unsigned long long),
                                     cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess)
                return bail(e, "cudaMemcpyAsync(total)");
            continue;
        }
        const unsigned n_tiles = (unsigned)((n + RPX_FILTER_TILE - 1) / RPX_FILTER_TILE);
        CaptureArgs a;
        a.S = ctx->cap_ds;
        a.in = gens[j]->soa;
        a.out = dst->soa;
        a.tile_state = d_state + tile_off;
        a.tile_counter = d_cnt + j;
        a.d_base = d_totals + j;
        a.d_next = d_totals + j + 1;
        a.wl_offset = wl_offsets ? wl_offsets[j] : 0u;
        a.wl_map = d_map;
        a.face_ids = ctx->cap_face_ids;
        a.smem_bytes = ctx->cap_smem;
        if ((e = launch_capture(is_g, ctx->cap_face_class, ctx->stream, n_tiles, a)) != cudaSuccess)
            return bail(e, "k_capture launch");
        tile_off += rpx_state_words(n_tiles);
    }
    std::vector<unsigned long long> h_totals((size_t)n_gens + 1);
    if ((e = cudaMemcpyAsync(h_totals.data(), d_totals, sizeof(unsigned long long) * h_totals.size(),
                             cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        return bail(e, "capture");
    cudaFreeAsync(scratch, ctx->stream);
    dst->soa.n = h_totals[(size_t)n_gens];
    if (counts)
        for (int j = 0; j < n_gens; j++) counts[j] = h_totals[(size_t)j + 1] - h_totals[(size_t)j];
    *out = dst;
    return RPX_OK;
}

// ------------------------------------------------------------------ streamed trace
// rpx_trace over a host-resident source that is cut into contiguous chunks: the upload of chunk
// c+1 (stream_in) and the download of chunk c's generations (stream_out) overlap the tracing of
// chunk c and each other, so PCIe runs full duplex and the device never holds more than two
// chunks.  Same results as one rpx_trace call: generation g is the concatenation of the chunks'
// generation g in source order (children are emitted in parent order) with parent_idx shifted
// by the number of generation g-1 rays of earlier chunks -- the single-node form of the multi-GPU
// sharding of SURVEY 8e.  It is also how a source larger than HBM is traced.
static void launch_soa_to_aos(rpx_ctx* ctx, const rpx_rays* rays, void* d_aos, uint32_t parent_offset) {
    const uint64_t n = rays->soa.n;
    if (rays->is_gausslet) {
        const int T = 64;
        k_soa_to_aos<RPX_WORDS_GAUSSLET, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, ctx->stream>>>(
            rays->soa, (uint32_t*)d_aos, parent_offset);
    } else {
        const int T = 256;
        k_soa_to_aos<RPX_WORDS_RAY, T><<<(unsigned)((n + T - 1) / T), T, T * RPX_RAY_BYTES, ctx->stream>>>(
            rays->soa, (uint32_t*)d_aos, parent_offset);
    }
}

extern "C" int rpx_trace_streamed(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                                  int recursion_limit, uint64_t chunk_rays, void* const* out_gens,
                                  const uint64_t* out_capacity, int max_gens, uint64_t* out_counts, int* n_gens,
                                  uint32_t* face_counts, double* device_ms) {
    if (!ctx || !out_gens || !out_capacity || !out_counts || !n_gens || max_gens <= 0 || (!rays_aos && n))
        return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called before tracing");
    CU(ctx, cudaSetDevice(ctx->device));
    if (!ctx->have_copy_streams) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
        CU(ctx, cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) CU(ctx, cudaEventCreateWithFlags(&ctx->st_out_done[k], cudaEventDisableTiming));
        ctx->have_copy_streams = true;
    }
    if (chunk_rays == 0) chunk_rays = 262144;
    const size_t rec = is_gausslet ? RPX_GAUSSLET_BYTES : RPX_RAY_BYTES;
    const uint64_t n_chunks = n ? (n + chunk_rays - 1) / chunk_rays : 0;
    for (int g = 0; g < max_gens; g++) out_counts[g] = 0;
    *n_gens = 0;
    std::vector<uint64_t> totals((size_t)max_gens, 0);
    std::vector<uint32_t> fc((size_t)(ctx->n_traced > 0 ? ctx->n_traced : 1), 0);
    double ms = 0.0;
    void* d_in[2] = {nullptr, nullptr};
    int ring = 0;  // next download staging buffer
    cudaEvent_t ev_in[2], ev_used[2], ev_ready;
    bool used_once[2] = {false, false};
    int rc = RPX_OK;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2; k++) {
        cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_used[k], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming);
    // out_gens[0] may alias rays_aos (the reference's convention: traced_rays[0] IS input_rays, mutated in
    // place, ctracer.pyx:2086-2087 / 1900-1903): chunk c's generation 0 lands on the records chunk c was
    // uploaded from, after that upload has completed.  (A 12-byte write-back + host-side scatter instead of
    // the 188-byte record was measured SLOWER on B200: 1.69e8 vs 2.51e8 seg/s end to end, the scatter loop
    // on the calling thread being the bottleneck -- profiles/r02_notes.md; removed.)
    const uint64_t in_bytes = (n < chunk_rays ? n : chunk_rays) * rec;
    if (n_chunks && rc == RPX_OK) {
        for (int k = 0; k < (n_chunks > 1 ? 2 : 1) && e == cudaSuccess; k++) {
            if (ctx->st_in_bytes[k] < in_bytes) {  // grow once; kept for later calls
                if (ctx->st_in[k]) cudaFree(ctx->st_in[k]);
                ctx->st_in[k] = nullptr;
                ctx->st_in_bytes[k] = 0;
                if ((e = cudaMalloc(&ctx->st_in[k], in_bytes)) == cudaSuccess) ctx->st_in_bytes[k] = in_bytes;
            }
            d_in[k] = ctx->st_in[k];
        }
        if (e != cudaSuccess) rc = fail(ctx, RPX_ERR_NOMEM, "chunk staging (%llu bytes): %s", (unsigned long long)in_bytes, cudaGetErrorString(e));
    }
    auto issue_upload = [&](uint64_t c) -> cudaError_t {
        const int b = (int)(c & 1);
        const uint64_t lo = c * chunk_rays, cnt = (n - lo < chunk_rays) ? n - lo : chunk_rays;
        cudaError_t err = cudaSuccess;
        if (used_once[b]) err = cudaStreamWaitEvent(ctx->stream_in, ev_used[b], 0);  // buffer b consumed by chunk c-2
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(d_in[b], (const unsigned char*)rays_aos + lo * rec, cnt * rec, cudaMemcpyHostToDevice,
                                  ctx->stream_in);
        if (err == cudaSuccess) err = cudaEventRecord(ev_in[b], ctx->stream_in);
        return err;
    };
    if (rc == RPX_OK && n_chunks && (e = issue_upload(0)) != cudaSuccess)
        rc = fail(ctx, RPX_ERR_CUDA, "chunk upload: %s", cudaGetErrorString(e));
    for (uint64_t c = 0; c < n_chunks && rc == RPX_OK; c++) {
        const int b = (int)(c & 1);
        const uint64_t lo = c * chunk_rays, cnt = (n - lo < chunk_rays) ? n - lo : chunk_rays;
        if (c + 1 < n_chunks && (e = issue_upload(c + 1)) != cudaSuccess) {
            rc = fail(ctx, RPX_ERR_CUDA, "chunk upload: %s", cudaGetErrorString(e));
            break;
        }
        // AoS -> SoA on the compute stream once the chunk has landed
        rpx_rays* r = nullptr;
        if ((rc = rpx_rays_alloc(ctx, cnt, is_gausslet, &r)) != RPX_OK) break;
        r->soa.n = cnt;
        cudaStreamWaitEvent(ctx->stream, ev_in[b], 0);
        if (is_gausslet) {
            const int T = 64;
            k_aos_to_soa<RPX_WORDS_GAUSSLET, T><<<(unsigned)((cnt + T - 1) / T), T, T * RPX_GAUSSLET_BYTES, ctx->stream>>>(
                (const uint32_t*)d_in[b], r->soa);
        } else {
            const int T = 256;
            k_aos_to_soa<RPX_WORDS_RAY, T><<<(unsigned)((cnt + T - 1) / T), T, T * RPX_RAY_BYTES, ctx->stream>>>(
                (const uint32_t*)d_in[b], r->soa);
        }
        cudaEventRecord(ev_used[b], ctx->stream);
        used_once[b] = true;
        rpx_result* res = nullptr;
        rc = rpx_trace_device(ctx, r, max_length, recursion_limit, RPX_TRACE_DEFAULT, &res);  // owns r
        if (rc != RPX_OK) break;
        ms += res->device_ms;
        for (size_t i = 0; i < res->face_counts.size() && i < fc.size(); i++) fc[i] += res->face_counts[i];
        const int ng = (int)res->gens.size();
        if (ng > max_gens) {
            rc = fail(ctx, RPX_ERR_INVALID, "trace produced %d generations, the caller provided %d output buffers", ng, max_gens);
            rpx_result_free(ctx, res);
            break;
        }
        for (int g = 0; g < ng && rc == RPX_OK; g++) {
            const rpx_rays* gen = res->gens[(size_t)g];
            const uint64_t m = gen ? gen->soa.n : 0;
            if (!m) continue;
            if (totals[(size_t)g] + m > out_capacity[g]) {
                rc = fail(ctx, RPX_ERR_NOMEM, "output buffer of generation %d holds %llu records, needs more than %llu", g,
                          (unsigned long long)out_capacity[g], (unsigned long long)(totals[(size_t)g] + m));
                break;
            }
            // global parent index = local + rays of generation g-1 in earlier chunks (generation 0 keeps its own);
            // ray_t.parent_idx is 32 bits wide: a generation beyond that cannot be numbered
            if (totals[(size_t)g] + m >= 0xFFFFFFFFull) {
                rc = fail(ctx, RPX_ERR_INVALID, "generation %d would exceed the 32-bit parent_idx of ray_t (%llu rays)", g,
                          (unsigned long long)(totals[(size_t)g] + m));
                break;
            }
            const uint32_t poff = g > 0 ? (uint32_t)(totals[(size_t)g - 1]) : 0u;
            // next buffer of the download ring: wait (on the compute stream) until its previous
            // D2H has left it, grow it if this generation is larger than anything seen so far
            const int k = ring;
            ring = (ring + 1) & 3;
            if (ctx->st_out_busy[k]) cudaStreamWaitEvent(ctx->stream, ctx->st_out_done[k], 0);
            if (ctx->st_out_bytes[k] < m * rec) {
                if (ctx->st_out_busy[k]) cudaEventSynchronize(ctx->st_out_done[k]);
                if (ctx->st_out[k]) cudaFree(ctx->st_out[k]);
                ctx->st_out[k] = nullptr;
                ctx->st_out_bytes[k] = 0;
                const size_t want = (size_t)(m * rec) + (size_t)(m * rec) / 4;  // head room: later chunks vary
                if ((e = cudaMalloc(&ctx->st_out[k], want)) != cudaSuccess) {
                    rc = fail(ctx, RPX_ERR_NOMEM, "download staging (%zu bytes): %s", want, cudaGetErrorString(e));
                    break;
                }
                ctx->st_out_bytes[k] = want;
            }
            void* d_stage = ctx->st_out[k];
            launch_soa_to_aos(ctx, gen, d_stage, poff);
            cudaEventRecord(ev_ready, ctx->stream);
            cudaStreamWaitEvent(ctx->stream_out, ev_ready, 0);
            e = cudaMemcpyAsync((unsigned char*)out_gens[g] + totals[(size_t)g] * rec, d_stage, m * rec,
                                cudaMemcpyDeviceToHost, ctx->stream_out);
            cudaEventRecord(ctx->st_out_done[k], ctx->stream_out);
            ctx->st_out_busy[k] = true;
            if (e != cudaSuccess) rc = fail(ctx, RPX_ERR_CUDA, "generation download: %s", cudaGetErrorString(e));
        }
        // offsets of the NEXT chunk are the totals including this one; update after all generations
        // of this chunk used the old values for their parent offsets
        if (rc == RPX_OK) {
            std::vector<uint64_t> add((size_t)ng, 0);
            for (int g = 0; g < ng; g++) add[(size_t)g] = res->gens[(size_t)g] ? res->gens[(size_t)g]->soa.n : 0;
            for (int g = 0; g < ng; g++) totals[(size_t)g] += add[(size_t)g];
            if (ng > *n_gens) *n_gens = ng;
        }
        rpx_result_free(ctx, res);  // generation buffers go back to the pool in compute-stream order
    }
    cudaStreamSynchronize(ctx->stream_out);
    cudaStreamSynchronize(ctx->stream_in);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 4; k++) ctx->st_out_busy[k] = false;  // everything drained above
    for (int k = 0; k < 2; k++) {
        cudaEventDestroy(ev_in[k]);
        cudaEventDestroy(ev_used[k]);
    }
    cudaEventDestroy(ev_ready);
    if (rc != RPX_OK) return rc;
    for (int g = 0; g < *n_gens; g++) out_counts[g] = totals[(size_t)g];
    if (face_counts)
        for (int i = 0; i < ctx->n_traced; i++) face_counts[i] = fc[(size_t)i];
    if (device_ms) *device_ms = ms;
    return RPX_OK;
}

extern "C" int rpx_result_n_generations(const rpx_result* res) { return res ? (int)res->counts.size() : 0; }

extern "C" int rpx_result_counts(const rpx_result* res, uint64_t* counts) {
    if (!res || !counts) return RPX_ERR_INVALID;
    for (size_t i = 0; i < res->counts.size(); i++) counts[i] = res->counts[i];
    return RPX_OK;
}

extern "C" int rpx_result_generation(rpx_ctx* ctx, const rpx_result* res, int g, void* out_aos, uint64_t capacity) {
    if (!ctx || !res) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (g < 0 || g >= (int)res->gens.size()) return fail(ctx, RPX_ERR_INVALID, "generation %d out of range", g);
    if (!res->gens[g]) return fail(ctx, RPX_ERR_STATE, "generation %d was dropped (RPX_TRACE_KEEP_LAST_ONLY)", g);
    return rpx_rays_download(ctx, res->gens[g], out_aos, capacity);
}

extern "C" int rpx_result_face_counts(const rpx_result* res, uint32_t* counts) {
    if (!res || !counts) return RPX_ERR_INVALID;
    for (size_t i = 0; i < res->face_counts.size(); i++) counts[i] = res->face_counts[i];
    return RPX_OK;
}

extern "C" double rpx_result_device_ms(const rpx_result* res) { return res ? res->device_ms : 0.0; }
extern "C" uint64_t rpx_result_launches(const rpx_result* res) { return res ? res->launches : 0; }

extern "C" int rpx_result_kernel_ms(const rpx_result* res, int which, double* total_ms, uint64_t* launches) {
    if (!res || which < 0 || which > 1) return RPX_ERR_INVALID;
    if (total_ms) *total_ms = res->k_ms[which];
    if (launches) *launches = res->k_launches[which];
    return RPX_OK;
}

// ------------------------------------------------------------------ unit entry points
// Host buffers in / out; each call stages through temporary device buffers.
namespace {
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, st); }
    cudaError_t put(const void* src, size_t bytes) {
        cudaError_t e = alloc(bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st);
    }
    cudaError_t get(void* dst, size_t bytes) { return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, st); }
};
}  // namespace

static int unit_precheck(rpx_ctx* ctx) {
    if (!ctx) return RPX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, RPX_ERR_STATE, "rpx_scene_set must be called first");
    CU(ctx, cudaSetDevice(ctx->device));
    return RPX_OK;
}

extern "C" int rpx_unit_face_intersect(rpx_ctx* ctx, int face, const double* p1, const double* p2, uint64_t n,
                                       int is_base_ray, double* out_dist) {
    int rc = unit_precheck(ctx);
    if (rc != RPX_OK) return rc;
    if (face < 0 || face >= ctx->ds.n_faces) return fail(ctx, RPX_ERR_INVALID, "face %d out of range", face);
    if (!n) return RPX_OK;
    cudaStream_t st = ctx->stream;
    DevBuf a(st), b(st), o(st);
    CU(ctx, a.put(p1, n * 24));
    CU(ctx, b.put(p2, n * 24));
    CU(ctx, o.alloc(n * 8));
    CU(ctx, launch_unit_face_intersect(st, ctx->ds, face, (const double*)a.p, (const double*)b.p, n, is_base_ray,
                                       (double*)o.p));
    CU(ctx, o.get(out_dist, n * 8));
    CU(ctx, cudaStreamSynchronize(st));
    return RPX_OK;
}

extern "C" int rpx_unit_face_normal(rpx_ctx* ctx, int face, const double* points, uint64_t n, double* out_normal,
                                    double* out_tangent) {
    int rc = unit_precheck(ctx);
    if (rc != RPX_OK) return rc;
    if (face < 0 || face >= ctx->ds.n_faces) return fail(ctx, RPX_ERR_INVALID, "face %d out of range", face);
    if (!n) return RPX_OK;
    cudaStream_t st = ctx->stream;
    DevBuf a(st), nn(st), tt(st);
    CU(ctx, a.put(points, n * 24));
    CU(ctx, nn.alloc(n * 24));
    CU(ctx, tt.alloc(n * 24));
    CU(ctx, launch_unit_face_normal(st, ctx->ds, face, (const double*)a.p, n, (double*)nn.p, (double*)tt.p));
    CU(ctx, nn.get(out_normal, n * 24));
    CU(ctx, tt.get(out_tangent, n * 24));
    CU(ctx, cudaStreamSynchronize(st));
    return RPX_OK;
}

extern "C" int rpx_unit_material_eval(rpx_ctx* ctx, int material, const void* rays_aos, uint64_t n, const double* point,
                                      const double* normal, const double* tangent, void* out_aos_2n,
                                      uint32_t* out_counts) {
    int rc = unit_precheck(ctx);
    if (rc != RPX_OK) return rc;
    if (material < 0 || material >= ctx->ds.n_mats)
        return fail(ctx, RPX_ERR_INVALID, "material %d out of range", material);
    if (!n) return RPX_OK;
    cudaStream_t st = ctx->stream;
    DevBuf r(st), p(st), nn(st), tt(st), o(st), c(st);
    CU(ctx, r.put(rays_aos, n * RPX_RAY_BYTES));
    CU(ctx, p.put(point, n * 24));
    CU(ctx, nn.put(normal, n * 24));
    CU(ctx, tt.put(tangent, n * 24));
    CU(ctx, o.alloc(2 * n * RPX_RAY_BYTES));
    CU(ctx, cudaMemsetAsync(o.p, 0, 2 * n * RPX_RAY_BYTES, st));
    CU(ctx, c.alloc(n * 4));
    CU(ctx, launch_unit_material_eval(st, ctx->ds, material, (const uint32_t*)r.p, n, (const double*)p.p,
                                      (const double*)nn.p, (const double*)tt.p, (uint32_t*)o.p, (uint32_t*)c.p));
    CU(ctx, o.get(out_aos_2n, 2 * n * RPX_RAY_BYTES));
    CU(ctx, c.get(out_counts, n * 4));
    CU(ctx, cudaStreamSynchronize(st));
    return RPX_OK;
}

extern "C" int rpx_unit_distortion(rpx_ctx* ctx, int distortion, const double* x, const double* y, uint64_t n,
                                   double* out_z, double* out_grad) {
    int rc = unit_precheck(ctx);
    if (rc != RPX_OK) return rc;
    if (distortion < 0 || distortion >= ctx->ds.n_dists)
        return fail(ctx, RPX_ERR_INVALID, "distortion %d out of range", distortion);
    if (!n) return RPX_OK;
    cudaStream_t st = ctx->stream;
    DevBuf a(st), b(st), z(st), g(st);
    CU(ctx, a.put(x, n * 8));
    CU(ctx, b.put(y, n * 8));
    CU(ctx, z.alloc(n * 8));
    CU(ctx, g.alloc(n * 24));
    CU(ctx, launch_unit_distortion(st, ctx->ds, distortion, (const double*)a.p, (const double*)b.p, n, (double*)z.p,
                                   (double*)g.p));
    CU(ctx, z.get(out_z, n * 8));
    CU(ctx, g.get(out_grad, n * 24));
    CU(ctx, cudaStreamSynchronize(st));
    return RPX_OK;
}
