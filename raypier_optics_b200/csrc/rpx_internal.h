// rpx_internal.h -- host-side structures shared by the translation units that implement the
// C ABI (rpx_api.cu: context / scene / trace / capture; rpx_field.cu: E-field summation).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rpx.h"
#include "rpx_kernels.cuh"

#define RPX_MAX_PIPE_GENS 1024  /* generations a pipelined trace can hold counts for */
#define RPX_PIPE_STATE_TILES (4u << 20) /* 32 MB of look-back state: 5e8 parents per trace without a re-zero */

struct rpx_rays {
    rpx::Soa soa;
    void* block;  // single allocation backing soa.f / soa.u / soa.p
    size_t bytes;
    int is_gausslet;
};

struct KernelStat {
    std::vector<cudaEvent_t> start, stop;
};

struct rpx_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t stream_in, stream_out;  // copy streams of rpx_trace_streamed (created on first use)
    bool have_copy_streams;
    // persistent staging of rpx_trace_streamed: 2 upload buffers, a ring of download buffers
    void* st_in[2];
    size_t st_in_bytes[2];
    void* st_out[4];
    size_t st_out_bytes[4];
    cudaEvent_t st_out_done[4];  // D2H out of st_out[k] finished (recorded on stream_out)
    bool st_out_busy[4];
    std::string err;
    // scene
    bool have_scene;
    rpx::DevScene ds;
    void* scene_block;
    int n_traced;
    int max_kids;       // upper bound of children per hit over all materials in the scene
    int scene_smem;     // bytes of shared memory the staged scene needs (0 = use global)
    int face_class;     // RPX_FC_SIMPLE / RPX_FC_FULL / RPX_FC_MESH kernel variant for this scene
    int mm_idx;         // material-mask kernel variant: 0 LIGHT, 1 COATED, 2 FULLDIEL, 3 ALL
    // capture-plane scene (rpx_capture_scene_set): a second, independent face list
    bool have_capture;
    rpx::DevScene cap_ds;
    void* cap_block;
    uint32_t* cap_face_ids;  // device copy of the Python-side Face.idx values, or NULL
    int cap_smem;
    int cap_face_class;
    // scratch
    unsigned long long* tile_state;
    size_t tile_state_cap;  // tiles
    uint32_t* tile_counter;
    unsigned long long* d_count;
    unsigned long long* h_count;  // pinned
    uint32_t* d_face_counts;
    unsigned long long* d_counts;  // per-generation counts of a pipelined trace (RPX_MAX_PIPE_GENS)
    unsigned long long* h_counts;  // pinned + mapped: the kernels write len(new_rays) straight into it
    unsigned long long* h_counts_dev;  // device alias of h_counts
    unsigned long long* pipe_state;  // tile-state slices of a pipelined trace, zeroed ahead of use
    size_t pipe_state_cap;           // tiles
    uint32_t* pipe_counters;         // one ticket counter per generation
    uint32_t* pipe_hits;             // pipe_hits[g] != 0: some ray of generation g hit a face (set by the launch that built it)
    uint32_t* pipe_miss;             // pipe_miss[g] = rays of generation g that hit nothing (drives the tile-local compaction)
    // event pool
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used;
};

struct rpx_result {
    std::vector<rpx_rays*> gens;  // nullptr for generations dropped in KEEP_LAST_ONLY mode
    std::vector<uint64_t> counts;
    std::vector<uint32_t> face_counts;
    double device_ms;
    uint64_t launches;
    double k_ms[2];
    uint64_t k_launches[2];
};

// records the message for rpx_last_error() and returns `code`
int rpx_fail(rpx_ctx* ctx, int code, const char* fmt, ...);
#define fail rpx_fail

#define CU(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)


// a generation buffer from the stream-ordered pool (rpx_api.cu)
int rpx_rays_alloc(rpx_ctx* ctx, unsigned long long cap_req, int is_gausslet, rpx_rays** out);
