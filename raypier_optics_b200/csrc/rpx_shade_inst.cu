// rpx_shade_inst.cu -- instantiates ONE k_shade<RPX_I_GAUSS, RPX_I_FC, MM[RPX_I_MM]> variant.
// Compiled once per (gausslet, face class, material mask) triple (see Makefile) so all 16
// variants build in parallel.
#include <mutex>

#include "rpx_launch.h"

#if !defined(RPX_I_GAUSS) || !defined(RPX_I_FC) || !defined(RPX_I_MM)
#error "compile with -DRPX_I_GAUSS=0|1 -DRPX_I_FC=0|1|2 -DRPX_I_MM=0..3"
#endif

// RPX_I_SS=0: the instantiation for scenes whose tables do not fit in shared memory (suffix _nss)
#ifndef RPX_I_SS
#define RPX_I_SS 1
#endif
#if RPX_I_SS
#define RPX_CAT_(a, b, c) launch_shade_g##a##_f##b##_m##c
#else
#define RPX_CAT_(a, b, c) launch_shade_g##a##_f##b##_m##c##_nss
#endif
#define RPX_CAT(a, b, c) RPX_CAT_(a, b, c)

namespace rpx {

#if RPX_I_MM == 0
static constexpr uint32_t kMask = RPX_MM_LIGHT;
#elif RPX_I_MM == 1
static constexpr uint32_t kMask = RPX_MM_COATED;
#elif RPX_I_MM == 2
static constexpr uint32_t kMask = RPX_MM_FULLDIEL;
#else
static constexpr uint32_t kMask = RPX_MM_ALL;
#endif

cudaError_t RPX_CAT(RPX_I_GAUSS, RPX_I_FC, RPX_I_MM)(cudaStream_t st, const ShadeArgs& a) {
    // dynamic shared memory = child staging (35 KB lean, 47 KB full) + the scene copy: may need the > 48 KB opt-in.
    // Persistent grid: one wave of resident CTAs (SMs x occupancy), never more than the tiles.
    // per-device state (cudaFuncSetAttribute and the occupancy answer are per device; one process may
    // open several GPUs through rpx_init): indexed by the current device, guarded by a mutex
    struct DevState {
        int resident_ctas = 0;
        int resident_dyn = -1;  // recomputed when another scene changes the shared-memory footprint
        bool attr_set = false;
    };
    static DevState states[RPX_MAX_DEVICES];
    static std::mutex mu;
    const int dyn = RPX_STAGE_BYTES + (RPX_I_SS ? a.smem_bytes : 0);
    auto kern = k_shade<(RPX_I_GAUSS != 0), RPX_I_FC, kMask, (RPX_I_SS != 0)>;
    int dev = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= RPX_MAX_DEVICES) return cudaErrorInvalidDevice;
    int resident_ctas;
    {
        std::lock_guard<std::mutex> lock(mu);
        DevState& ds = states[dev];
        if (!ds.attr_set) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     RPX_STAGE_BYTES + 40 * 1024);
            if (e != cudaSuccess) return e;
            ds.attr_set = true;
        }
        if (ds.resident_dyn != dyn) {
            int sms = 0, per_sm = 0;
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RPX_TILE, dyn)) != cudaSuccess) return e;
            ds.resident_ctas = sms * (per_sm < 1 ? 1 : per_sm);
            ds.resident_dyn = dyn;
        }
        resident_ctas = ds.resident_ctas;
    }
    const unsigned grid = a.n_tiles < (unsigned)resident_ctas ? a.n_tiles : (unsigned)resident_ctas;
    kern<<<grid, RPX_TILE, dyn, st>>>(a.S, a.in, a.out, a.max_length, a.tile_state, a.tile_counter, a.d_count,
                                      a.face_counts, a.n_tiles, a.ahead_face, a.n_dev, a.h_count, a.hits_in, a.hits_out, a.miss_in,
                                      a.miss_out);
    return cudaGetLastError();
}

}  // namespace rpx
