"""ctypes binding of librpx.so (the C ABI declared in include/rpx.h).

There is deliberately no fallback: if the CUDA library has not been built, or there is
no CUDA device, every entry point raises.  (The plain-C oracle under oracle/ is test
infrastructure and is never imported from here.)
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
# RPX_LIB: developer override for A/B-testing a differently built librpx (never a fallback)
LIB_PATH = os.environ.get("RPX_LIB") or os.path.join(_HERE, "csrc", "librpx.so")

# every symbol include/rpx.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "rpx_init", "rpx_shutdown", "rpx_last_error", "rpx_abi_version", "rpx_scene_set",
    "rpx_host_alloc", "rpx_host_free", "rpx_rays_upload", "rpx_rays_download", "rpx_rays_count",
    "rpx_rays_free", "rpx_rays_clone", "rpx_trace_device", "rpx_trace", "rpx_trace_sequence", "rpx_result_n_generations",
    "rpx_result_counts", "rpx_result_generation", "rpx_result_face_counts",
    "rpx_result_device_ms", "rpx_result_launches", "rpx_result_kernel_ms", "rpx_result_free",
    "rpx_stream", "rpx_unit_face_intersect", "rpx_unit_face_normal", "rpx_unit_material_eval",
    "rpx_unit_distortion", "rpx_capture_scene_set", "rpx_result_rays", "rpx_capture",
    "rpx_field_prepare", "rpx_field_count", "rpx_field_modes", "rpx_field_evaluate",
    "rpx_field_evaluate_device", "rpx_field_last_ms", "rpx_field_free", "rpx_trace_streamed",
    "rpx_rays_project_to_sphere", "rpx_field_prepare_neighbours", "rpx_unit_evaluate_modes",
    "rpx_select_terminal", "rpx_rays_export_device", "rpx_rays_import_device",
    "rpx_detector_create", "rpx_detector_reset", "rpx_detector_accumulate", "rpx_detector_read",
    "rpx_detector_field_device", "rpx_detector_npoints", "rpx_detector_modes", "rpx_detector_ms",
    "rpx_detector_free", "rpx_trace_consume", "rpx_trace_step",
)


class RpxError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "%s (%d): %s" % (A.STATUS_NAMES.get(code, "RPX_ERR"), code, message))
        self.code = code


_LIB = None


def load():
    """Load librpx.so and declare prototypes.  Raises if the library is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "librpx.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C raypier_optics_b200/csrc`.  raypier_optics_b200 has no CPU "
            "fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, d = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    pvp = C.POINTER(C.c_void_p)
    L.rpx_init.argtypes = [i32, pvp]
    L.rpx_init.restype = i32
    L.rpx_shutdown.argtypes = [vp]
    L.rpx_shutdown.restype = None
    L.rpx_last_error.argtypes = [vp]
    L.rpx_last_error.restype = C.c_char_p
    L.rpx_abi_version.restype = i32
    L.rpx_scene_set.argtypes = [vp, vp]
    L.rpx_scene_set.restype = i32
    L.rpx_host_alloc.argtypes = [C.c_size_t]
    L.rpx_host_alloc.restype = vp
    L.rpx_host_free.argtypes = [vp]
    L.rpx_host_free.restype = None
    L.rpx_rays_upload.argtypes = [vp, vp, u64, i32, pvp]
    L.rpx_rays_upload.restype = i32
    L.rpx_rays_download.argtypes = [vp, vp, vp, u64]
    L.rpx_rays_download.restype = i32
    L.rpx_rays_count.argtypes = [vp]
    L.rpx_rays_count.restype = u64
    L.rpx_rays_clone.argtypes = [vp, vp, pvp]
    L.rpx_rays_clone.restype = i32
    L.rpx_rays_free.argtypes = [vp, vp]
    L.rpx_rays_free.restype = None
    L.rpx_trace_device.argtypes = [vp, vp, d, i32, u32, pvp]
    L.rpx_trace_device.restype = i32
    L.rpx_trace.argtypes = [vp, vp, u64, i32, d, i32, u32, pvp]
    L.rpx_trace.restype = i32
    L.rpx_trace_sequence.argtypes = [vp, vp, u64, i32, d, i32, vp, i32, pvp]
    L.rpx_trace_sequence.restype = i32
    L.rpx_result_n_generations.argtypes = [vp]
    L.rpx_result_n_generations.restype = i32
    L.rpx_result_counts.argtypes = [vp, vp]
    L.rpx_result_counts.restype = i32
    L.rpx_result_generation.argtypes = [vp, vp, i32, vp, u64]
    L.rpx_result_generation.restype = i32
    L.rpx_result_face_counts.argtypes = [vp, vp]
    L.rpx_result_face_counts.restype = i32
    L.rpx_result_device_ms.argtypes = [vp]
    L.rpx_result_device_ms.restype = d
    L.rpx_result_launches.argtypes = [vp]
    L.rpx_result_launches.restype = u64
    L.rpx_result_kernel_ms.argtypes = [vp, i32, C.POINTER(d), C.POINTER(u64)]
    L.rpx_result_kernel_ms.restype = i32
    L.rpx_result_free.argtypes = [vp, vp]
    L.rpx_result_free.restype = None
    L.rpx_stream.argtypes = [vp]
    L.rpx_stream.restype = vp
    L.rpx_unit_face_intersect.argtypes = [vp, i32, vp, vp, u64, i32, vp]
    L.rpx_unit_face_intersect.restype = i32
    L.rpx_unit_face_normal.argtypes = [vp, i32, vp, u64, vp, vp]
    L.rpx_unit_face_normal.restype = i32
    L.rpx_unit_material_eval.argtypes = [vp, i32, vp, u64, vp, vp, vp, vp, vp]
    L.rpx_unit_material_eval.restype = i32
    L.rpx_unit_distortion.argtypes = [vp, i32, vp, vp, u64, vp, vp]
    L.rpx_unit_distortion.restype = i32
    L.rpx_capture_scene_set.argtypes = [vp, vp, vp]
    L.rpx_capture_scene_set.restype = i32
    L.rpx_result_rays.argtypes = [vp, i32]
    L.rpx_result_rays.restype = vp
    L.rpx_capture.argtypes = [vp, vp, i32, vp, vp, u32, pvp, vp]
    L.rpx_capture.restype = i32
    L.rpx_field_prepare.argtypes = [vp, vp, vp, vp, i32, d, pvp]
    L.rpx_field_prepare.restype = i32
    L.rpx_rays_project_to_sphere.argtypes = [vp, vp, vp, d, vp, vp]
    L.rpx_rays_project_to_sphere.restype = i32
    L.rpx_field_prepare_neighbours.argtypes = [vp, vp, vp, i32, vp, i32, d, vp, pvp]
    L.rpx_field_prepare_neighbours.restype = i32
    L.rpx_unit_evaluate_modes.argtypes = [vp, vp, vp, vp, vp, u64, i32, d, vp]
    L.rpx_unit_evaluate_modes.restype = i32
    L.rpx_field_count.argtypes = [vp]
    L.rpx_field_count.restype = u64
    L.rpx_field_modes.argtypes = [vp, vp, vp]
    L.rpx_field_modes.restype = i32
    L.rpx_field_evaluate.argtypes = [vp, vp, vp, u64, d, vp]
    L.rpx_field_evaluate.restype = i32
    L.rpx_field_evaluate_device.argtypes = [vp, vp, vp, u64, d, vp]
    L.rpx_field_evaluate_device.restype = i32
    L.rpx_field_last_ms.argtypes = [vp]
    L.rpx_field_last_ms.restype = d
    L.rpx_field_free.argtypes = [vp, vp]
    L.rpx_field_free.restype = None
    L.rpx_trace_streamed.argtypes = [vp, vp, u64, i32, d, i32, u64, vp, vp, i32, vp, vp, vp, vp]
    L.rpx_trace_streamed.restype = i32
    L.rpx_select_terminal.argtypes = [vp, vp, i32, i32, vp, pvp, vp]
    L.rpx_select_terminal.restype = i32
    L.rpx_rays_export_device.argtypes = [vp, vp, vp, u64]
    L.rpx_rays_export_device.restype = i32
    L.rpx_rays_import_device.argtypes = [vp, vp, u64, i32, pvp]
    L.rpx_rays_import_device.restype = i32
    L.rpx_detector_create.argtypes = [vp, vp, u64, vp, i32, d, d, pvp]
    L.rpx_detector_create.restype = i32
    L.rpx_detector_reset.argtypes = [vp, vp]
    L.rpx_detector_reset.restype = i32
    L.rpx_detector_accumulate.argtypes = [vp, vp, vp]
    L.rpx_detector_accumulate.restype = i32
    L.rpx_detector_read.argtypes = [vp, vp, vp]
    L.rpx_detector_read.restype = i32
    L.rpx_detector_field_device.argtypes = [vp]
    L.rpx_detector_field_device.restype = vp
    L.rpx_detector_npoints.argtypes = [vp]
    L.rpx_detector_npoints.restype = u64
    L.rpx_detector_modes.argtypes = [vp]
    L.rpx_detector_modes.restype = u64
    L.rpx_detector_ms.argtypes = [vp, vp]
    L.rpx_detector_ms.restype = d
    L.rpx_detector_free.argtypes = [vp, vp]
    L.rpx_detector_free.restype = None
    L.rpx_trace_step.argtypes = [vp, vp, d, pvp, vp]
    L.rpx_trace_step.restype = i32
    L.rpx_trace_consume.argtypes = [vp, vp, u64, i32, d, i32, vp, vp, vp]
    L.rpx_trace_consume.restype = i32
    if L.rpx_abi_version() != A.RPX_ABI_VERSION:
        raise RuntimeError("librpx.so ABI %d != binding ABI %d" % (L.rpx_abi_version(), A.RPX_ABI_VERSION))
    _LIB = L
    return L
