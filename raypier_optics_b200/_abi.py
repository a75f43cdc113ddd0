"""ctypes / numpy mirror of include/rpx.h (the C ABI of librpx).

Everything here is layout only: enum values, the numpy dtypes of the flat scene
tables and the ctypes ``rpx_scene`` struct that carries pointers to them.
"""
import ctypes as C

import numpy as np

RPX_ABI_VERSION = 1

# ---- ray records (ctracer.pyx:45-76 of the reference) -----------------------
ray_dtype = np.dtype([('origin', np.double, (3,)),
                      ('direction', np.double, (3,)),
                      ('normal', np.double, (3,)),
                      ('E_vector', np.double, (3,)),
                      ('refractive_index', np.complex128),
                      ('E1_amp', np.complex128),
                      ('E2_amp', np.complex128),
                      ('length', np.double),
                      ('phase', np.double),
                      ('accumulated_path', np.double),
                      ('wavelength_idx', np.uint32),
                      ('parent_idx', np.uint32),
                      ('end_face_idx', np.uint32),
                      ('ray_ident', np.uint32),
                      ('ray_type_id', np.uint32)])
para_dtype = np.dtype([('origin', np.double, (3,)),
                       ('direction', np.double, (3,)),
                       ('normal', np.double, (3,)),
                       ('length', np.double)])
NPARA = 6
gausslet_dtype = np.dtype([('base_ray', ray_dtype), ('para_rays', para_dtype, (NPARA,))])
assert ray_dtype.itemsize == 188 and para_dtype.itemsize == 80 and gausslet_dtype.itemsize == 668

REFL_RAY = 1
GAUSSLET = 2
PARABASAL = 4
NO_FACE = 0xFFFFFFFF

# ---- enums -------------------------------------------------------------------
FACE_CIRCULAR = 1
FACE_SHAPED_PLANAR = 2
FACE_IMPLICIT_PLANAR = 3
FACE_ELLIPTICAL_PLANE = 4
FACE_RECTANGULAR = 5
FACE_SPHERICAL = 6
FACE_SHAPED_SPHERICAL = 7
FACE_EXTRUDED_PLANAR = 8
FACE_POLYGON = 9
FACE_ORIENTED_POLYGON = 10
FACE_OFFAXIS_PARABOLIC = 11
FACE_ELLIPSOIDAL = 12
FACE_SADDLE = 13
FACE_CYLINDRICAL = 14
FACE_AXICON = 15
FACE_CONIC = 16
FACE_ASPHERIC = 17
FACE_EXT_POLY = 18
FACE_DISTORTION = 19
FACE_EXTRUDED_BEZIER = 20
FACE_MESH = 21
FACE_UVPATCH = 22

SHAPE_TRUE, SHAPE_CIRCLE, SHAPE_RECT, SHAPE_POLYGON, SHAPE_NOT, SHAPE_AND, SHAPE_OR, SHAPE_XOR = range(8)
IMPL_NULL, IMPL_PLANE, IMPL_SPHERE, IMPL_CYLINDER, IMPL_NEG, IMPL_MIN, IMPL_MAX, IMPL_SUB = range(8)
DIST_ZERNIKE_J7 = 1
DIST_ZERNIKE = 2
ZERNIKE_MAX_K = 64

MAT_OPAQUE = 1
MAT_TRANSPARENT = 2
MAT_PEC = 3
MAT_PARTIALLY_REFLECTIVE = 4
MAT_LINEAR_POLARISING = 5
MAT_WAVEPLATE = 6
MAT_DIELECTRIC = 7
MAT_FULL_DIELECTRIC = 8
MAT_COATED = 9
MAT_GRATING = 10
MAT_CIRC_APERTURE = 11
MAT_RECT_APERTURE = 12

PARA_DEFAULT, PARA_SNELL, PARA_GRATING = 0, 1, 2

TRACE_DEFAULT = 0
TRACE_KEEP_LAST_ONLY = 1
TRACE_EXACT_SYNC = 2

RPX_OK = 0
STATUS_NAMES = {0: "RPX_OK", -1: "RPX_ERR_INVALID", -2: "RPX_ERR_UNSUPPORTED", -3: "RPX_ERR_CUDA",
                -4: "RPX_ERR_NOMEM", -5: "RPX_ERR_NODEVICE", -6: "RPX_ERR_STATE"}

# ---- table dtypes (align=True reproduces the C struct layout) ---------------
FACE_NPARAM = 16
MAT_NPARAM = 12
face_dtype = np.dtype([('type', np.int32), ('face_set', np.int32), ('material', np.int32),
                       ('invert_normal', np.int32), ('shape_off', np.int32), ('shape_len', np.int32),
                       ('aux_off', np.int32), ('aux_n', np.int32), ('aux_m', np.int32),
                       ('base_face', np.int32), ('tolerance', np.double),
                       ('p', np.double, (FACE_NPARAM,))], align=True)
shape_op_dtype = np.dtype([('type', np.int32), ('aux_off', np.int32), ('aux_n', np.int32),
                           ('pad_', np.int32), ('p', np.double, (4,))], align=True)
implicit_op_dtype = np.dtype([('type', np.int32), ('pad_', np.int32), ('p', np.double, (7,))],
                             align=True)
ztape_op_dtype = np.dtype([('kind', np.int32), ('dst', np.int32), ('a', np.int32), ('b', np.int32),
                           ('c', np.int32), ('d', np.int32), ('e', np.int32), ('pad_', np.int32)],
                          align=True)
zcoef_dtype = np.dtype([('j', np.int32), ('n', np.int32), ('m', np.int32), ('k', np.int32),
                        ('value', np.double), ('opR', np.int32), ('opRp', np.int32),
                        ('opRr', np.int32), ('opR_z', np.int32)], align=True)
distortion_dtype = np.dtype([('type', np.int32), ('n_coefs', np.int32), ('coef_off', np.int32),
                             ('tape_z_off', np.int32), ('tape_z_len', np.int32),
                             ('tape_g_off', np.int32), ('tape_g_len', np.int32), ('k_max', np.int32),
                             ('p', np.double, (4,))], align=True)
material_dtype = np.dtype([('type', np.int32), ('para_model', np.int32), ('ntab_off', np.int32),
                           ('pad_', np.int32), ('p', np.double, (MAT_NPARAM,))], align=True)
transform_dtype = np.dtype([('m', np.double, (9,)), ('t', np.double, (3,))], align=True)
face_set_dtype = np.dtype([('trans', transform_dtype), ('inv_trans', transform_dtype),
                           ('face_begin', np.int32), ('face_end', np.int32)], align=True)

assert face_dtype.itemsize == 176
assert shape_op_dtype.itemsize == 48
assert implicit_op_dtype.itemsize == 64
assert ztape_op_dtype.itemsize == 32
assert zcoef_dtype.itemsize == 40
assert distortion_dtype.itemsize == 64
assert material_dtype.itemsize == 112
assert transform_dtype.itemsize == 96
assert face_set_dtype.itemsize == 200


class rpx_scene(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_traced_faces", C.c_int32), ("n_faces", C.c_int32),
                ("n_face_sets", C.c_int32), ("n_materials", C.c_int32), ("n_shape_ops", C.c_int32),
                ("n_implicit_ops", C.c_int32), ("n_distortions", C.c_int32), ("n_zcoefs", C.c_int32),
                ("n_ztape", C.c_int32), ("n_wavelengths", C.c_int32), ("n_ntab", C.c_int32),
                ("n_pool", C.c_int32), ("pad_", C.c_int32),
                ("faces", C.c_void_p), ("face_sets", C.c_void_p), ("materials", C.c_void_p),
                ("shape_ops", C.c_void_p), ("implicit_ops", C.c_void_p), ("distortions", C.c_void_p),
                ("zcoefs", C.c_void_p), ("ztape", C.c_void_p), ("wavelengths", C.c_void_p),
                ("ntab", C.c_void_p), ("pool", C.c_void_p)]


assert C.sizeof(rpx_scene) == 144


# ---- rpx_trace_consume (include/rpx.h) ----------------------------------------
CONSUME_SOURCE_ON_DEVICE = 1
CONSUME_TERMINAL = 2
CONSUME_CAPTURE = 4
CONSUME_FIELD = 8
CONSUME_MAX_GENS = 256


class rpx_consume_opts(C.Structure):
    _fields_ = [("chunk_rays", C.c_uint64), ("flags", C.c_uint32), ("max_gens", C.c_int32),
                ("terminal_faces", C.c_void_p), ("terminal_capacity", C.c_uint64),
                ("captured_capacity", C.c_uint64), ("detector", C.c_void_p),
                ("per_chunk_terminal", C.c_void_p), ("per_chunk_captured", C.c_void_p)]


class rpx_consume_result(C.Structure):
    _fields_ = [("n_gens", C.c_int32), ("n_chunks", C.c_int32), ("counts", C.c_uint64 * CONSUME_MAX_GENS),
                ("n_terminal", C.c_uint64), ("n_captured", C.c_uint64), ("terminal", C.c_void_p),
                ("captured", C.c_void_p), ("device_ms", C.c_double), ("trace_ms", C.c_double),
                ("shade_ms", C.c_double), ("intersect_ms", C.c_double), ("shade_launches", C.c_uint64),
                ("intersect_launches", C.c_uint64), ("launches", C.c_uint64)]


assert C.sizeof(rpx_consume_opts) == 64
assert C.sizeof(rpx_consume_result) == 8 + 8 * CONSUME_MAX_GENS + 16 + 16 + 32 + 24
