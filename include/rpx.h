/*
 * rpx.h -- C ABI of librpx, the B200-native (sm_100a) replacement for the
 * non-sequential ray-tracing core of raypier (raypier/core).
 *
 * Drop-in boundary: librpx replaces what
 *     raypier.core.tracer.trace_rays            (raypier/core/tracer.py:9-47)
 * does between "scene and source rays are known" and "every generation of
 * rays is known":  ctracer.trace_segment_c (raypier/core/ctracer.pyx:2062-2118),
 * ctracer.trace_gausslet_c + trace_parabasal_rays (ctracer.pyx:2214-2281,
 * 2350-2385), FaceList.intersect_c / compute_orientation_c (ctracer.pyx:1882-1953),
 * every cfaces.*.intersect_c / compute_normal_c (raypier/core/cfaces.pyx) and
 * every cmaterials.*.eval_child_ray_c / eval_parabasal_ray_c
 * (raypier/core/cmaterials.pyx).
 *
 * The reference has no FFI for this path (it is Cython calling Cython); the
 * entry points below are what a `ctypes`/`cffi` binding inside
 * raypier/core/tracer.py would bind (see INTEGRATION.md for that stub).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions cross the ABI
 *   - every function returns RPX_OK (0) or a negative rpx_status; the message
 *     is available from rpx_last_error()
 *   - the caller owns every buffer it passes in; the library owns device
 *     buffers and rpx_result/rpx_rays handles until the matching *_free
 *   - ray records cross the ABI in the reference's own packed AoS layouts
 *     (ray_t 188 B, gausslet_t 668 B, raypier/core/ctracer.pxd:40-64)
 *   - there is NO CPU fallback: without a CUDA device rpx_init fails
 */
#ifndef RPX_H_
#define RPX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPX_ABI_VERSION 1

/* ------------------------------------------------------------------ status */
typedef enum rpx_status {
    RPX_OK = 0,
    RPX_ERR_INVALID = -1,     /* bad argument / malformed scene               */
    RPX_ERR_UNSUPPORTED = -2, /* face / material type not implemented         */
    RPX_ERR_CUDA = -3,        /* CUDA runtime error                           */
    RPX_ERR_NOMEM = -4,       /* device or host allocation failed             */
    RPX_ERR_NODEVICE = -5,    /* no CUDA device: there is no CPU fallback     */
    RPX_ERR_STATE = -6        /* call order (e.g. trace before scene_set)     */
} rpx_status;

/* --------------------------------------------------------- ray record AoS */
/* Bit-compatible with ray_t / para_t / gausslet_t (ctracer.pxd:40-64) and the
 * numpy ray_dtype / gausslet_dtype (ctracer.pyx:45-76).                      */
#pragma pack(push, 1)
typedef struct rpx_ray {
    double origin[3], direction[3], normal[3], E_vector[3];
    double refractive_index[2], E1_amp[2], E2_amp[2]; /* complex128 (re, im)  */
    double length, phase, accumulated_path;
    uint32_t wavelength_idx, parent_idx, end_face_idx, ray_ident, ray_type_id;
} rpx_ray; /* 188 bytes */

typedef struct rpx_para {
    double origin[3], direction[3], normal[3];
    double length;
} rpx_para; /* 80 bytes */

typedef struct rpx_gausslet {
    rpx_ray base_ray;
    rpx_para para[6];
} rpx_gausslet; /* 668 bytes */
#pragma pack(pop)

#define RPX_RAY_BYTES 188u
#define RPX_GAUSSLET_BYTES 668u
#define RPX_NPARA 6

/* ray_type_id bit flags (ctracer.pxd:35-37) */
#define RPX_REFL_RAY 1u
#define RPX_GAUSSLET 2u
#define RPX_PARABASAL 4u
/* end_face_idx of an unterminated ray: (unsigned)-1 (ctracer.pyx:2087) */
#define RPX_NO_FACE 0xFFFFFFFFu

/* ------------------------------------------------------------- face types */
/* One enum value per concrete cfaces.pyx class; p[] holds its parameters.   */
typedef enum rpx_face_type {
    RPX_FACE_CIRCULAR = 1,         /* cfaces.pyx:140  p: diameter, offset, z_plane, invert_normals */
    RPX_FACE_SHAPED_PLANAR = 2,    /* :194  p: z_height ; shape                                     */
    RPX_FACE_IMPLICIT_PLANAR = 3,  /* :251  p: origin[3], normal[3] ; aux = implicit program        */
    RPX_FACE_ELLIPTICAL_PLANE = 4, /* :313  p: g_x, g_y, diameter                                   */
    RPX_FACE_RECTANGULAR = 5,      /* :354  p: length, width, offset, z_plane                       */
    RPX_FACE_SPHERICAL = 6,        /* :410  p: diameter, curvature, z_height                        */
    RPX_FACE_SHAPED_SPHERICAL = 7, /* :501  p: curvature, z_height ; shape                          */
    RPX_FACE_EXTRUDED_PLANAR = 8,  /* :611  p: x1, y1, x2, y2, z1, z2, normal[3]                    */
    RPX_FACE_POLYGON = 9,          /* :1077 p: z_plane ; aux = xy points (pool)                     */
    RPX_FACE_ORIENTED_POLYGON = 10,/* :1121 p: origin[3], normal[3], x_axis[3], y_axis[3]; aux pts  */
    RPX_FACE_OFFAXIS_PARABOLIC = 11,/* :1224 p: EFL, diameter, height                               */
    RPX_FACE_ELLIPSOIDAL = 12,     /* :1320 p: major, minor, x1,x2,y1,y2,z1,z2; aux = 2 transforms  */
    RPX_FACE_SADDLE = 13,          /* :1429 p: z_height, curvature ; shape                          */
    RPX_FACE_CYLINDRICAL = 14,     /* :1516 p: z_height, radius ; shape                             */
    RPX_FACE_AXICON = 15,          /* :1607 p: z_height, gradient ; shape                           */
    RPX_FACE_CONIC = 16,           /* :1751 p: curvature, z_height, conic_const, invert_normals; shape */
    RPX_FACE_ASPHERIC = 17,        /* :1885 p: curvature, z_height, conic_const, invert_normals,
                                              A4,A6,A8,A10,A12,A14,A16, atol ; shape                */
    RPX_FACE_EXT_POLY = 18,        /* :2130 p: R(=-curvature), beta(=1+k), norm_radius, z_height,
                                              atol, invert_normals ; aux = coefs[Nx][Ny] (pool)     */
    RPX_FACE_DISTORTION = 19,      /* :2323 p: accuracy ; base_face, aux = distortion idx ; shape   */
    RPX_FACE_EXTRUDED_BEZIER = 20, /* :795  p: z_height_1, z_height_2, mincorner[2], maxcorner[2];
                                              aux = cubic Bezier segments [n][4][2] (pool)          */
    RPX_FACE_MESH = 21,            /* obbtree.pyx:880-946 OBBTreeFace over an OBBTree (:200-400): a
                                      triangle mesh.  p: tree tolerance (OBBTree.tolerance, 0.1);
                                      aux_off = mesh block in pool, aux_n = cells, aux_m = BVH nodes.
                                      Block (doubles): header[8] = n_points, n_cells, n_nodes,
                                      off_points, off_cells, off_tris, off_nodes (relative to aux_off),
                                      tolerance; points[n_points][3]; cells[n_cells][3] (vertex ids);
                                      tris[n_cells][16] in BVH leaf order = p1, v1 = p2 - p1,
                                      v2 = p3 - p1, n = v1 x v2, cell id, 3 pad; nodes[n_nodes][8] =
                                      box min[3], box max[3], then (left, right) child ids for an inner
                                      node or (-(first tri) - 1, count) for a leaf.  The reference's OBB
                                      tree is an acceleration structure (its node test only prunes); this
                                      library carries its own BVH, built by the host side.  rpx_scene_set
                                      repacks it for the device walk (64-byte nodes with both children's
                                      boxes in outward-rounded fp32; triangle tests stay fp64 on these
                                      records, so hits are bit-identical to a walk of the fp64 nodes) and
                                      validates links and depth.  The facet a ray hit travels with the device
                                      collection in a side array (intersect_t.piece_idx, ctracer.pxd), so the
                                      generation that shades the ray does not walk the tree again.          */
    RPX_FACE_UVPATCH = 22          /* cbezier.pyx:391-550 UVPatchFace over a BezierPatch (:200-286) or BSplinePatch
                                      (:290-388).  p: atol, invert_normals, patch kind (0 Bezier, 1 B-spline),
                                      N, M (orders: (N+1) x (M+1) control points), u_degree, v_degree, offset of
                                      the patch block in the pool, number of u knots, number of v knots.
                                      aux_off / aux_n / aux_m = mesh block of the patch's (u_res x v_res)
                                      tessellation (get_mesh, :153-197) exactly as for RPX_FACE_MESH.  Patch block
                                      (doubles): uvs[n_points][2] (the vertex (u, v) of the tessellation),
                                      ctrl[N+1][M+1][3], then binom_n[N+1], binom_m[M+1] (Bezier) or
                                      u_knots[], v_knots[] (B-spline).  The nearest facet seeds a Newton
                                      iteration on the patch itself (<= 100 steps, |du|, |dv| < atol).         */
} rpx_face_type;

#define RPX_FACE_NPARAM 16

typedef struct rpx_face {
    int32_t type;          /* rpx_face_type                                              */
    int32_t face_set;      /* index into scene.face_sets (the owning FaceList)           */
    int32_t material;      /* index into scene.materials                                 */
    int32_t invert_normal; /* Face.invert_normal (ctracer.pyx:1948)                      */
    int32_t shape_off;     /* first op of this face's shape program, -1 = none           */
    int32_t shape_len;     /* number of ops                                              */
    int32_t aux_off;       /* type-specific: offset into pool / implicit ops / distortions */
    int32_t aux_n;         /* type-specific count (points, Nx, ops ...)                  */
    int32_t aux_m;         /* type-specific second count (Ny)                            */
    int32_t base_face;     /* DISTORTION: index into scene.faces of the wrapped face     */
    double tolerance;      /* Face.tolerance (default 1e-4, ctracer.pyx:1736)            */
    double p[RPX_FACE_NPARAM];
} rpx_face;

/* --------------------------------------------------- 2-D aperture shapes */
/* cshapes.pyx trees flattened to a postfix (RPN) program evaluated on a bit
 * stack.  Leaves push, NOT pops 1 / pushes 1, AND/OR/XOR pop 2 / push 1.    */
typedef enum rpx_shape_op_type {
    RPX_SHAPE_TRUE = 0,    /* base Shape: always inside (ctracer.pyx:1659)  */
    RPX_SHAPE_CIRCLE = 1,  /* cshapes.pyx:102  p: cx, cy, radius             */
    RPX_SHAPE_RECT = 2,    /* :119  p: cx, cy, width, height                 */
    RPX_SHAPE_POLYGON = 3, /* :139  aux_off/aux_n -> xy points in pool       */
    RPX_SHAPE_NOT = 4,     /* :34 */
    RPX_SHAPE_AND = 5,     /* :56 */
    RPX_SHAPE_OR = 6,      /* :60 */
    RPX_SHAPE_XOR = 7      /* :64 */
} rpx_shape_op_type;

typedef struct rpx_shape_op {
    int32_t type;
    int32_t aux_off, aux_n;
    int32_t pad_;
    double p[4];
} rpx_shape_op;

/* ---------------------------------------------------- implicit surfaces */
/* cimplicit_surfs.pyx trees as an RPN program on a double stack.           */
typedef enum rpx_implicit_op_type {
    RPX_IMPL_NULL = 0,     /* :23  pushes -1.0                               */
    RPX_IMPL_PLANE = 1,    /* :27  p: origin[3], normal[3] (normalised)      */
    RPX_IMPL_SPHERE = 2,   /* :60  p: centre[3], radius                      */
    RPX_IMPL_CYLINDER = 3, /* :83  p: origin[3], axis[3] (normalised), radius */
    RPX_IMPL_NEG = 4,      /* Invert :122                                    */
    RPX_IMPL_MIN = 5,      /* Union.apply_op :173                            */
    RPX_IMPL_MAX = 6,      /* Intersection.apply_op :181                     */
    RPX_IMPL_SUB = 7       /* Difference.apply_op :189                       */
} rpx_implicit_op_type;

typedef struct rpx_implicit_op {
    int32_t type;
    int32_t pad_;
    double p[7];
} rpx_implicit_op;

/* ------------------------------------------------------------ distortions */
typedef enum rpx_distortion_type {
    RPX_DIST_ZERNIKE_J7 = 1, /* SimpleTestZernikeJ7 cdistortions.pyx:39  p: unit_radius, amplitude */
    RPX_DIST_ZERNIKE = 2     /* ZernikeDistortion   cdistortions.pyx:323 p: unit_radius            */
} rpx_distortion_type;

/* The reference evaluates Zernike radial polynomials by memoised recursion
 * (cdistortions.pyx:149-316).  Which memo slot is filled when depends only on
 * the coefficient set, never on the ray, so the host unrolls the recursion
 * ONCE into a straight-line "tape" (exactly the reference's evaluation order,
 * including its slot-aliasing quirks) and the device just runs the tape.
 * Operand encoding: 0 -> constant 0.0, 1 -> constant 1.0,
 *                   2 + 3*slot + w -> workspace[w][slot], w in {0:R, 1:R', 2:R/r} */
typedef struct rpx_ztape_op {
    int32_t kind;  /* 0: R      ws0[dst] = r*(a + b) - c
                      1: R'     ws1[dst] = (a + b) + r*(d + e) - c
                      2: R/r    ws2[dst] = (a + b) - c                      */
    int32_t dst;   /* slot k                                                 */
    int32_t a, b, c, d, e; /* operands (see encoding above)                  */
    int32_t pad_;
} rpx_ztape_op;

typedef struct rpx_zcoef {
    int32_t j, n, m, k;   /* ANSI index and (n, m, k) (cdistortions.pyx:104-136) */
    double value;
    int32_t opR, opRp, opRr; /* operands holding R, R', R/r after the tape ran
                                (gradient tape); opR_z: R after the z-only tape */
    int32_t opR_z;
} rpx_zcoef;

typedef struct rpx_distortion {
    int32_t type;
    int32_t n_coefs, coef_off;       /* into scene.zcoefs                    */
    int32_t tape_z_off, tape_z_len;  /* tape for z_offset_c (R only)         */
    int32_t tape_g_off, tape_g_len;  /* tape for z_offset_and_gradient_c     */
    int32_t k_max;                   /* workspace slots                      */
    double p[4];
} rpx_distortion;

#define RPX_ZERNIKE_MAX_K 64 /* device workspace bound: n_max <= 12 */

/* --------------------------------------------------------------- materials */
typedef enum rpx_material_type {
    RPX_MAT_OPAQUE = 1,               /* cmaterials.pyx:242                                   */
    RPX_MAT_TRANSPARENT = 2,          /* :254                                                 */
    RPX_MAT_PEC = 3,                  /* :281                                                 */
    RPX_MAT_PARTIALLY_REFLECTIVE = 4, /* :322  p: reflectivity                                */
    RPX_MAT_LINEAR_POLARISING = 5,    /* :400                                                 */
    RPX_MAT_WAVEPLATE = 6,            /* :458  p: retardance_re, retardance_im, fast_axis[3]  */
    RPX_MAT_DIELECTRIC = 7,           /* :554  n tables                                       */
    RPX_MAT_FULL_DIELECTRIC = 8,      /* :727 and :875 (dispersive)  p: refl_thr, trans_thr   */
    RPX_MAT_COATED = 9,               /* :1017 and :1187 (dispersive) p: refl_thr, trans_thr, thickness */
    RPX_MAT_GRATING = 10,             /* :1437 p: lines_per_mm, order, efficiency, origin[3]  */
    RPX_MAT_CIRC_APERTURE = 11,       /* :1602 p: outer_radius, radius, edge_width, invert, origin[3] */
    RPX_MAT_RECT_APERTURE = 12        /* :1677 p: outer_width, outer_height, width, height, edge_width, invert, origin[3] */
} rpx_material_type;

/* rpx_material.para_model: which eval_parabasal_ray_c the reference class has */
#define RPX_PARA_DEFAULT 0   /* InterfaceMaterial default, ctracer.pyx:1588-1610   */
#define RPX_PARA_SNELL 1     /* DielectricMaterial :683 / CoatedDispersive :1393   */
#define RPX_PARA_GRATING 2   /* DiffractionGratingMaterial :1544                   */

#define RPX_MAT_NPARAM 12

typedef struct rpx_material {
    int32_t type;
    int32_t para_model;
    int32_t ntab_off; /* offset (in complex elements) into scene.ntab of this material's
                         [3][n_wavelengths] table: row 0 n_inside, 1 n_outside, 2 n_coating;
                         non-dispersive materials carry their constant n replicated, so the
                         device has one code path (on_set_wavelengths, cmaterials.pyx:883,1220) */
    int32_t pad_;
    double p[RPX_MAT_NPARAM];
} rpx_material;

/* ------------------------------------------------------------- transforms */
typedef struct rpx_transform {
    double m[9]; /* row-major m00..m22 (transform_t, ctracer.pxd:67-69) */
    double t[3];
} rpx_transform;

typedef struct rpx_face_set {
    rpx_transform trans;     /* FaceList.trans      */
    rpx_transform inv_trans; /* FaceList.inv_trans  */
    int32_t face_begin, face_end; /* [begin, end) into the traced face list */
} rpx_face_set;

/* ------------------------------------------------------------------ scene */
typedef struct rpx_scene {
    int32_t abi_version;       /* RPX_ABI_VERSION */
    int32_t n_traced_faces;    /* len(all_faces): faces[0..n) are traced, idx == position;
                                  faces[n..n_faces) are only referenced as DistortionFace bases */
    int32_t n_faces;
    int32_t n_face_sets;
    int32_t n_materials;
    int32_t n_shape_ops;
    int32_t n_implicit_ops;
    int32_t n_distortions;
    int32_t n_zcoefs;
    int32_t n_ztape;
    int32_t n_wavelengths;
    int32_t n_ntab;            /* complex elements in ntab */
    int32_t n_pool;            /* doubles in pool          */
    int32_t pad_;
    const rpx_face* faces;
    const rpx_face_set* face_sets;
    const rpx_material* materials;
    const rpx_shape_op* shape_ops;
    const rpx_implicit_op* implicit_ops;
    const rpx_distortion* distortions;
    const rpx_zcoef* zcoefs;
    const rpx_ztape_op* ztape;
    const double* wavelengths; /* microns, RayCollection.wavelengths */
    const double* ntab;        /* interleaved (re, im) */
    const double* pool;        /* polygon points, ext-poly coefs, ellipsoid transforms */
} rpx_scene;

/* ------------------------------------------------------------ the library */
typedef struct rpx_ctx rpx_ctx;       /* one per process per GPU              */
typedef struct rpx_rays rpx_rays;     /* a device-resident SoA ray generation */
typedef struct rpx_result rpx_result; /* all generations of one trace         */

/* rpx_trace flags */
#define RPX_TRACE_DEFAULT 0u
#define RPX_TRACE_KEEP_LAST_ONLY 1u /* streaming mode: keep counts + face counts, free
                                       generation g-1 once g is built (N too big to keep) */
#define RPX_TRACE_EXACT_SYNC 2u     /* read len(new_rays) back after every generation instead of
                                       pipelining launches on a one-generation-old count       */

/* Replaces: module import of raypier.core.ctracer (no device state exists there). */
int rpx_init(int device, rpx_ctx** out_ctx);
void rpx_shutdown(rpx_ctx* ctx);
/* Last error text for this context (or for a failed rpx_init when ctx==NULL). */
const char* rpx_last_error(const rpx_ctx* ctx);
int rpx_abi_version(void);

/* Replaces the per-trace set-up loop of trace_rays (core/tracer.py:28-37):
 * f.idx = i, f.update(), f.material.wavelengths = wavelengths,
 * fs.sync_transforms() (ctracer.pyx:1820-1837) -- the host flattens the objects
 * into rpx_scene after doing those calls; this uploads the tables.            */
int rpx_scene_set(rpx_ctx* ctx, const rpx_scene* scene);

/* Pinned host memory for ray records (replaces malloc in RayCollection.__cinit__,
 * ctracer.pyx:983-987, when the caller wants full-speed DMA).                 */
void* rpx_host_alloc(size_t bytes);
void rpx_host_free(void* p);

/* Replaces RayCollection.from_array / GaussletCollection.from_array
 * (ctracer.pyx:1142-1154, 1304-1319): copies n packed AoS records to the device
 * and transposes them into the SoA generation buffer.                         */
int rpx_rays_upload(rpx_ctx* ctx, const void* aos, uint64_t n, int is_gausslet,
                    rpx_rays** out_rays);
/* Replaces copy_as_array (ctracer.pyx:1048-1054, 1286-1292).                  */
int rpx_rays_download(rpx_ctx* ctx, const rpx_rays* rays, void* out_aos, uint64_t capacity);
uint64_t rpx_rays_count(const rpx_rays* rays);
/* Device-to-device copy of a generation (a trace consumes and mutates its input, so a
 * caller that re-traces the same source keeps a pristine copy).                       */
int rpx_rays_clone(rpx_ctx* ctx, const rpx_rays* rays, rpx_rays** out_rays);
void rpx_rays_free(rpx_ctx* ctx, rpx_rays* rays);

/* Replaces the generation loop of trace_rays (core/tracer.py:22,39-45) over
 * trace_segment_c / trace_gausslet_c, inputs already resident on the device.
 * `rays` becomes generation 0 of the result (it is mutated like the reference
 * mutates its parent collection: length, end_face_idx).  The call takes ownership
 * of `rays` whether it succeeds or fails: never rpx_rays_free it afterwards.  max_length is rounded to float for plain rays exactly as
 * trace_segment_c's `float max_length` argument does (ctracer.pyx:2066).      */
int rpx_trace_device(rpx_ctx* ctx, rpx_rays* rays, double max_length, int recursion_limit,
                     uint32_t flags, rpx_result** out_result);

/* Convenience = rpx_rays_upload + rpx_trace_device (host buffers in).         */
int rpx_trace(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet,
              double max_length, int recursion_limit, uint32_t flags,
              rpx_result** out_result);

/* ONE generation (trace_segment_c / trace_gausslet_c, ctracer.pyx:2062-2118, 2214-2281) for traces that
 * need the host between generations: ResampleGaussletMaterial (cmaterials.pyx:1766-1831) hands the
 * gausslets that reached it to a Python callback and appends what it returns to the new generation
 * (eval_decomposed_rays_c, ctracer.pyx:2274-2278).  `rays` (still owned by the caller) is intersected
 * with every face and written back in place; *out_children owns the new generation, NOT yet intersected
 * (the next step does that; length = INF, or max_length for gausslets, :2280).  face_counts[n_traced_faces]
 * (may be NULL) is ADDED to: Face.count accumulates over the generations of a trace (:2108).          */
int rpx_trace_step(rpx_ctx* ctx, rpx_rays* rays, double max_length, rpx_rays** out_children,
                   uint32_t* face_counts);

/* Replaces trace_ray_sequence (core/tracer.py:50-99) over trace_one_face_segment_c /
 * trace_one_face_gausslet_c (ctracer.pyx:2121-2170, 2284-2347): step s intersects ONLY the
 * face with global index face_seq[s] (FaceList.intersect_one_face_c, ctracer.pyx:1861-1879).
 * traced_rays[0] is the input; the generation produced by the last step is returned
 * untraced (length INF / max_length for gausslets, end_face_idx = the parent's).         */
int rpx_trace_sequence(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet,
                       double max_length, int recursion_limit, const int32_t* face_seq,
                       int n_seq, rpx_result** out_result);

/* len(traced_rays) */
int rpx_result_n_generations(const rpx_result* res);
/* [len(traced_rays[g]) for g]; counts has room for n_generations entries */
int rpx_result_counts(const rpx_result* res, uint64_t* counts);
/* traced_rays[g].copy_as_array() into out_aos (capacity in records) */
int rpx_result_generation(rpx_ctx* ctx, const rpx_result* res, int g, void* out_aos,
                          uint64_t capacity);
/* Face.count for every traced face (ctracer.pyx:2108), n_traced_faces entries */
int rpx_result_face_counts(const rpx_result* res, uint32_t* counts);
/* Device time of the generation loop (CUDA events on the tracing stream), ms */
double rpx_result_device_ms(const rpx_result* res);
/* Kernels launched by this trace (for bench.py's gpu_launches claim) */
uint64_t rpx_result_launches(const rpx_result* res);
/* Average device time (ms) and launch count per kernel family over this trace:
 * which = 0 intersect, 1 shade (orientation + material + ordered emit)       */
int rpx_result_kernel_ms(const rpx_result* res, int which, double* total_ms, uint64_t* launches);
void rpx_result_free(rpx_ctx* ctx, rpx_result* res);

/* Raw CUDA stream the context launches on (so a caller can bracket the trace
 * with its own events); returned as void* to keep CUDA types out of the ABI. */
void* rpx_stream(rpx_ctx* ctx);

/* rpx_trace for a host-resident source cut into contiguous chunks of `chunk_rays` (0 = default):
 * the upload of chunk c+1 and the download of chunk c's generations overlap the tracing (three
 * streams, full-duplex PCIe), and the device holds two chunks at most -- also the way to trace a
 * source larger than HBM.  Results are those of one rpx_trace call: out_gens[g] (caller memory,
 * out_capacity[g] records; pinned memory for real overlap) receives generation g = the chunks'
 * generation g concatenated in source order, parent_idx renumbered globally; out_counts[g] =
 * len(traced_rays[g]); *n_gens = len(traced_rays) (<= max_gens, else RPX_ERR_INVALID; a too small
 * buffer gives RPX_ERR_NOMEM); face_counts[n_traced_faces] = Face.count; *device_ms = summed
 * device time of the generation loops.  (core/tracer.py:9-47 + the sharding of SURVEY 8e.)
 * In-place tracing: out_gens[0] == rays_aos is allowed -- the reference's own convention, traced_rays[0]
 * IS input_rays with `length` and `end_face_idx` written back (ctracer.pyx:2086-2087, 1900-1903).        */
int rpx_trace_streamed(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                       int recursion_limit, uint64_t chunk_rays, void* const* out_gens,
                       const uint64_t* out_capacity, int max_gens, uint64_t* out_counts, int* n_gens,
                       uint32_t* face_counts, double* device_ms);

/* ---------------------------------------------------------- capture planes
 * select_ray_intersections / select_gausslet_intersections (ctracer.pyx:1981-2058), the
 * filter behind probes.py RayCapturePlane / GaussletCapturePlane (:119-143): every ray of every
 * generation is re-intersected, between its origin and its traced end point, with ONE FaceList
 * (normally a single RectangularFace); hits are appended in (generation, ray) order with
 * length = distance to the capture face and end_face_idx = that face's idx.  Run on the device
 * it removes the need to ship all generations to the host.                                    */
/* The capture FaceList as a scene of its own (one face set; materials are ignored).
 * face_ids[i] = the Python-side Face.idx of face i, written into end_face_idx of captured rays
 * (the reference never assigns idx to capture faces, so it is whatever the caller left there);
 * NULL = position in the table.                                                                */
int rpx_capture_scene_set(rpx_ctx* ctx, const rpx_scene* capture_scene, const uint32_t* face_ids);
/* Borrowed handle on generation g of a finished trace (NULL if dropped / out of range).        */
const rpx_rays* rpx_result_rays(const rpx_result* res, int g);
/* Filter n_gens device-resident collections.  wl_offsets[j] is added to wavelength_idx of the
 * rays taken from collection j (the running `wl_offset` of the reference loop) and the sum is
 * then mapped through wl_map[n_wl_map] (the `inverse` of np.unique over the concatenated
 * wavelength lists, :2011-2014); wl_map == NULL leaves the offset index.  counts[j] (may be NULL)
 * receives the number of rays captured from collection j.  *out owns a new device collection. */
int rpx_capture(rpx_ctx* ctx, const rpx_rays* const* gens, int n_gens, const uint32_t* wl_offsets,
                const uint32_t* wl_map, uint32_t n_wl_map, rpx_rays** out, uint64_t* counts);

/* ---------------------------------------------------------- E-field summation
 * The field at a set of points as the sum of general astigmatic Gaussian modes, one per ray:
 * raypier/core/cfields.pyx sum_gaussian_modes (:51-115) + calc_mode_U (:118-153), with the
 * gausslet front end of raypier/core/fields.py (evaluate_neighbours_gc :114-137 ->
 * cfields.evaluate_modes :217-228), i.e. eval_Efield_from_gausslets / EFieldSummation
 * (fields.py:206-277).  O(N_ray x N_pt) complex exponentials in fp64.                         */
typedef struct rpx_field rpx_field;
/* Prepare the per-ray mode records on the device (EFieldSummation.__init__, fields.py:212-221).
 * modes == NULL: `rays` must be gausslets; the modes (A, B, C) are fitted on the device from the
 * six parabasal rays with the given blending.  modes != NULL: n x 3 complex128 host array, used
 * as given with the (base) rays -- the exact argument list of sum_gaussian_modes.
 * wavelengths[n_wavelengths] in microns (the `wavelengths` argument, indexed by wavelength_idx). */
int rpx_field_prepare(rpx_ctx* ctx, const rpx_rays* rays, const double* modes, const double* wavelengths,
                      int n_wavelengths, double blending, rpx_field** out);
/* The plain-ray front end, eval_Efield_from_rays (raypier/core/fields.py:206-229):
 *   project_to_sphere (fields.py:50-77) on a device-resident collection of plain rays, IN PLACE: every
 *   ray whose line meets the sphere (centre[3], radius) is moved to the intersection with the most
 *   negative path and its accumulated_path corrected by alpha * n.real.  selector (host, n bytes, may
 *   be NULL) receives the reference's boolean `selector`; *n_selected (may be NULL) its sum.  The
 *   reference returns rays[selector]; rays with selector 0 are left untouched here.              */
int rpx_rays_project_to_sphere(rpx_ctx* ctx, rpx_rays* rays, const double* centre, double radius, uint8_t* selector,
                               uint64_t* n_selected);
/*   evaluate_neighbours (fields.py:80-111) -> cfields.evaluate_modes (cfields.pyx:217-228) -> mode
 *   records: neighbours is the n x row_size int32 host array of RayCollection.neighbours
 *   (ctracer.pyx:1084-1131; -1 = no neighbour; row_size must be 6); rays lacking a neighbour are
 *   dropped (`mask`, fields.py:97), so rpx_field_count() = number of kept rays.  xy_out (host, may be
 *   NULL) receives x, y, dx, dy of the kept rays as four consecutive n_kept x 6 blocks.           */
int rpx_field_prepare_neighbours(rpx_ctx* ctx, const rpx_rays* rays, const int32_t* neighbours, int row_size,
                                 const double* wavelengths, int n_wavelengths, double blending, double* xy_out,
                                 rpx_field** out);
/*   cfields.evaluate_modes(x, y, dx, dy, blending) (cfields.pyx:217-228) on explicit n x 6 host arrays;
 *   modes_out is n x 3 complex128.                                                                 */
int rpx_unit_evaluate_modes(rpx_ctx* ctx, const double* x, const double* y, const double* dx, const double* dy,
                            uint64_t n, int row_size, double blending, double* modes_out);
uint64_t rpx_field_count(const rpx_field* field);
/* The (A, B, C) modes, n x 3 complex128 (fields.py ExtractGamma :196-203) */
int rpx_field_modes(rpx_ctx* ctx, const rpx_field* field, double* modes_out);
/* sum_gaussian_modes(rays, modes, wavelengths, points, time_ps): points is npt x 3 doubles, out
 * npt x 3 complex128 (host memory; overwritten).                                               */
int rpx_field_evaluate(rpx_ctx* ctx, rpx_field* field, const double* points, uint64_t npt, double time_ps,
                       double* out);
/* Same with DEVICE pointers; d_out is ACCUMULATED into (zero it first) so partial fields of
 * several ray shards can share one buffer before an all-reduce.  Returns after the kernel ended.
 * The kernel runs on rpx_stream(ctx), a NON-BLOCKING stream: a caller that produced d_points / d_out
 * on another stream (a torch tensor's zero fill, say) must synchronise that stream first.          */
int rpx_field_evaluate_device(rpx_ctx* ctx, rpx_field* field, const double* d_points, uint64_t npt,
                              double time_ps, double* d_out);
/* Device time (ms) of the last summation kernel of this field (CUDA events) */
double rpx_field_last_ms(const rpx_field* field);
void rpx_field_free(rpx_ctx* ctx, rpx_field* field);

/* ---------------------------------------------------------- terminal rays
 * The rays a trace ends with: rays that hit nothing -- end_face_idx left at (unsigned)-1 by the write-back
 * of trace_segment_c / FaceList.intersect_c (ctracer.pyx:2086-2087, 1900-1903) -- and / or rays that end
 * on chosen faces (absorbers and detectors: BeamStop, an OpaqueMaterial target).  A device filter over
 * device-resident generations; selected records are appended unchanged in (collection, ray) order.
 * face_select: n_traced_faces bytes, non-zero = rays ending on that face are selected (NULL = none).
 * counts[j] (may be NULL) = rays selected from collection j.  *out owns a new device collection.        */
int rpx_select_terminal(rpx_ctx* ctx, const rpx_rays* const* gens, int n_gens, int select_unterminated,
                        const uint8_t* face_select, rpx_rays** out, uint64_t* counts);
/* Device-resident AoS export / import of a collection (the packed ray_t / gausslet_t records of
 * copy_as_array / from_array, ctracer.pyx:1048-1054, 1142-1154) into / from CALLER-OWNED DEVICE memory,
 * e.g. a torch tensor: what an NCCL gather of terminal rays sends and receives, with no host bounce.
 * d_aos must be 4-byte aligned and hold `capacity` records; both calls return after the copy finished. */
int rpx_rays_export_device(rpx_ctx* ctx, const rpx_rays* rays, void* d_aos, uint64_t capacity);
int rpx_rays_import_device(rpx_ctx* ctx, const void* d_aos, uint64_t n, int is_gausslet, rpx_rays** out_rays);

/* ---------------------------------------------------------- detector (accumulating E-field)
 * EFieldSummation / eval_Efield_from_gausslets (fields.py:206-277) as a long-lived device object: a fixed
 * set of points and a complex field that gausslet collections are summed INTO, one collection after the
 * other (the sum over rays is associative), so that a trace too large to keep can feed it chunk by chunk.
 * The partial fields of several GPUs are combined with one all-reduce of rpx_detector_field_device().   */
typedef struct rpx_detector rpx_detector;
int rpx_detector_create(rpx_ctx* ctx, const double* points, uint64_t npt, const double* wavelengths, int n_wavelengths,
                        double blending, double time_ps, rpx_detector** out);
/* field := 0, modes := 0 */
int rpx_detector_reset(rpx_ctx* ctx, rpx_detector* det);
/* Fit the modes of a device-resident gausslet collection and add their field at every point. */
int rpx_detector_accumulate(rpx_ctx* ctx, rpx_detector* det, const rpx_rays* gausslets);
/* npt x 3 complex128 into host memory */
int rpx_detector_read(rpx_ctx* ctx, rpx_detector* det, double* field_out);
/* device pointer of the npt x 6 doubles (re, im interleaved), e.g. for ncclAllReduce */
void* rpx_detector_field_device(rpx_detector* det);
uint64_t rpx_detector_npoints(const rpx_detector* det);
/* gausslets summed since the last reset / device time (ms, CUDA events) of the summation kernels since then */
uint64_t rpx_detector_modes(const rpx_detector* det);
double rpx_detector_ms(rpx_ctx* ctx, rpx_detector* det);
void rpx_detector_free(rpx_ctx* ctx, rpx_detector* det);

/* ---------------------------------------------------------- streamed trace with device-side consumers
 * trace_rays (core/tracer.py:9-47) for sources whose generations cannot be kept or shipped: the BASELINE
 * configs at 1e8 - 1e9 rays (one generation of 1e9 gausslets is 668 GB).  The source -- host memory, or
 * device memory with RPX_CONSUME_SOURCE_ON_DEVICE -- is cut into contiguous chunks; every chunk is traced
 * through all its generations on the device (the upload of the next chunk overlaps), handed to the
 * consumers below, and freed.  What survives a chunk is what the reference's post-trace consumers keep:
 *   - len(traced_rays[g]) and Face.count, summed over chunks (always);
 *   - RPX_CONSUME_TERMINAL: the terminal rays (rpx_select_terminal) of every generation;
 *   - RPX_CONSUME_CAPTURE: the rays crossing the capture plane set with rpx_capture_scene_set
 *     (select_ray_intersections / select_gausslet_intersections, ctracer.pyx:1981-2058);
 *   - RPX_CONSUME_FIELD (gausslets, needs RPX_CONSUME_CAPTURE): the captured gausslets summed into
 *     `detector` (cfields.pyx:51-117), then dropped unless captured_capacity asks to keep them.
 * Kept collections stay on the device (*terminal / *captured, to download, export or feed onwards) in
 * (chunk, generation, ray) order with parent_idx numbered globally as in rpx_trace_streamed;
 * per_chunk_* (optional, n_chunks x max_gens entries, row-major) give the piece sizes needed to restore
 * the reference's (generation, ray) order.  A capacity of 0 counts without keeping; a selection larger
 * than a non-zero capacity is RPX_ERR_NOMEM.                                                            */
#define RPX_CONSUME_SOURCE_ON_DEVICE 1u
#define RPX_CONSUME_TERMINAL 2u       /* rays that hit nothing ...                       */
#define RPX_CONSUME_CAPTURE 4u
#define RPX_CONSUME_FIELD 8u
#define RPX_CONSUME_MAX_GENS 256

typedef struct rpx_consume_opts {
    uint64_t chunk_rays;           /* source rays per chunk; 0 = default (2^20 gausslets / 2^22 rays)   */
    uint32_t flags;                /* RPX_CONSUME_*                                                      */
    int32_t max_gens;              /* row length of per_chunk_* (<= RPX_CONSUME_MAX_GENS)                */
    const uint8_t* terminal_faces; /* ... and rays ending on these faces (n_traced_faces bytes or NULL)  */
    uint64_t terminal_capacity;    /* records of terminal rays to keep on the device                     */
    uint64_t captured_capacity;    /* records of captured rays to keep on the device                     */
    rpx_detector* detector;        /* RPX_CONSUME_FIELD                                                  */
    uint64_t* per_chunk_terminal;  /* optional outputs, n_chunks x max_gens                              */
    uint64_t* per_chunk_captured;
} rpx_consume_opts;

typedef struct rpx_consume_result {
    int32_t n_gens;                /* len(traced_rays)                                                   */
    int32_t n_chunks;
    uint64_t counts[RPX_CONSUME_MAX_GENS]; /* len(traced_rays[g])                                        */
    uint64_t n_terminal, n_captured;       /* rays selected (kept or not)                                */
    rpx_rays* terminal;            /* kept terminal rays (NULL when terminal_capacity == 0)              */
    rpx_rays* captured;            /* kept captured rays (NULL when captured_capacity == 0)              */
    double device_ms;              /* first to last device operation of the call on the tracing stream   */
    double trace_ms;               /* summed device time of the generation loops                         */
    double shade_ms, intersect_ms; /* summed device time per kernel family ...                           */
    uint64_t shade_launches, intersect_launches; /* ... and their launch counts                          */
    uint64_t launches;             /* every kernel launched by this call                                 */
} rpx_consume_result;

int rpx_trace_consume(rpx_ctx* ctx, const void* rays_aos, uint64_t n, int is_gausslet, double max_length,
                      int recursion_limit, const rpx_consume_opts* opts, uint32_t* face_counts,
                      rpx_consume_result* result);

/* ---------------------------------------------------------- unit entry points
 * Batch evaluation of ONE device function over host arrays -- the GPU counterpart of the
 * reference's Python-callable test wrappers ("mostly for testing",
 * doc/source/creating_new_optics.rst:89-93).  Not on the trace path.                  */
/* Face.intersect(p1, p2, is_base_ray) (ctracer.pyx:1769-1777): p1/p2 are n x 3 local points */
int rpx_unit_face_intersect(rpx_ctx* ctx, int face, const double* p1, const double* p2, uint64_t n,
                            int is_base_ray, double* out_dist);
/* FaceList.compute_orientation(face, point) (ctracer.pyx:1955-1964): n x 3 GLOBAL points.  The
 * reference's wrapper also takes a `piece`; this entry evaluates piece 0 (for RPX_FACE_MESH: the facet
 * normal of cell 0 -- the trace itself uses the piece that was hit).                              */
int rpx_unit_face_normal(rpx_ctx* ctx, int face, const double* points, uint64_t n,
                         double* out_normal, double* out_tangent);
/* InterfaceMaterial.eval_child_ray(ray, idx, point, normal, tangent, new_rays)
 * (ctracer.pyx:1620-1632): children of ray i land in out_aos_2n[2*i .. 2*i+count[i])   */
int rpx_unit_material_eval(rpx_ctx* ctx, int material, const void* rays_aos, uint64_t n,
                           const double* point, const double* normal, const double* tangent,
                           void* out_aos_2n, uint32_t* out_counts);
/* Distortion.z_offset / z_offset_and_gradient (ctracer.pyx:1703-1729): out_grad is n x 3
 * (dz/dx, dz/dy, z)                                                                      */
int rpx_unit_distortion(rpx_ctx* ctx, int distortion, const double* x, const double* y, uint64_t n,
                        double* out_z, double* out_grad);

#ifdef __cplusplus
}
#endif
#endif /* RPX_H_ */
