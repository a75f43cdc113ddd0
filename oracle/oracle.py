"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/librpx_oracle.so
(the plain-C CPU restatement of the reference trace, oracle/rpx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from raypier_optics_b200 import _abi as A  # layout only (dtypes / rpx_scene struct)

_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "librpx_oracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "librpx_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, d, i, u32, u64 = C.c_void_p, C.c_double, C.c_int, C.c_uint32, C.c_uint64
        L.rpxo_trace_segment.restype = u64
        L.rpxo_trace_segment.argtypes = [vp, vp, u64, d, vp, vp]
        L.rpxo_trace_gausslet.restype = u64
        L.rpxo_trace_gausslet.argtypes = [vp, vp, u64, d, vp, vp]
        L.rpxo_trace_segment_ex.restype = u64
        L.rpxo_trace_segment_ex.argtypes = [vp, vp, u64, d, vp, vp, i]
        L.rpxo_trace_gausslet_ex.restype = u64
        L.rpxo_trace_gausslet_ex.argtypes = [vp, vp, u64, d, vp, vp, i]
        L.rpxo_capture_rays.restype = u64
        L.rpxo_capture_rays.argtypes = [vp, vp, vp, u64, u32, vp]
        L.rpxo_capture_gausslets.restype = u64
        L.rpxo_capture_gausslets.argtypes = [vp, vp, vp, u64, u32, vp]
        L.rpxo_evaluate_neighbours_gc.argtypes = [vp, u64, vp, vp, vp, vp]
        L.rpxo_evaluate_modes.argtypes = [vp, vp, vp, vp, u64, i, d, vp]
        L.rpxo_sum_gaussian_modes.argtypes = [vp, u64, vp, vp, vp, u64, d, vp]
        L.rpxo_project_to_sphere.restype = u64
        L.rpxo_project_to_sphere.argtypes = [vp, u64, vp, d, vp]
        L.rpxo_evaluate_neighbours.restype = u64
        L.rpxo_evaluate_neighbours.argtypes = [vp, u64, vp, i, vp, vp, vp, vp, vp]
        L.rpxo_face_intersect.restype = d
        L.rpxo_face_intersect.argtypes = [vp, i, vp, vp, i]
        L.rpxo_face_normal.argtypes = [vp, i, vp, vp]
        L.rpxo_orientation.argtypes = [vp, i, vp, vp, vp]
        L.rpxo_convert_to_sp.argtypes = [vp, vp, vp]
        L.rpxo_material_eval.restype = i
        L.rpxo_material_eval.argtypes = [vp, i, vp, u32, vp, vp, vp, vp]
        L.rpxo_material_eval_para.argtypes = [vp, i, vp, vp, vp, vp, vp, u32, vp]
        L.rpxo_distortion_z.restype = d
        L.rpxo_distortion_z.argtypes = [vp, i, d, d]
        L.rpxo_distortion_zgrad.argtypes = [vp, i, d, d, vp]
        L.rpxo_shape_inside.restype = i
        L.rpxo_shape_inside.argtypes = [vp, i, d, d]
        L.rpxo_implicit_eval.restype = d
        L.rpxo_implicit_eval.argtypes = [vp, i, i, vp]
        for name in ("rpxo_zernike_R", "rpxo_zernike_Rprime", "rpxo_zernike_R_over_r"):
            f = getattr(L, name)
            f.restype = d
            f.argtypes = [d, i, i, i, vp, i]
        assert L.rpxo_sizeof_ray() == 188 and L.rpxo_sizeof_gausslet() == 668
        _LIB = L
    return _LIB


def _v3(v):
    return np.ascontiguousarray(v, dtype=np.double).reshape(3)


def trace_generation(scene, rays, max_length, face_counts=None, only_face=-1):
    """One generation (trace_segment_c / trace_gausslet_c).  ``rays`` (ray_dtype or
    gausslet_dtype array) is mutated in place like the reference mutates the parent
    collection; returns the child array."""
    L = lib()
    rays_c = rays
    assert rays_c.flags.c_contiguous
    n = rays_c.shape[0]
    out = np.zeros(max(2 * n, 1), dtype=rays_c.dtype)
    fc = face_counts.ctypes.data if face_counts is not None else None
    if rays_c.dtype == A.ray_dtype:
        n_out = L.rpxo_trace_segment_ex(scene.byref_ptr(), rays_c.ctypes.data, n, float(max_length),
                                        out.ctypes.data, fc, int(only_face))
    elif rays_c.dtype == A.gausslet_dtype:
        n_out = L.rpxo_trace_gausslet_ex(scene.byref_ptr(), rays_c.ctypes.data, n, float(max_length),
                                         out.ctypes.data, fc, int(only_face))
    else:
        raise TypeError("rays must be ray_dtype or gausslet_dtype")
    return out[:n_out].copy()


class OracleScene:
    """Adapter giving the oracle a stable pointer to a flattened Scene."""

    def __init__(self, scene):
        self.scene = scene

    def byref_ptr(self):
        return C.cast(C.pointer(self.scene.c_scene), C.c_void_p)


def trace_rays(scene, input_rays, recursion_limit=100, max_length=100.0, decomp=None):
    """The generation loop of raypier.core.tracer.trace_rays (core/tracer.py:9-47) on
    numpy arrays.  Returns (list of generation arrays, face_counts).

    ``decomp``: {face index: callable(gausslet array) -> gausslet array} for faces that carry a
    ResampleGaussletMaterial (cmaterials.pyx:1766-1831; the C oracle treats the face as an absorber, which
    is all eval_child_ray_c does to the ray itself).  After the ray loop of a generation, for each such
    face that was hit (ctracer.pyx:2274-2278): the gausslets captured by it -- copies taken when they hit,
    base ray written back, parabasal lengths still max_length -- go through the callable, the result is
    appended to the new generation, the face's count is zeroed; then every length of the new generation
    is reset to max_length (:2280)."""
    osc = scene if isinstance(scene, OracleScene) else OracleScene(scene)
    rays = np.ascontiguousarray(input_rays).copy()
    is_g = rays.dtype == A.gausslet_dtype
    # input_rays.reset_length(max_length), core/tracer.py:22
    if is_g:
        rays['base_ray']['length'] = max_length
        rays['para_rays']['length'] = max_length
    else:
        rays['length'] = max_length
    counts = np.zeros(max(scene.c_scene.n_traced_faces, 1), dtype=np.uint32)
    traced = []
    count = 0
    while rays.shape[0] > 0 and count < recursion_limit:
        traced.append(rays)
        before = counts.copy()
        parents = rays
        rays = trace_generation(osc, rays, max_length, counts)
        if decomp and is_g:
            extra = []
            ef = parents['base_ray']['end_face_idx']
            for j in sorted(decomp):
                if counts[j] == before[j]:
                    continue
                sel = (ef == j) & ((parents['base_ray']['ray_type_id'] & A.GAUSSLET) != 0)
                cap = parents[sel].copy()
                cap['para_rays']['length'] = max_length
                extra.append(np.ascontiguousarray(decomp[j](cap)).view(A.gausslet_dtype))
                counts[j] = 0
            if extra:
                rays = np.concatenate([rays] + extra)
                rays['base_ray']['length'] = max_length
                rays['para_rays']['length'] = max_length
        count += 1
    return traced, counts[:scene.c_scene.n_traced_faces]


def trace_rays_blocks(scene, input_rays, recursion_limit=100, max_length=100.0, block=65536, workers=None):
    """The same trace as ``trace_rays`` for BASELINE-size sources: the source is cut into
    contiguous blocks that are traced independently on ``workers`` host threads (the C oracle
    releases the GIL and keeps no shared state).  Yields ``(lo, hi, generations, face_counts)``
    per block as blocks complete, in no particular order; a ray never interacts with another
    ray and children are appended in parent order (ctracer.pyx:2084-2117), so generation g of
    the whole trace is the blocks' generation g concatenated in source order with
    ``parent_idx`` shifted by the generation g-1 rays of earlier blocks."""
    import concurrent.futures as cf
    import os
    osc = scene if isinstance(scene, OracleScene) else OracleScene(scene)
    sc = osc.scene
    n = input_rays.shape[0]
    bounds = [(lo, min(lo + block, n)) for lo in range(0, n, block)]
    workers = workers or min(len(bounds), os.cpu_count() or 1) or 1

    def one(b):
        lo, hi = b
        gens, fc = trace_rays(sc, input_rays[lo:hi], recursion_limit, max_length)
        return lo, hi, gens, fc

    with cf.ThreadPoolExecutor(workers) as ex:
        # keep at most 2 * workers blocks in flight so that the finished ones do not pile up
        pending, it = set(), iter(bounds)
        for b in it:
            pending.add(ex.submit(one, b))
            if len(pending) >= 2 * workers:
                break
        while pending:
            done, pending = cf.wait(pending, return_when=cf.FIRST_COMPLETED)
            for f in done:
                yield f.result()
                nxt = next(it, None)
                if nxt is not None:
                    pending.add(ex.submit(one, nxt))


def trace_ray_sequence(scene, input_rays, face_seq, recursion_limit=100, max_length=100.0):
    """raypier.core.tracer.trace_ray_sequence (core/tracer.py:50-99) on numpy arrays;
    ``face_seq`` holds the global face index of every step."""
    osc = scene if isinstance(scene, OracleScene) else OracleScene(scene)
    rays = np.ascontiguousarray(input_rays).copy()
    if rays.dtype == A.gausslet_dtype:
        rays['base_ray']['length'] = max_length
        rays['para_rays']['length'] = max_length
    else:
        rays['length'] = max_length
    counts = np.zeros(max(scene.c_scene.n_traced_faces, 1), dtype=np.uint32)
    traced = [rays]
    count = 0
    for fidx in face_seq:
        rays = trace_generation(osc, rays, max_length, counts, only_face=int(fidx))
        if (count > recursion_limit) or rays.shape[0] == 0:
            break
        traced.append(rays)
        count += 1
    return traced, counts[:scene.c_scene.n_traced_faces]


def reference_trace_ray_sequence(core, input_rays, face_sequence, recursion_limit=100, max_length=100.0):
    """trace_ray_sequence (core/tracer.py:50-99) over the REAL reference kernels."""
    ct = core.ctracer
    input_rays.reset_length(max_length)
    traced_rays = [input_rays]
    face_lists = [fl for fl, fidx in face_sequence]
    face_idx_list = [fidx for fl, fidx in face_sequence]
    trace_func = (ct.trace_one_face_segment if isinstance(input_rays, ct.RayCollection)
                  else ct.trace_one_face_gausslet)
    count = 0
    wavelengths = np.asarray(input_rays.wavelengths)
    all_faces = [f for fs in face_lists for f in fs.faces]
    for i, f in enumerate(all_faces):
        f.idx = i
        f.count = 0
        f.update()
        f.material.wavelengths = wavelengths
        f.max_length = max_length
    decomp_faces = [f for f in all_faces if f.material.is_decomp_material()]
    rays = input_rays
    for next_face_list, next_face_idx in zip(face_lists, face_idx_list):
        rays = trace_func(rays, next_face_list, next_face_idx, all_faces, max_length=max_length,
                          decomp_faces=decomp_faces)
        if (count > recursion_limit) or (rays.n_rays == 0):
            break
        traced_rays.append(rays)
        count += 1
    return traced_rays, all_faces


def select_intersections(capture_scene, collections, wavelength_lists, face_ids=None):
    """select_ray_intersections / select_gausslet_intersections (ctracer.pyx:1981-2058) on numpy
    arrays: ``capture_scene`` is the flattened capture FaceList, ``collections`` the ray arrays
    (one per generation), ``wavelength_lists`` their wavelength tables.  Returns
    (captured array, reduced wavelengths, rays captured per collection)."""
    L = lib()
    osc = capture_scene if isinstance(capture_scene, OracleScene) else OracleScene(capture_scene)
    ids = None if face_ids is None else np.ascontiguousarray(face_ids, dtype=np.uint32)
    parts, counts = [], []
    wl_offset = 0
    for rays, wls in zip(collections, wavelength_lists):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(max(rays.shape[0], 1), dtype=rays.dtype)
        fn = L.rpxo_capture_gausslets if rays.dtype == A.gausslet_dtype else L.rpxo_capture_rays
        n = fn(osc.byref_ptr(), None if ids is None else ids.ctypes.data, rays.ctypes.data, rays.shape[0],
               wl_offset, out.ctypes.data)
        parts.append(out[:n])
        counts.append(int(n))
        wl_offset += len(wls)
    captured = np.concatenate(parts) if parts else np.zeros(0, dtype=A.ray_dtype)
    # np.unique re-mapping of the offset wavelength indices, ctracer.pyx:2011-2016
    reduced, inverse = np.unique(np.concatenate([np.asarray(w, dtype=np.double) for w in wavelength_lists]),
                                 return_inverse=True)
    wl = captured['base_ray']['wavelength_idx'] if captured.dtype == A.gausslet_dtype else captured['wavelength_idx']
    wl[:] = inverse[wl].astype(np.uint32)
    return captured, reduced, counts


def reference_select_intersections(core, face_list, collections):
    """The REAL select_ray_intersections / select_gausslet_intersections of the reference."""
    ct = core.ctracer
    face_list.sync_transforms()
    is_g = isinstance(collections[0], ct.GaussletCollection)
    fn = ct.select_gausslet_intersections if is_g else ct.select_ray_intersections
    rc = fn(face_list, list(collections))
    arr = rc.copy_as_array()
    return arr.view(A.gausslet_dtype if is_g else A.ray_dtype), np.asarray(rc.wavelengths), rc


# ---- E-field summation: raypier/core/cfields.pyx + core/fields.py front end -------
def evaluate_neighbours_gc(gausslets):
    """core/fields.py:114-137 -> (x, y, dx, dy), each n x 6."""
    L = lib()
    g = np.ascontiguousarray(gausslets)
    assert g.dtype == A.gausslet_dtype
    n = g.shape[0]
    x, y, dx, dy = (np.zeros((n, 6)) for _ in range(4))
    L.rpxo_evaluate_neighbours_gc(g.ctypes.data, n, x.ctypes.data, y.ctypes.data, dx.ctypes.data, dy.ctypes.data)
    return x, y, dx, dy


def evaluate_modes(x, y, dx, dy, blending=1.0):
    """cfields.evaluate_modes (cfields.pyx:217-228) -> n x 3 complex."""
    L = lib()
    x, y, dx, dy = (np.ascontiguousarray(a, dtype=np.double) for a in (x, y, dx, dy))
    n, row = x.shape
    out = np.zeros((n, 3), dtype=np.complex128)
    L.rpxo_evaluate_modes(x.ctypes.data, y.ctypes.data, dx.ctypes.data, dy.ctypes.data, n, row, float(blending),
                          out.ctypes.data)
    return out


def sum_gaussian_modes(rays, modes, wavelengths, points, time_ps=0.0):
    """cfields.sum_gaussian_modes (cfields.pyx:51-115) -> npt x 3 complex."""
    L = lib()
    rays = np.ascontiguousarray(rays)
    assert rays.dtype == A.ray_dtype
    modes = np.ascontiguousarray(modes, dtype=np.complex128)
    wl = np.ascontiguousarray(wavelengths, dtype=np.double)
    pts = np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)
    out = np.zeros((pts.shape[0], 3), dtype=np.complex128)
    L.rpxo_sum_gaussian_modes(rays.ctypes.data, rays.shape[0], modes.ctypes.data, wl.ctypes.data, pts.ctypes.data,
                              pts.shape[0], float(time_ps), out.ctypes.data)
    return out


def eval_Efield_from_gausslets(gausslets, points, wavelengths, blending=1.0, time_ps=0.0):
    """core/fields.py:252-277 on numpy arrays."""
    g = np.ascontiguousarray(gausslets)
    x, y, dx, dy = evaluate_neighbours_gc(g)
    modes = evaluate_modes(x, y, dx, dy, blending)
    return sum_gaussian_modes(np.ascontiguousarray(g['base_ray']), modes, wavelengths, points, time_ps)


def project_to_sphere(rays, centre=(0, 0, 0), radius=10.0):
    """core/fields.py:50-77 -> the selected rays, moved to the sphere (a copy, like rays[selector])."""
    L = lib()
    r = np.array(rays, dtype=A.ray_dtype, copy=True)
    c = np.ascontiguousarray(centre, dtype=np.double).reshape(3)
    sel = np.zeros(len(r), dtype=np.uint8)
    L.rpxo_project_to_sphere(r.ctypes.data, len(r), c.ctypes.data, float(radius), sel.ctypes.data)
    return r[sel.astype(bool)]


def evaluate_neighbours(rays, neighbours_idx):
    """core/fields.py:80-111 -> (rays[mask], x, y, dx, dy)."""
    L = lib()
    r = np.ascontiguousarray(rays)
    assert r.dtype == A.ray_dtype
    nb = np.ascontiguousarray(neighbours_idx, dtype=np.int32)
    n, row = nb.shape
    if n != len(r):  # numpy: boolean index did not match indexed array
        raise IndexError("neighbours_idx has %d rows for %d rays" % (n, len(r)))
    if nb.max(initial=-1) >= len(r):
        raise IndexError("neighbour index out of bounds")
    x, y, dx, dy = (np.zeros((n, row)) for _ in range(4))
    mask = np.zeros(n, dtype=np.uint8)
    k = L.rpxo_evaluate_neighbours(r.ctypes.data, n, nb.ctypes.data, row, x.ctypes.data, y.ctypes.data, dx.ctypes.data,
                                   dy.ctypes.data, mask.ctypes.data)
    return r[mask.astype(bool)], x[:k], y[:k], dx[:k], dy[:k]


def eval_Efield_from_rays(rays, neighbours_idx, points, wavelengths, blending=1.0, time_ps=0.0,
                          exit_pupil_offset=0.0, exit_pupil_centre=(0.0, 0.0, 0.0)):
    """core/fields.py:206-229 on numpy arrays (rays = ray_collection.copy_as_array(), neighbours_idx =
    ray_collection.neighbours)."""
    rays = np.ascontiguousarray(rays)
    projected = project_to_sphere(rays, exit_pupil_centre, exit_pupil_offset) if exit_pupil_offset else rays
    kept, x, y, dx, dy = evaluate_neighbours(projected, neighbours_idx)
    modes = evaluate_modes(x, y, dx, dy, blending)
    return sum_gaussian_modes(np.ascontiguousarray(kept), modes, wavelengths, points, time_ps)


def reference_fields(core):
    """The reference's raypier/core/fields.py, imported in place from /root/reference (it is pure
    Python over the compiled cfields of oracle/_ref); None where /root/reference is absent."""
    import importlib
    import importlib.util
    name = core.__name__ + ".fields"
    if name in sys.modules:
        return sys.modules[name]
    src = "/root/reference/raypier/core"
    if not os.path.exists(os.path.join(src, "fields.py")):
        return None
    try:
        importlib.import_module(core.__name__ + ".cfields")
        for mod in ("utils", "fields"):
            full = core.__name__ + "." + mod
            if full in sys.modules:
                continue
            spec = importlib.util.spec_from_file_location(full, os.path.join(src, mod + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        return sys.modules[name]
    except Exception:
        sys.modules.pop(core.__name__ + ".fields", None)
        return None


# ---- unit entry points (for pinning against the reference's own KATs) -------
# NOTE: every array passed by pointer is bound to a local first, so it outlives the call.
def face_intersect(scene, face_idx, p1, p2, is_base_ray=1):
    osc = OracleScene(scene)
    a, b = _v3(p1), _v3(p2)
    return lib().rpxo_face_intersect(osc.byref_ptr(), face_idx, a.ctypes.data, b.ctypes.data,
                                     int(is_base_ray))


def face_normal(scene, face_idx, p):
    osc = OracleScene(scene)
    a, out = _v3(p), np.zeros(3)
    lib().rpxo_face_normal(osc.byref_ptr(), face_idx, a.ctypes.data, out.ctypes.data)
    return out


def orientation(scene, face_idx, point):
    osc = OracleScene(scene)
    a, n, t = _v3(point), np.zeros(3), np.zeros(3)
    lib().rpxo_orientation(osc.byref_ptr(), face_idx, a.ctypes.data, n.ctypes.data, t.ctypes.data)
    return n, t


def convert_to_sp(ray, normal):
    r = np.ascontiguousarray(ray, dtype=A.ray_dtype).reshape(1).copy()
    nn = _v3(normal)
    out = np.zeros(1, dtype=A.ray_dtype)
    lib().rpxo_convert_to_sp(r.ctypes.data, nn.ctypes.data, out.ctypes.data)
    return out[0]


def material_eval(scene, mat_idx, ray, idx, point, normal, tangent=(1.0, 0.0, 0.0)):
    osc = OracleScene(scene)
    r = np.ascontiguousarray(ray, dtype=A.ray_dtype).reshape(1).copy()
    p, nn, tt = _v3(point), _v3(normal), _v3(tangent)
    out = np.zeros(2, dtype=A.ray_dtype)
    n = lib().rpxo_material_eval(osc.byref_ptr(), mat_idx, r.ctypes.data, int(idx), p.ctypes.data,
                                 nn.ctypes.data, tt.ctypes.data, out.ctypes.data)
    return out[:n].copy()


def material_eval_para(scene, mat_idx, base_ray, direction, point, normal, tangent=(1.0, 0.0, 0.0),
                       ray_type_id=0):
    osc = OracleScene(scene)
    r = np.ascontiguousarray(base_ray, dtype=A.ray_dtype).reshape(1).copy()
    d, p, nn, tt = _v3(direction), _v3(point), _v3(normal), _v3(tangent)
    out = np.zeros(1, dtype=A.para_dtype)
    lib().rpxo_material_eval_para(osc.byref_ptr(), mat_idx, r.ctypes.data, d.ctypes.data, p.ctypes.data,
                                  nn.ctypes.data, tt.ctypes.data, int(ray_type_id), out.ctypes.data)
    return out[0]


def distortion_z(scene, dist_idx, x, y):
    return lib().rpxo_distortion_z(OracleScene(scene).byref_ptr(), dist_idx, float(x), float(y))


def distortion_zgrad(scene, dist_idx, x, y):
    out = np.zeros(3)
    lib().rpxo_distortion_zgrad(OracleScene(scene).byref_ptr(), dist_idx, float(x), float(y),
                                out.ctypes.data)
    return out


def shape_inside(scene, face_idx, x, y):
    return lib().rpxo_shape_inside(OracleScene(scene).byref_ptr(), face_idx, float(x), float(y))


def implicit_eval(scene, off, length, p):
    a = _v3(p)
    return lib().rpxo_implicit_eval(OracleScene(scene).byref_ptr(), off, length, a.ctypes.data)


def zernike(which, r, k, n, m, kmax):
    ws = np.full(3 * kmax, np.nan)
    f = {"R": lib().rpxo_zernike_R, "Rprime": lib().rpxo_zernike_Rprime,
         "R_over_r": lib().rpxo_zernike_R_over_r}[which]
    return f(float(r), int(k), int(n), int(m), ws.ctypes.data, int(kmax))


# ---- the real reference, when it was built here (oracle/_ref) ----------------
def reference_path(flavour="parity"):
    return os.path.join(_HERE, "_ref", flavour)


def import_reference(flavour="parity"):
    """Import the unmodified reference core from oracle/_ref/<flavour> (built by
    oracle/build_ref.sh).  Returns the ``raypier.core`` package or None."""
    path = reference_path(flavour)
    if not os.path.isdir(os.path.join(path, "raypier", "core")):
        return None
    for m in [k for k in sys.modules if k == "raypier" or k.startswith("raypier.")]:
        # a different flavour may already be loaded; extension modules cannot be reloaded
        mod = sys.modules[m]
        f = getattr(mod, "__file__", "") or ""
        if f and not f.startswith(path):
            return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        import importlib
        core = importlib.import_module("raypier.core")
        for name in ("ctracer", "cfaces", "cmaterials", "cshapes", "cdistortions", "cimplicit_surfs"):
            importlib.import_module("raypier.core." + name)
        try:  # triangle-mesh faces (needs PIL at import time: obbtree.pyx:6); optional
            importlib.import_module("raypier.core.obbtree")
            # UV patch faces.  cbezier.pyx:16 does `from numpy import math`, an alias numpy >= 2 (the only
            # numpy in this image) no longer has.  The reference SOURCE is built unmodified; the missing
            # alias is supplied at run time, in this process only, before the module is imported:
            import math
            import numpy
            if not hasattr(numpy, "math"):
                numpy.math = math
            importlib.import_module("raypier.core.cbezier")
        except Exception:
            pass
        return core
    except Exception:
        return None


def reference_trace_rays(core, input_rays, face_lists, recursion_limit=100, max_length=100.0):
    """raypier.core.tracer.trace_rays (core/tracer.py:9-47) driven over the REAL reference
    kernels (ctracer.trace_segment / trace_gausslet from oracle/_ref).  The 40-line Python
    driver is restated here rather than copied into oracle/_ref."""
    ct = core.ctracer
    input_rays.reset_length(max_length)
    traced_rays = []
    trace_func = ct.trace_segment if isinstance(input_rays, ct.RayCollection) else ct.trace_gausslet
    count = 0
    wavelengths = np.asarray(input_rays.wavelengths)
    all_faces = [f for fs in face_lists for f in fs.faces]
    for i, f in enumerate(all_faces):
        f.idx = i
        f.count = 0
        f.update()
        f.material.wavelengths = wavelengths
        f.max_length = max_length
    face_sets = list(face_lists)
    decomp_faces = [f for f in all_faces if f.material.is_decomp_material()]
    rays = input_rays
    while rays.n_rays > 0 and count < recursion_limit:
        traced_rays.append(rays)
        rays = trace_func(rays, face_sets, all_faces, max_length=max_length,
                          decomp_faces=decomp_faces)
        count += 1
    return traced_rays, all_faces


def reference_collection(core, rays, wavelengths):
    """numpy ray_dtype / gausslet_dtype array -> reference RayCollection / GaussletCollection."""
    ct = core.ctracer
    if rays.dtype == A.gausslet_dtype:
        rc = ct.GaussletCollection.from_array(np.ascontiguousarray(rays).view(ct.gausslet_dtype))
    else:
        a = np.empty(rays.shape[0], dtype=ct.ray_dtype)  # from_array asserts dtype identity
        a.view(np.uint8)[:] = np.ascontiguousarray(rays).view(np.uint8)
        rc = ct.RayCollection.from_array(a)
    rc.wavelengths = np.ascontiguousarray(wavelengths, dtype=np.double)
    return rc
