"""Drop-in replacement for ``raypier.core.tracer`` (raypier/core/tracer.py:9-99 of the
reference): same function name, arguments, return value and side effects, but the
generation loop runs on the GPU through librpx.

    from raypier_optics_b200.core.tracer import trace_rays     # instead of raypier.core.tracer

``input_rays`` / ``face_lists`` may be this package's host mirrors or genuine
``raypier.core`` objects (RayCollection / GaussletCollection, FaceList).
"""
import numpy as np

from .. import _abi as A
from ..engine import get_engine
from ..scene import Scene


def _prepare_faces(input_rays, face_lists, max_length):
    """The per-trace set-up of trace_rays (core/tracer.py:22-37) and the per-generation
    ``fs.sync_transforms()`` of trace_segment / trace_gausslet (ctracer.pyx:2180-2181)."""
    input_rays.reset_length(max_length)
    wavelengths = np.asarray(input_rays.wavelengths)
    face_lists = list(face_lists)
    all_faces = [f for fs in face_lists for f in fs.faces]
    for i, f in enumerate(all_faces):
        f.idx = i
        f.count = 0
        f.update()
        f.material.wavelengths = wavelengths
        f.max_length = max_length
    for fs in face_lists:
        fs.sync_transforms()
    return wavelengths, face_lists, all_faces


def _records_of(input_rays):
    """The packed records of a collection, read-only for the upload.  Host mirrors lend their array (after a
    first trace that is the page-locked block generation 0 came back in, so the next upload of the same
    collection runs at PCIe speed); genuine ``raypier.core`` collections only offer ``copy_as_array``."""
    if hasattr(input_rays, "_assign_array"):
        return input_rays._data
    return input_rays.copy_as_array()


def _assign_in_place(input_rays, arr, ref_array):
    """Overwrite the records of ``input_rays`` with ``arr`` WITHOUT replacing the object: the
    reference mutates its parent collection in place (``length`` / ``end_face_idx`` write-back,
    ctracer.pyx:2086-2087, 1900-1903; gausslets also the six parabasal lengths, :2371) and callers
    rely on ``traced_rays[0] is input_rays``.  Host mirrors re-point their array;
    genuine ``raypier.core`` collections own malloc'd memory with no bulk setter, so their records
    are replaced through the collection's own C-speed list API."""
    if hasattr(input_rays, "_assign_array"):
        input_rays._assign_array(arr)
        return
    cls = type(input_rays)
    fresh = cls.from_array(ref_array(arr))
    input_rays.clear_ray_list()
    if hasattr(input_rays, "extend"):          # GaussletCollection.extend: one memcpy (ctracer.pyx:1294-1302)
        input_rays.extend(fresh)
    else:                                      # RayCollection: Cython loops over Ray objects (:1030-1046, 743-753)
        input_rays.add_ray_list(fresh.get_ray_list())


def _wrap_generations(input_rays, arrays, wavelengths):
    """Return the reference's container classes around the device results.
    ``traced_rays[0] is input_rays`` (mutated in place), each later generation's
    ``.parent`` is the previous one."""
    cls = type(input_rays)
    native = getattr(cls, "_dtype", None) is not None  # this package's host mirrors

    ref_array = _ref_array_for(cls)
    out = []
    prev = None
    for g, arr in enumerate(arrays):
        if g == 0:
            _assign_in_place(input_rays, arr, ref_array)
            rc = input_rays
        else:
            if native:  # the array came fresh from the device: adopt it, no second copy
                rc = cls(0)
                rc._assign_array(arr)
            else:
                rc = cls.from_array(ref_array(arr))
            rc.wavelengths = wavelengths
            rc.parent = prev
        out.append(rc)
        prev = rc
    return out


def _ref_array_for(cls):
    """numpy dtype bridge for genuine reference collections (their from_array checks their own dtype)."""
    if getattr(cls, "_dtype", None) is not None:
        return lambda arr: arr
    mod = __import__(cls.__module__, fromlist=["ray_dtype"])

    def conv(arr):
        # from_array checks the IDENTITY of the reference's own dtype object (ctracer.pyx:1150): a view with
        # that very object passes it without copying (the layouts are byte-identical, tests/test_abi.py)
        ref_dtype = mod.gausslet_dtype if arr.dtype.itemsize == A.gausslet_dtype.itemsize else mod.ray_dtype
        return np.ascontiguousarray(arr).view(ref_dtype)
    return conv


def _trace_with_decomposition(eng, input_rays, native, all_faces, wavelengths, recursion_limit, max_length):
    """The generation loop of trace_rays (core/tracer.py:39-45) when a face carries a decomposition
    material (ResampleGaussletMaterial, cmaterials.pyx:1766-1831): one device step per generation
    (``rpx_trace_step``), and between two steps what trace_gausslet_c does after its ray loop
    (ctracer.pyx:2274-2280): for every decomposition face that was hit, hand the gausslets it captured
    -- copies of the parents taken when they hit, i.e. with the base ray's length / end_face_idx written
    back and the parabasal lengths still at max_length -- to the material's Python callback, append what
    it returns to the new generation AFTER the regular children, zero that face's count, and reset every
    length of the new generation to max_length.  Only generations in which a decomposition face was hit
    make the round trip through the host."""
    cls = type(input_rays)
    to_ref = _ref_array_for(cls)
    n_faces = len(all_faces)
    decomp = [j for j, f in enumerate(all_faces) if f.material.is_decomp_material()]
    fc = np.zeros(max(n_faces, 1), dtype=np.uint32)
    arrays = []
    cur = eng.upload(native)
    count = 0
    try:
        while len(cur) > 0 and count < recursion_limit:
            fc_gen = np.zeros_like(fc)
            children = eng.trace_step(cur, max_length, fc_gen)
            arr = eng.download(cur)
            cur.free()
            cur = children
            arrays.append(arr)
            fc += fc_gen
            extra = []
            for j in decomp:
                if not fc_gen[j]:
                    continue
                mat = all_faces[j].material
                base = arr['base_ray']
                sel = (base['end_face_idx'] == j) & ((base['ray_type_id'] & A.GAUSSLET) != 0)
                cap = arr[sel].copy()
                cap['para_rays']['length'] = max_length  # captured before trace_parabasal_rays wrote them
                mat.captured_rays.clear_ray_list()
                mat.captured_rays.extend(cls.from_array(to_ref(cap)))
                mat.captured_rays.wavelengths = wavelengths
                mat.capture_count += int(sel.sum())
                out = mat.eval_func(mat.captured_rays)
                mat.captured_rays.clear_ray_list()
                new = np.ascontiguousarray(out.copy_as_array()).view(A.gausslet_dtype).copy()
                extra.append(new)
                fc[j] = 0  # `face.count = 0` (ctracer.pyx:2277)
            if extra:
                kids = eng.download(cur)
                cur.free()
                allk = np.concatenate([kids] + extra)
                allk['base_ray']['length'] = max_length  # new_gausslets.reset_length_c (:2280)
                allk['para_rays']['length'] = max_length
                cur = eng.upload(allk)
            count += 1
    finally:
        cur.free()
    return arrays, fc[:n_faces]


def trace_rays(input_rays, face_lists, recursion_limit=100, max_length=100.0, device=0):
    """Core ray-tracing routine: traces a RayCollection or GaussletCollection
    non-sequentially through the given list of FaceList objects.

    returns - (traced_rays, all_faces), as raypier.core.tracer.trace_rays does: the list of
    ray generations and the flat face list that ``end_face_idx`` indexes (``face.count`` =
    hits in this trace).
    """
    wavelengths, face_lists, all_faces = _prepare_faces(input_rays, face_lists, max_length)
    eng = get_engine(device)
    scene = Scene(face_lists, wavelengths)
    eng.set_scene(scene)
    rays = _records_of(input_rays)
    native = np.ascontiguousarray(rays).view(
        A.gausslet_dtype if rays.dtype.itemsize == A.gausslet_dtype.itemsize else A.ray_dtype)
    if native.dtype == A.gausslet_dtype and any(f.material.is_decomp_material() for f in all_faces):
        # (plain rays never trigger a decomposition: eval_child_ray_c captures gausslets only and
        # trace_segment_c has no callback step -- for them the face simply absorbs)
        arrays, counts = _trace_with_decomposition(eng, input_rays, native, all_faces, wavelengths, recursion_limit,
                                                   max_length)
        for f, c in zip(all_faces, counts):
            f.count = int(c)
        trace_rays.last_device_ms = None
        return _wrap_generations(input_rays, arrays, wavelengths), all_faces
    res = eng.trace(native, max_length, recursion_limit)
    try:
        arrays = res.generations()
        for f, c in zip(all_faces, res.face_counts):
            f.count = int(c)
    finally:
        res.free()
    traced = _wrap_generations(input_rays, arrays, wavelengths)
    trace_rays.last_device_ms = res.device_ms
    return traced, all_faces


def sequence_face_indices(face_sequence):
    """Global face index of every step of a face sequence, numbered the way
    trace_ray_sequence numbers them (core/tracer.py:71-80): ``all_faces`` chains the faces of
    EVERY FaceList entry of the sequence (a FaceList listed twice contributes twice) and
    ``f.idx`` keeps the position of its LAST occurrence."""
    face_lists = [fl for fl, _ in face_sequence]
    all_faces = [f for fs in face_lists for f in fs.faces]
    last_idx = {}
    for i, f in enumerate(all_faces):
        last_idx[id(f)] = i
    return [last_idx[id(fl.faces[fidx])] for fl, fidx in face_sequence]


def trace_ray_sequence(input_rays, face_sequence, recursion_limit=100, max_length=100.0, device=0):
    """Sequential ray-trace (raypier/core/tracer.py:50-99): ``face_sequence`` is a list of
    ``(FaceList, face_idx)``; step s intersects the rays of generation s with that one face only.

    returns - (traced_rays, all_faces) like the reference: ``traced_rays[0]`` is the input, the
    last generation is returned as the materials left it (not intersected with anything).
    """
    face_sequence = list(face_sequence)
    face_lists = [fl for fl, _ in face_sequence]
    wavelengths, face_lists, all_faces = _prepare_faces(input_rays, face_lists, max_length)
    seq = sequence_face_indices(face_sequence)
    eng = get_engine(device)
    scene = Scene(face_lists, wavelengths)
    eng.set_scene(scene)
    rays = _records_of(input_rays)
    native = np.ascontiguousarray(rays).view(
        A.gausslet_dtype if rays.dtype.itemsize == A.gausslet_dtype.itemsize else A.ray_dtype)
    res = eng.trace_sequence(native, seq, max_length, recursion_limit)
    try:
        arrays = res.generations()
        for f, c in zip(all_faces, res.face_counts):
            f.count = int(c)
    finally:
        res.free()
    if not arrays:  # empty input: the reference still returns [input_rays] (core/tracer.py:66)
        return [input_rays], all_faces
    return _wrap_generations(input_rays, arrays, wavelengths), all_faces
