"""Thin object layer over the C ABI: one ``Engine`` = one ``rpx_ctx`` (one GPU)."""
import ctypes as C

import numpy as np

from . import _abi as A
from ._lib import RpxError, load
from ._hostpool import get_pool


class TraceResult(object):
    """Handle on a finished trace (``rpx_result``): generation counts, per-face hit
    counts, device timings; generations are copied to the host on demand."""

    def __init__(self, engine, handle, is_gausslet):
        self._e = engine
        self._h = handle
        self.is_gausslet = bool(is_gausslet)
        L = engine._L
        n = L.rpx_result_n_generations(handle)
        counts = np.zeros(max(n, 1), dtype=np.uint64)
        L.rpx_result_counts(handle, counts.ctypes.data)
        self.counts = [int(c) for c in counts[:n]]
        fc = np.zeros(max(engine.n_traced_faces, 1), dtype=np.uint32)
        L.rpx_result_face_counts(handle, fc.ctypes.data)
        self.face_counts = fc[:engine.n_traced_faces].copy()
        self.device_ms = float(L.rpx_result_device_ms(handle))
        self.launches = int(L.rpx_result_launches(handle))
        self.kernel_ms = {}
        for which, name in ((0, "intersect"), (1, "shade")):
            ms, ln = C.c_double(), C.c_uint64()
            L.rpx_result_kernel_ms(handle, which, C.byref(ms), C.byref(ln))
            self.kernel_ms[name] = (ms.value, int(ln.value))

    @property
    def n_generations(self):
        return len(self.counts)

    @property
    def segments(self):
        """ray-segments of this trace = sum over generations of len(traced_rays[g])."""
        return int(sum(self.counts))

    def generation(self, g, out=None):
        dtype = A.gausslet_dtype if self.is_gausslet else A.ray_dtype
        n = self.counts[g]
        if out is None:
            out = self._e.result_empty(n, dtype)  # pooled page-locked block (see _hostpool.py)
        assert out.dtype == dtype and out.shape[0] >= n and out.flags.c_contiguous
        self._e._check(self._e._L.rpx_result_generation(self._e._ctx, self._h, g, out.ctypes.data,
                                                         out.shape[0]))
        return out[:n]

    def generations(self):
        return [self.generation(g) for g in range(self.n_generations)]

    def device_generations(self):
        """Borrowed ``rpx_rays`` handles of the generations still resident on the device
        (input of ``Engine.capture``); valid until this result is freed."""
        return [self._e._L.rpx_result_rays(self._h, g) for g in range(self.n_generations)]

    def capture(self, wavelengths, out=None):
        """Filter every generation through the engine's capture plane on the device
        (``rpx_capture``) and return ``(captured array, reduced wavelengths, per-generation
        counts)`` -- select_ray_intersections over ``traced_rays`` without shipping the
        generations to the host."""
        return self._e.capture_collections(self.device_generations(),
                                           [wavelengths] * self.n_generations, self.is_gausslet, out=out)

    def free(self):
        if self._h is not None:
            self._e._L.rpx_result_free(self._e._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceRays(object):
    """A device-resident generation (``rpx_rays``)."""

    def __init__(self, engine, handle, is_gausslet):
        self._e, self._h, self.is_gausslet = engine, handle, bool(is_gausslet)

    def __len__(self):
        return int(self._e._L.rpx_rays_count(self._h))

    def free(self):
        if self._h is not None:
            self._e._L.rpx_rays_free(self._e._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FieldModes(object):
    """Prepared Gaussian modes of a set of rays / gausslets on the device (``rpx_field``):
    evaluate the E-field at as many point sets as needed."""

    def __init__(self, engine, handle):
        self._e, self._h = engine, handle

    def __len__(self):
        return int(self._e._L.rpx_field_count(self._h))

    @property
    def modes(self):
        """(A, B, C) per ray, n x 3 complex128."""
        out = np.zeros((len(self), 3), dtype=np.complex128)
        self._e._check(self._e._L.rpx_field_modes(self._e._ctx, self._h, out.ctypes.data))
        return out

    def evaluate(self, points, time_ps=0.0):
        pts = np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)
        out = np.zeros((pts.shape[0], 3), dtype=np.complex128)
        self._e._check(self._e._L.rpx_field_evaluate(self._e._ctx, self._h, pts.ctypes.data, pts.shape[0],
                                                     float(time_ps), out.ctypes.data))
        return out

    def evaluate_device(self, d_points, npt, d_out, time_ps=0.0):
        """Device pointers (e.g. torch ``tensor.data_ptr()``); ``d_out`` is accumulated into."""
        self._e._check(self._e._L.rpx_field_evaluate_device(self._e._ctx, self._h, int(d_points), int(npt),
                                                            float(time_ps), int(d_out)))

    @property
    def last_ms(self):
        return float(self._e._L.rpx_field_last_ms(self._h))

    def free(self):
        if self._h is not None:
            self._e._L.rpx_field_free(self._e._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Detector(object):
    """A fixed set of points with an accumulating complex E-field on the device (``rpx_detector``):
    gausslet collections are summed into it one after the other (EFieldSummation, fields.py:206-249)."""

    def __init__(self, engine, points, wavelengths, blending=1.0, time_ps=0.0):
        self._e = engine
        pts = np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)
        wl = np.ascontiguousarray(wavelengths, dtype=np.double).reshape(-1)
        h = C.c_void_p()
        engine._check(engine._L.rpx_detector_create(engine._ctx, pts.ctypes.data, pts.shape[0], wl.ctypes.data,
                                                    wl.shape[0], float(blending), float(time_ps), C.byref(h)))
        self._h = h
        self.npt = pts.shape[0]

    def reset(self):
        self._e._check(self._e._L.rpx_detector_reset(self._e._ctx, self._h))

    def accumulate(self, dev_rays):
        handle = dev_rays._h if isinstance(dev_rays, DeviceRays) else dev_rays
        self._e._check(self._e._L.rpx_detector_accumulate(self._e._ctx, self._h, handle))

    def read(self):
        """The accumulated field, (npt, 3) complex128."""
        out = np.zeros((self.npt, 3), dtype=np.complex128)
        self._e._check(self._e._L.rpx_detector_read(self._e._ctx, self._h, out.ctypes.data))
        return out

    @property
    def field_device_ptr(self):
        """Device address of the npt x 6 doubles (for an in-place NCCL all-reduce)."""
        return int(self._e._L.rpx_detector_field_device(self._h))

    @property
    def modes(self):
        return int(self._e._L.rpx_detector_modes(self._h))

    @property
    def ms(self):
        return float(self._e._L.rpx_detector_ms(self._e._ctx, self._h))

    def free(self):
        if self._h is not None:
            self._e._L.rpx_detector_free(self._e._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ConsumeResult(object):
    """What a streamed trace with device-side consumers leaves behind (``rpx_consume_result``)."""

    def __init__(self, engine, r, face_counts, is_gausslet, per_chunk_terminal, per_chunk_captured):
        self.counts = [int(r.counts[g]) for g in range(r.n_gens)]
        self.n_chunks = int(r.n_chunks)
        self.n_terminal, self.n_captured = int(r.n_terminal), int(r.n_captured)
        self.terminal = DeviceRays(engine, C.c_void_p(r.terminal), is_gausslet) if r.terminal else None
        self.captured = DeviceRays(engine, C.c_void_p(r.captured), is_gausslet) if r.captured else None
        self.device_ms, self.trace_ms = float(r.device_ms), float(r.trace_ms)
        self.kernel_ms = {"intersect": (float(r.intersect_ms), int(r.intersect_launches)),
                          "shade": (float(r.shade_ms), int(r.shade_launches))}
        self.launches = int(r.launches)
        self.face_counts = face_counts
        self.per_chunk_terminal = per_chunk_terminal
        self.per_chunk_captured = per_chunk_captured

    @property
    def segments(self):
        return int(sum(self.counts))

    @staticmethod
    def reference_order(arr, per_chunk):
        """Kept records come in (chunk, generation, ray) order; the reference's consumers
        (select_ray_intersections, ctracer.pyx:1999-2010) produce (generation, ray) order."""
        n_chunks, n_g = per_chunk.shape
        starts = np.concatenate([[0], np.cumsum(per_chunk.reshape(-1))]).astype(np.int64)
        pieces = []
        for g in range(n_g):
            for c in range(n_chunks):
                k = c * n_g + g
                pieces.append(arr[starts[k]:starts[k + 1]])
        return np.concatenate(pieces) if pieces else arr[:0]

    def free(self):
        for d in (self.terminal, self.captured):
            if d is not None:
                d.free()
        self.terminal = self.captured = None


class Engine(object):
    def __init__(self, device=0):
        self._L = load()
        ctx = C.c_void_p()
        rc = self._L.rpx_init(int(device), C.byref(ctx))
        if rc != A.RPX_OK:
            raise RpxError(rc, (self._L.rpx_last_error(None) or b"").decode())
        self._ctx = ctx
        self.device = int(device)
        self.scene = None
        self.n_traced_faces = 0

    def _check(self, rc):
        if rc != A.RPX_OK:
            raise RpxError(rc, (self._L.rpx_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx is not None:
            self._L.rpx_shutdown(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene --------------------------------------------------------------------
    def set_scene(self, scene):
        self._check(self._L.rpx_scene_set(self._ctx, scene.byref()))
        self.scene = scene  # keeps the host tables alive
        self.n_traced_faces = scene.n_traced_faces

    def set_capture_scene(self, scene, face_ids=None):
        """The capture FaceList (flattened like a traced scene) of select_ray_intersections."""
        ids = None
        if face_ids is not None:
            ids = np.ascontiguousarray(face_ids, dtype=np.uint32)
            assert ids.shape[0] == scene.c_scene.n_faces
        self._check(self._L.rpx_capture_scene_set(self._ctx, scene.byref(),
                                                  None if ids is None else ids.ctypes.data))
        self.capture_scene = scene

    def capture_collections(self, handles, wavelength_lists, is_gausslet, out=None):
        """``rpx_capture`` over device-resident collections + the reference's wavelength merge
        (np.unique over the concatenated tables, ctracer.pyx:2011-2016)."""
        n = len(handles)
        arr = (C.c_void_p * n)(*handles)
        sizes = [len(w) for w in wavelength_lists]
        offsets = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)[:-1]]), dtype=np.uint32)
        reduced, inverse = np.unique(np.concatenate([np.asarray(w, dtype=np.double) for w in wavelength_lists]),
                                     return_inverse=True)
        wl_map = np.ascontiguousarray(inverse, dtype=np.uint32)
        counts = np.zeros(n, dtype=np.uint64)
        h = C.c_void_p()
        self._check(self._L.rpx_capture(self._ctx, arr, n, offsets.ctypes.data, wl_map.ctypes.data,
                                        wl_map.shape[0], C.byref(h), counts.ctypes.data))
        dev = DeviceRays(self, h, is_gausslet)
        try:
            out = self.download(dev, out=out)  # ``out``: a (pinned) staging array to fill, or None
        finally:
            dev.free()
        return out, reduced, [int(c) for c in counts]

    # -- terminal rays / device-resident AoS -----------------------------------------
    def select_terminal(self, handles, is_gausslet, unterminated=True, faces=None):
        """``rpx_select_terminal`` over device-resident collections -> (DeviceRays, per-collection counts).
        ``faces``: indices of the faces whose hits count as terminal (absorbers, detectors)."""
        n = len(handles)
        arr = (C.c_void_p * n)(*handles)
        sel = None
        if faces is not None:
            sel = np.zeros(max(self.n_traced_faces, 1), dtype=np.uint8)
            sel[np.asarray(list(faces), dtype=np.int64)] = 1
        counts = np.zeros(n, dtype=np.uint64)
        h = C.c_void_p()
        self._check(self._L.rpx_select_terminal(self._ctx, arr, n, 1 if unterminated else 0,
                                                None if sel is None else sel.ctypes.data, C.byref(h), counts.ctypes.data))
        return DeviceRays(self, h, is_gausslet), [int(c) for c in counts]

    def export_device(self, dev_rays, d_ptr, capacity):
        """Packed AoS records of a device collection into caller-owned DEVICE memory (e.g. a torch uint8
        tensor's ``data_ptr()``): the send buffer of an NCCL gather."""
        self._check(self._L.rpx_rays_export_device(self._ctx, dev_rays._h, int(d_ptr), int(capacity)))

    def import_device(self, d_ptr, n, is_gausslet):
        h = C.c_void_p()
        self._check(self._L.rpx_rays_import_device(self._ctx, int(d_ptr), int(n), 1 if is_gausslet else 0, C.byref(h)))
        return DeviceRays(self, h, is_gausslet)

    def detector(self, points, wavelengths, blending=1.0, time_ps=0.0):
        return Detector(self, points, wavelengths, blending, time_ps)

    def trace_consume(self, rays, max_length, recursion_limit, n=None, is_gausslet=None, chunk_rays=0, terminal=False,
                      terminal_faces=None, terminal_capacity=0, capture=False, captured_capacity=0, detector=None,
                      per_chunk=False):
        """``rpx_trace_consume``: chunked trace whose generations stay on the device and are handed to the
        consumers (terminal-ray selection, capture plane, detector field).  ``rays``: a host array, or an
        integer DEVICE address of packed AoS records (then ``n`` and ``is_gausslet`` are required)."""
        flags = 0
        if isinstance(rays, np.ndarray):
            rays = np.ascontiguousarray(rays)
            is_g = self._is_gausslet(rays)
            ptr, n = rays.ctypes.data, rays.shape[0]
        else:
            ptr, is_g = int(rays), 1 if is_gausslet else 0
            flags |= A.CONSUME_SOURCE_ON_DEVICE
        if terminal:
            flags |= A.CONSUME_TERMINAL
        if capture:
            flags |= A.CONSUME_CAPTURE
        if detector is not None:
            flags |= A.CONSUME_FIELD | A.CONSUME_CAPTURE
        o = A.rpx_consume_opts()
        o.chunk_rays = int(chunk_rays)
        o.flags = flags
        sel = None
        if terminal_faces is not None:
            sel = np.zeros(max(self.n_traced_faces, 1), dtype=np.uint8)
            sel[np.asarray(list(terminal_faces), dtype=np.int64)] = 1
            o.terminal_faces = sel.ctypes.data
        o.terminal_capacity = int(terminal_capacity)
        o.captured_capacity = int(captured_capacity)
        o.detector = detector._h if detector is not None else None
        pct = pcc = None
        if per_chunk:
            default_chunk = (1 << 20) if is_g else (1 << 22)
            n_chunks = max(1, -(-int(n) // int(chunk_rays or default_chunk)))
            o.max_gens = A.CONSUME_MAX_GENS
            pct = np.zeros((n_chunks, A.CONSUME_MAX_GENS), dtype=np.uint64)
            pcc = np.zeros((n_chunks, A.CONSUME_MAX_GENS), dtype=np.uint64)
            o.per_chunk_terminal = pct.ctypes.data
            o.per_chunk_captured = pcc.ctypes.data
        r = A.rpx_consume_result()
        fc = np.zeros(max(self.n_traced_faces, 1), dtype=np.uint32)
        self._check(self._L.rpx_trace_consume(self._ctx, ptr, int(n), is_g, float(max_length), int(recursion_limit),
                                              C.byref(o), fc.ctypes.data, C.byref(r)))
        ng = int(r.n_gens)
        return ConsumeResult(self, r, fc[:self.n_traced_faces].copy(), is_g,
                             None if pct is None else pct[:, :ng].copy(), None if pcc is None else pcc[:, :ng].copy())

    # -- E-field summation --------------------------------------------------------------------
    def field_prepare(self, rays, wavelengths, modes=None, blending=1.0):
        """``rays``: a DeviceRays, a borrowed ``rpx_rays`` handle, or a host array (uploaded).
        Gausslets with ``modes=None`` get their modes fitted on the device."""
        own = None
        if isinstance(rays, np.ndarray):
            own = self.upload(rays)
            handle = own._h
        elif isinstance(rays, DeviceRays):
            handle = rays._h
        else:
            handle = rays
        wl = np.ascontiguousarray(wavelengths, dtype=np.double).reshape(-1)
        m = None
        if modes is not None:
            m = np.ascontiguousarray(modes, dtype=np.complex128).reshape(-1, 3)
        h = C.c_void_p()
        try:
            self._check(self._L.rpx_field_prepare(self._ctx, handle, None if m is None else m.ctypes.data,
                                                  wl.ctypes.data, wl.shape[0], float(blending), C.byref(h)))
        finally:
            if own is not None:
                own.free()
        return FieldModes(self, h)

    def project_to_sphere(self, dev_rays, centre=(0.0, 0.0, 0.0), radius=10.0):
        """fields.py:50-77 in place on a DeviceRays of plain rays; returns the boolean ``selector``."""
        c = np.ascontiguousarray(centre, dtype=np.double).reshape(3)
        sel = np.zeros(len(dev_rays), dtype=np.uint8)
        n_sel = C.c_uint64(0)
        self._check(self._L.rpx_rays_project_to_sphere(self._ctx, dev_rays._h, c.ctypes.data, float(radius),
                                                       sel.ctypes.data, C.byref(n_sel)))
        assert int(n_sel.value) == int(sel.sum())
        return sel.astype(bool)

    def field_prepare_neighbours(self, rays, neighbours_idx, wavelengths, blending=1.0, want_xy=False):
        """evaluate_neighbours -> evaluate_modes -> mode records (fields.py:80-111, cfields.pyx:217-228) for
        plain rays with their ``RayCollection.neighbours`` array.  Returns FieldModes, or
        (FieldModes, (x, y, dx, dy)) with ``want_xy``."""
        own = None
        if isinstance(rays, np.ndarray):
            own = self.upload(rays)
            handle = own._h
            n = len(own)
        else:
            handle = rays._h
            n = len(rays)
        nb = np.ascontiguousarray(neighbours_idx, dtype=np.int32)
        if nb.ndim != 2 or nb.shape[0] != n:  # numpy: boolean index did not match indexed array (fields.py:99)
            if own is not None:
                own.free()
            raise IndexError("neighbours_idx has shape %s for %d rays" % (nb.shape, n))
        wl = np.ascontiguousarray(wavelengths, dtype=np.double).reshape(-1)
        kept = int((nb >= 0).all(axis=1).sum())
        xy = np.zeros((4, kept, nb.shape[1])) if want_xy else None
        h = C.c_void_p()
        try:
            self._check(self._L.rpx_field_prepare_neighbours(self._ctx, handle, nb.ctypes.data, nb.shape[1], wl.ctypes.data,
                                                             wl.shape[0], float(blending),
                                                             None if xy is None else xy.ctypes.data, C.byref(h)))
        finally:
            if own is not None:
                own.free()
        fm = FieldModes(self, h)
        return (fm, tuple(xy)) if want_xy else fm

    def evaluate_modes(self, x, y, dx, dy, blending=1.0):
        """cfields.evaluate_modes (cfields.pyx:217-228) on explicit n x 6 arrays -> n x 3 complex128."""
        x, y, dx, dy = (np.ascontiguousarray(a, dtype=np.double) for a in (x, y, dx, dy))
        n, row = x.shape
        out = np.zeros((n, 3), dtype=np.complex128)
        self._check(self._L.rpx_unit_evaluate_modes(self._ctx, x.ctypes.data, y.ctypes.data, dx.ctypes.data, dy.ctypes.data,
                                                    n, row, float(blending), out.ctypes.data))
        return out

    # -- rays -----------------------------------------------------------------------
    @staticmethod
    def _is_gausslet(arr):
        if arr.dtype == A.gausslet_dtype:
            return 1
        if arr.dtype == A.ray_dtype:
            return 0
        raise TypeError("rays must be a ray_dtype or gausslet_dtype array, got %s" % (arr.dtype,))

    def upload(self, rays):
        rays = np.ascontiguousarray(rays)
        is_g = self._is_gausslet(rays)
        h = C.c_void_p()
        self._check(self._L.rpx_rays_upload(self._ctx, rays.ctypes.data, rays.shape[0], is_g, C.byref(h)))
        return DeviceRays(self, h, is_g)

    def clone(self, dev_rays):
        h = C.c_void_p()
        self._check(self._L.rpx_rays_clone(self._ctx, dev_rays._h, C.byref(h)))
        return DeviceRays(self, h, dev_rays.is_gausslet)

    def result_empty(self, n, dtype):
        """Array for a result the caller keeps (a generation, a captured collection): a pooled page-locked
        block when the result is large -- no first-touch page faults, D2H at PCIe speed -- else numpy.empty."""
        return get_pool(self._L).empty(n, dtype)

    def pinned_empty(self, n, dtype):
        """numpy array over page-locked host memory (rpx_host_alloc)."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 1)
        ptr = self._L.rpx_host_alloc(nbytes)
        if not ptr:
            raise MemoryError("rpx_host_alloc(%d) failed" % nbytes)
        buf = (C.c_char * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        return arr

    def download(self, dev_rays, out=None):
        n = len(dev_rays)
        dtype = A.gausslet_dtype if dev_rays.is_gausslet else A.ray_dtype
        if out is None:
            out = self.result_empty(n, dtype)
        assert out.dtype == dtype and out.shape[0] >= n and out.flags.c_contiguous
        self._check(self._L.rpx_rays_download(self._ctx, dev_rays._h, out.ctypes.data, out.shape[0]))
        return out[:n]

    # -- tracing ----------------------------------------------------------------------
    def trace(self, rays, max_length, recursion_limit, flags=A.TRACE_DEFAULT):
        """Host buffers in (H2D inside): ``rpx_trace``."""
        rays = np.ascontiguousarray(rays)
        is_g = self._is_gausslet(rays)
        h = C.c_void_p()
        self._check(self._L.rpx_trace(self._ctx, rays.ctypes.data, rays.shape[0], is_g, float(max_length),
                                      int(recursion_limit), int(flags), C.byref(h)))
        return TraceResult(self, h, is_g)

    def trace_streamed(self, rays, max_length, recursion_limit, out, chunk_rays=0):
        """``rpx_trace_streamed``: chunked trace with upload / trace / download overlapped.
        ``out`` is a list of arrays (one per expected generation, ideally pinned) that receive the
        generations.  Returns ``(generation views, face_counts, device_ms)``."""
        rays = np.ascontiguousarray(rays)
        is_g = self._is_gausslet(rays)
        n_out = len(out)
        for o in out:
            assert o.dtype == rays.dtype and o.flags.c_contiguous
        ptrs = (C.c_void_p * n_out)(*[o.ctypes.data for o in out])
        caps = np.ascontiguousarray([o.shape[0] for o in out], dtype=np.uint64)
        counts = np.zeros(n_out, dtype=np.uint64)
        n_gens = C.c_int(0)
        fc = np.zeros(max(self.n_traced_faces, 1), dtype=np.uint32)
        ms = C.c_double(0.0)
        self._check(self._L.rpx_trace_streamed(self._ctx, rays.ctypes.data, rays.shape[0], is_g, float(max_length),
                                               int(recursion_limit), int(chunk_rays), ptrs, caps.ctypes.data, n_out,
                                               counts.ctypes.data, C.byref(n_gens), fc.ctypes.data, C.byref(ms)))
        gens = [out[g][:int(counts[g])] for g in range(n_gens.value)]
        return gens, fc[:self.n_traced_faces].copy(), ms.value

    def trace_sequence(self, rays, face_seq, max_length, recursion_limit):
        """Sequential trace (``rpx_trace_sequence``): step s intersects only the face with
        global index ``face_seq[s]``."""
        rays = np.ascontiguousarray(rays)
        is_g = self._is_gausslet(rays)
        seq = np.ascontiguousarray(face_seq, dtype=np.int32)
        h = C.c_void_p()
        self._check(self._L.rpx_trace_sequence(self._ctx, rays.ctypes.data, rays.shape[0], is_g,
                                               float(max_length), int(recursion_limit), seq.ctypes.data,
                                               int(seq.shape[0]), C.byref(h)))
        return TraceResult(self, h, is_g)

    def trace_step(self, dev_rays, max_length, face_counts=None):
        """``rpx_trace_step``: one generation.  ``dev_rays`` is written back in place and stays the
        caller's; returns the (un-intersected) children as a new DeviceRays.  ``face_counts`` (uint32
        array of n_traced_faces) is added to."""
        h = C.c_void_p()
        fc = None
        if face_counts is not None:
            assert face_counts.dtype == np.uint32 and face_counts.shape[0] >= self.n_traced_faces
            fc = face_counts.ctypes.data
        self._check(self._L.rpx_trace_step(self._ctx, dev_rays._h, float(max_length), C.byref(h), fc))
        return DeviceRays(self, h, dev_rays.is_gausslet)

    def trace_device(self, dev_rays, max_length, recursion_limit, flags=A.TRACE_DEFAULT):
        """Inputs already resident (``rpx_trace_device``); consumes ``dev_rays``."""
        h = C.c_void_p()
        handle, dev_rays._h = dev_rays._h, None  # ownership moves to the library
        self._check(self._L.rpx_trace_device(self._ctx, handle, float(max_length), int(recursion_limit),
                                             int(flags), C.byref(h)))
        return TraceResult(self, h, dev_rays.is_gausslet)

    # -- unit entry points (function-level parity tests) -----------------------------------
    def unit_face_intersect(self, face_idx, p1, p2, is_base_ray=1):
        p1 = np.ascontiguousarray(p1, dtype=np.double).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=np.double).reshape(-1, 3)
        out = np.empty(p1.shape[0])
        self._check(self._L.rpx_unit_face_intersect(self._ctx, face_idx, p1.ctypes.data, p2.ctypes.data,
                                                    p1.shape[0], int(is_base_ray), out.ctypes.data))
        return out

    def unit_face_normal(self, face_idx, points):
        """-> (normal, tangent) of FaceList.compute_orientation_c at global points."""
        p = np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)
        n, t = np.empty_like(p), np.empty_like(p)
        self._check(self._L.rpx_unit_face_normal(self._ctx, face_idx, p.ctypes.data, p.shape[0],
                                                 n.ctypes.data, t.ctypes.data))
        return n, t

    def unit_material_eval(self, mat_idx, rays, point, normal, tangent):
        rays = np.ascontiguousarray(rays, dtype=A.ray_dtype).reshape(-1)
        n = rays.shape[0]
        pt = np.ascontiguousarray(np.broadcast_to(np.asarray(point, dtype=np.double), (n, 3)))
        nm = np.ascontiguousarray(np.broadcast_to(np.asarray(normal, dtype=np.double), (n, 3)))
        tg = np.ascontiguousarray(np.broadcast_to(np.asarray(tangent, dtype=np.double), (n, 3)))
        out = np.zeros(2 * n, dtype=A.ray_dtype)
        cnt = np.zeros(n, dtype=np.uint32)
        self._check(self._L.rpx_unit_material_eval(self._ctx, mat_idx, rays.ctypes.data, n, pt.ctypes.data,
                                                   nm.ctypes.data, tg.ctypes.data, out.ctypes.data,
                                                   cnt.ctypes.data))
        return out.reshape(n, 2), cnt

    def unit_distortion(self, dist_idx, x, y):
        """-> (z_offset_c(x, y), z_offset_and_gradient_c(x, y) as (n, 3))."""
        x = np.ascontiguousarray(x, dtype=np.double).reshape(-1)
        y = np.ascontiguousarray(y, dtype=np.double).reshape(-1)
        z = np.empty(x.shape[0])
        g = np.empty((x.shape[0], 3))
        self._check(self._L.rpx_unit_distortion(self._ctx, dist_idx, x.ctypes.data, y.ctypes.data,
                                                x.shape[0], z.ctypes.data, g.ctypes.data))
        return z, g


_ENGINES = {}


def get_engine(device=0):
    """Process-wide engine per device (one context per process per GPU)."""
    e = _ENGINES.get(device)
    if e is None or e._ctx is None:
        e = _ENGINES[device] = Engine(device)
    return e
