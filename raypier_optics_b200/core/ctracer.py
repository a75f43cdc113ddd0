"""Host-side mirror of the reference's ``raypier.core.ctracer`` object model.

Same names, constructor keywords and defaults as raypier/core/ctracer.pyx, but the
objects are plain *parameter holders*: all trace arithmetic (intersection,
orientation, material evaluation) lives in librpx on the GPU.  They exist so a
model can be assembled -- and the parity tests can read like the reference's own
tests -- on a box where the reference itself is not installed.  Genuine
``raypier.core`` objects are accepted by ``trace_rays`` as well (see scene.py).
"""
import math

import numpy as np

from .._abi import (GAUSSLET, NPARA, PARABASAL, REFL_RAY, gausslet_dtype, para_dtype,  # noqa: F401
                    ray_dtype)

GAUSSLET_ = GAUSSLET
PARABASAL_ = PARABASAL
INF = float("inf")


def get_ray_size():
    """sizeof(ray_t) (ctracer.pyx:2399-2400)."""
    return ray_dtype.itemsize


class Transform(object):
    """ctracer.pyx:268-295"""

    def __init__(self, rotation=[[1, 0, 0], [0, 1, 0], [0, 0, 1]], translation=[0, 0, 0]):
        self.rotation = rotation
        self.translation = translation

    @property
    def rotation(self):
        return [list(r) for r in self._rot]

    @rotation.setter
    def rotation(self, rot):
        self._rot = [[float(v) for v in row] for row in rot]

    @property
    def translation(self):
        return tuple(self._dt)

    @translation.setter
    def translation(self, dt):
        self._dt = [float(v) for v in dt]


class Ray(object):
    """A single ray_t record (ctracer.pyx:396-714): attribute access to one element
    of a ``ray_dtype`` array."""
    _fields = ray_dtype.names

    def __init__(self, **kwds):
        object.__setattr__(self, "_rec", np.zeros(1, dtype=ray_dtype))
        self._rec['length'] = INF
        self._rec['refractive_index'] = 1.0
        object.__setattr__(self, "max_length", 1000.0)
        for k, v in kwds.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        if name in ray_dtype.names:
            v = self._rec[name][0]
            if isinstance(v, np.ndarray):
                return tuple(float(x) for x in v)
            return v.item()
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in ray_dtype.names:
            self._rec[name][0] = value
        else:
            object.__setattr__(self, name, value)

    @property
    def record(self):
        return self._rec[0]

    @property
    def power(self):
        """ray_power_ (ctracer.pyx:1971-1978)"""
        r = self._rec[0]
        n = r['refractive_index'].real
        return (abs(r['E1_amp']) ** 2) * n + (abs(r['E2_amp']) ** 2) * n

    @property
    def termination(self):
        r = self._rec[0]
        length = min(float(r['length']), self.max_length)
        return tuple(r['origin'] + r['direction'] * np.float32(length))


class _Collection(object):
    _dtype = None

    def __init__(self, max_size=0):
        self._data = np.zeros(0, dtype=self._dtype)
        self._parent = None
        self._wavelengths = None
        self._neighbours = None

    # -- container protocol -----------------------------------------------------
    def __len__(self):
        return int(self._data.shape[0])

    @property
    def n_rays(self):
        return int(self._data.shape[0])

    def copy_as_array(self):
        """Always a copy (ctracer.pyx:1048-1054, 1286-1292)."""
        return self._data.copy()

    @classmethod
    def from_array(cls, data):
        """ctracer.pyx:1142-1154 / 1304-1319 -- the data is copied."""
        data = np.asarray(data)
        if data.dtype != cls._dtype:
            raise ValueError("Array must have %s dtype" % cls.__name__)
        rc = cls(data.shape[0])
        rc._data = np.ascontiguousarray(data).copy()
        return rc

    def clear_ray_list(self):
        self._data = self._data[:0].copy()

    @property
    def wavelengths(self):
        return self._wavelengths

    @wavelengths.setter
    def wavelengths(self, wl_list):
        self._wavelengths = np.ascontiguousarray(wl_list, dtype=np.double)

    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, rc):
        self._parent = rc
        self._neighbours = None
        self._wavelengths = rc._wavelengths

    # -- used by the tracer shim: replace contents in place (the reference mutates the
    #    parent collection's length / end_face_idx, ctracer.pyx:2086-2087,1900-1903) ----
    def _assign_array(self, data):
        assert data.dtype == self._dtype
        self._data = data


def _ray_field(name):
    def get(self):
        return self._base()[name].copy()
    return property(get)


class RayCollection(_Collection):
    """ctracer.pyx:973-1154"""
    _dtype = ray_dtype

    def _base(self):
        return self._data

    def reset_length(self, max_length=INF):
        self._data['length'] = max_length

    @property
    def base_rays(self):
        return self

    def __getitem__(self, idx):
        if idx >= self.n_rays:
            raise IndexError("Requested index %d from a size %d array" % (idx, self.n_rays))
        r = Ray()
        r._rec[0] = self._data[idx]
        return r

    def __iter__(self):
        for i in range(self.n_rays):
            yield self[i]

    def add_ray(self, r):
        self._data = np.concatenate([self._data, r._rec])

    def add_ray_list(self, rays):
        for r in rays:
            if not isinstance(r, Ray):
                raise TypeError("ray list contains non-Ray instance")
        if rays:
            self._data = np.concatenate([self._data] + [r._rec for r in rays])

    @property
    def termination(self):
        d = self._data
        return d['origin'] + d['direction'] * d['length'][:, None]

    @property
    def neighbours(self):
        """Lazy neighbour map derived from parent_idx + REFL bit (ctracer.pyx:1084-1131)."""
        if self._parent is None:
            return self._neighbours
        if self._neighbours is None:
            pnb = self._parent.neighbours
            if pnb is None:
                return None
            d = self._data
            rtype = (d['ray_type_id'] & REFL_RAY).astype(np.int64)
            pidx = d['parent_idx'].astype(np.int64)
            rmap = np.full((self._parent.n_rays, 2), -1, dtype=np.int32)
            rmap[pidx, rtype] = np.arange(d.shape[0], dtype=np.int32)
            cnb = pnb[pidx]  # (n, k)
            nb = np.where(cnb >= 0, rmap[np.clip(cnb, 0, None), rtype[:, None]], -1).astype(np.int32)
            self._neighbours = nb
        return self._neighbours

    @neighbours.setter
    def neighbours(self, nb):
        self._neighbours = None if nb is None else np.asarray(nb, dtype=np.int32)


for _name in ray_dtype.names:
    setattr(RayCollection, _name, _ray_field(_name))


class GaussletBaseRayView(object):
    """ctracer.pyx:1157-1183: the base rays of a GaussletCollection seen through the RayArrayView API
    (ctracer.pyx:717-969): ``len``, ``copy_as_array`` and one array property per ray_t member --
    e.g. ``gausslets.base_rays.end_face_idx``, which BaseRaySource.set_face_sequence (sources.py:138-155)
    reads for every traced generation."""

    def __init__(self, owner):
        self.owner = owner

    def __len__(self):
        return self.owner.n_rays

    def _base(self):
        return self.owner._data['base_ray']

    def copy_as_array(self):
        return np.ascontiguousarray(self.owner._data['base_ray']).copy()

    @property
    def termination(self):
        d = self._base()
        return d['origin'] + d['direction'] * d['length'][:, None]


class GaussletCollection(_Collection):
    """ctracer.pyx:1185-1570"""
    _dtype = gausslet_dtype

    def _base(self):
        return self._data['base_ray']

    def reset_length(self, max_length=INF):
        self._data['base_ray']['length'] = max_length
        self._data['para_rays']['length'] = max_length

    @property
    def base_rays(self):
        return GaussletBaseRayView(self)

    @classmethod
    def from_rays(cls, data):
        """ctracer.pyx:1321-1345"""
        data = np.asarray(data)
        if data.dtype != ray_dtype:
            raise ValueError("Array must have ray_dtype dtype")
        n = data.shape[0]
        g = np.zeros(n, dtype=gausslet_dtype)
        g['base_ray'] = data
        for name in ('origin', 'direction', 'normal'):
            g['para_rays'][name] = data[name][:, None, :]
        g['para_rays']['length'] = data['length'][:, None]
        rc = cls(n)
        rc._data = g
        return rc

    def config_parabasal_rays(self, wavelength_list, radius, working_dist):
        """ctracer.pyx:1430-1481, vectorised with the same per-element operation order."""
        g = self._data
        n = g.shape[0]
        if n == 0:
            return
        wl = np.asarray(wavelength_list, dtype=np.double)
        b = g['base_ray']
        d = b['direction']
        mag = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        base_d = d / mag[:, None]
        o = np.where((base_d[:, 0] > base_d[:, 1])[:, None], np.array([0.0, 1.0, 0.0]),
                     np.array([1.0, 0.0, 0.0]))

        def cross(a, c):
            return np.stack([a[:, 1] * c[:, 2] - a[:, 2] * c[:, 1],
                             a[:, 2] * c[:, 0] - a[:, 0] * c[:, 2],
                             a[:, 0] * c[:, 1] - a[:, 1] * c[:, 0]], axis=1)

        def norm(a):
            m = np.sqrt(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])
            return a / m[:, None]

        d1 = norm(cross(base_d, o))
        d2 = norm(cross(base_d, d1))
        theta0 = wl[b['wavelength_idx']] / (math.pi * radius * 1000.0)
        for j in range(0, 6, 2):
            angle = (j * 2 * math.pi / 6)
            ca, sa = math.cos(angle), math.sin(angle)
            oo = d1 * (radius * ca) + d2 * (radius * sa)
            oo = oo + b['origin']
            oo = oo + base_d * working_dist
            dd = d1 * (-theta0 * sa)[:, None] + d2 * (theta0 * ca)[:, None]
            da = base_d + dd
            db = base_d - dd
            g['para_rays']['direction'][:, j] = norm(da)
            g['para_rays']['origin'][:, j] = oo - da * working_dist
            g['para_rays']['direction'][:, j + 1] = norm(db)
            g['para_rays']['origin'][:, j + 1] = oo - db * working_dist
            for jj in (j, j + 1):
                g['para_rays']['normal'][:, jj] = b['normal']
                g['para_rays']['length'][:, jj] = b['length']

    def project_to_plane(self, origin, direction):
        """ctracer.pyx:1391-1420: move every base ray and its six parabasal rays along their own
        directions onto the plane through ``origin`` with normal ``direction``; the base ray's
        accumulated_path grows by Re(n) * distance.  Vectorised with the per-element operation order of
        the reference's loop (dotprod_ = x*x' + y*y' + z*z', left to right)."""
        g = self._data
        if g.shape[0] == 0:
            return
        o = np.array([float(v) for v in origin], dtype=np.double)
        d = np.array([float(v) for v in direction], dtype=np.double)
        d = d / math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])  # norm_

        def dot(a):  # dotprod_(a, d) over the last axis
            return a[..., 0] * d[0] + a[..., 1] * d[1] + a[..., 2] * d[2]

        b = g['base_ray']
        a = dot(o - b['origin']) / dot(b['direction'])
        b['origin'] = b['origin'] + b['direction'] * a[:, None]
        b['accumulated_path'] += b['refractive_index'].real * a
        p = g['para_rays']
        a = dot(o - p['origin']) / dot(p['direction'])
        p['origin'] = p['origin'] + p['direction'] * a[..., None]

    @property
    def lagrange_invariant(self):
        """ctracer.pyx:1347-1389: the (normalised) Lagrange invariant of every gausslet, 1 for a
        diffraction-limited fundamental mode."""
        g = self._data
        b, p = g['base_ray'], g['para_rays']
        e = b['E_vector']
        axis1 = e / np.sqrt(e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])[:, None]
        dd = b['direction']
        axis2 = np.stack([axis1[:, 1] * dd[:, 2] - axis1[:, 2] * dd[:, 1],
                          axis1[:, 2] * dd[:, 0] - axis1[:, 0] * dd[:, 2],
                          axis1[:, 0] * dd[:, 1] - axis1[:, 1] * dd[:, 0]], axis=1)

        def dot3(a, c):
            return a[..., 0] * c[..., 0] + a[..., 1] * c[..., 1] + a[..., 2] * c[..., 2]

        off = p['origin'] - b['origin'][:, None, :]
        hx, hy = dot3(off, axis1[:, None, :]), dot3(off, axis2[:, None, :])
        ux, uy = dot3(p['direction'], axis1[:, None, :]), dot3(p['direction'], axis2[:, None, :])

        def term(i, j):  # (h_i . u_j - h_j . u_i)^2 with the z components zero
            hu = hx[:, i] * ux[:, j] + hy[:, i] * uy[:, j] + 0.0 * 0.0
            uh = hx[:, j] * ux[:, i] + hy[:, j] * uy[:, i] + 0.0 * 0.0
            return (hu - uh) ** 2

        v = np.zeros(g.shape[0])
        for i, j in ((0, 5), (1, 2), (3, 4), (1, 4), (0, 3), (2, 5)):
            v = v + term(i, j)
        wavelen = np.asarray(self.wavelengths, dtype=np.double)
        return 1000 * np.sqrt(v / 6) / wavelen[b['wavelength_idx']]

    def extend(self, gc):
        """ctracer.pyx:1294-1303: append the gausslets of another collection."""
        self._data = np.concatenate([self._data, np.asarray(gc.copy_as_array()).view(gausslet_dtype)])

    def scale_amplitude(self, scale):
        self._data['base_ray']['E1_amp'] *= scale
        self._data['base_ray']['E2_amp'] *= scale

    @property
    def total_power(self):
        # ctracer.pyx:1487-1502: four terms per ray added in sequence (cumsum keeps the loop's order,
        # np.sum would add pairwise)
        b = self._data['base_ray']
        if b.shape[0] == 0:
            return 0.0
        n = b['refractive_index'].real
        e1, e2 = b['E1_amp'], b['E2_amp']
        terms = np.stack([(e1.real * e1.real) * n, (e1.imag * e1.imag) * n,
                          (e2.real * e2.real) * n, (e2.imag * e2.imag) * n], axis=1).reshape(-1)
        return float(np.cumsum(terms)[-1])

    @property
    def para_origin(self):
        return self._data['para_rays']['origin'].copy()

    @property
    def para_direction(self):
        return self._data['para_rays']['direction'].copy()

    @property
    def para_normal(self):
        return self._data['para_rays']['normal'].copy()


for _name in ray_dtype.names:
    setattr(GaussletCollection, _name, _ray_field(_name))
    setattr(GaussletBaseRayView, _name, _ray_field(_name))


class InterfaceMaterial(object):
    """ctracer.pyx:1573-1656"""

    def __init__(self, **kwds):
        self._wavelengths = np.array([], dtype=np.double)

    def is_decomp_material(self):
        return False

    @property
    def wavelengths(self):
        return self._wavelengths

    @wavelengths.setter
    def wavelengths(self, wavelengths):
        self._wavelengths = np.asarray(wavelengths, dtype=np.double)
        self.on_set_wavelengths()

    def on_set_wavelengths(self):
        pass


class Shape(object):
    """ctracer.pyx:1659-1664 -- the base shape contains every point."""


class ImplicitSurface(object):
    """ctracer.pyx:1667-1679"""


class Distortion(object):
    """ctracer.pyx:1682-1729"""


class Face(object):
    """ctracer.pyx:1732-1809"""
    params = []

    def __init__(self, owner=None, tolerance=0.0001, max_length=100, material=None, **kwds):
        self.name = "base Face class"
        self.tolerance = tolerance
        self.owner = owner
        self.max_length = max_length
        if isinstance(material, InterfaceMaterial) or (
                material is not None and hasattr(material, "is_decomp_material")):
            self.material = material
        else:
            from .cmaterials import PECMaterial
            self.material = PECMaterial()
        self.invert_normal = int(kwds.get('invert_normal', 0))
        self.idx = 0
        self.count = 0

    def update(self):
        """Copy ``params`` attributes owner -> face (ctracer.pyx:1761-1767)."""
        for name in self.params:
            v = getattr(self.owner, name)
            setattr(self, name, v)


class FaceList(object):
    """A group of faces which share a transform (ctracer.pyx:1813-1964)."""

    def __init__(self, owner=None):
        self.transform = Transform()
        self.inverse_transform = Transform()
        self.owner = owner
        self.faces = []

    def sync_transforms(self):
        """ctracer.pyx:1820-1837: pull the 3x4 matrix and its inverse out of the owner's
        VTK-style transform (``owner.transform.matrix.get_element(i, j)``)."""
        try:
            trans = self.owner.transform
        except AttributeError:
            print("NO OWNER", self.owner)
            return
        m = trans.matrix
        rot = [[m.get_element(i, j) for j in range(3)] for i in range(3)]
        dt = [m.get_element(i, 3) for i in range(3)]
        self.transform = Transform(rotation=rot, translation=dt)
        inv_trans = trans.linear_inverse
        m = inv_trans.matrix
        rot = [[m.get_element(i, j) for j in range(3)] for i in range(3)]
        dt = [m.get_element(i, 3) for i in range(3)]
        self.inverse_transform = Transform(rotation=rot, translation=dt)

    def __getitem__(self, intidx):
        return self.faces[intidx]


# ---- capture planes -------------------------------------------------------------------------
def _select_intersections(face_set, ray_col_list, cls, device=0):
    from ..engine import get_engine
    from ..scene import Scene
    ray_col_list = list(ray_col_list)
    eng = get_engine(device)
    wl_lists = [np.asarray(rc.wavelengths, dtype=np.double) for rc in ray_col_list]
    # the geometry of the capture faces does not depend on the wavelengths; one entry keeps
    # the flattened tables well-formed
    cap = Scene([face_set], np.asarray([1.0]))
    eng.set_capture_scene(cap, [int(getattr(f, "idx", 0)) for f in cap.all_faces])
    devs = []
    try:
        for rc in ray_col_list:
            a = np.ascontiguousarray(rc.copy_as_array())
            a = a.view(gausslet_dtype if a.dtype.itemsize == gausslet_dtype.itemsize else ray_dtype)
            devs.append(eng.upload(a))
        arr, reduced, _ = eng.capture_collections([d._h for d in devs], wl_lists, cls is GaussletCollection)
    finally:
        for d in devs:
            d.free()
    out = cls.from_array(arr)
    out.wavelengths = reduced
    return out


def select_ray_intersections(face_set, ray_col_list, device=0):
    """ctracer.pyx:1981-2017 on the GPU: the rays of every collection that cross the capture
    ``face_set`` between their origin and their end point, in (collection, ray) order, cut at the
    capture face, with the wavelength tables merged.  Collections are uploaded; use
    ``TraceResult.capture`` to filter generations that are still resident on the device."""
    return _select_intersections(face_set, ray_col_list, RayCollection, device)


def select_gausslet_intersections(face_set, ray_col_list, device=0):
    """ctracer.pyx:2020-2058 on the GPU (see select_ray_intersections)."""
    return _select_intersections(face_set, ray_col_list, GaussletCollection, device)
