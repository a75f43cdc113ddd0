"""Scene flattening: FaceList / Face / InterfaceMaterial objects -> the flat POD
tables of include/rpx.h.

This is the host half of the drop-in boundary.  It walks ``face_lists`` exactly as
``raypier.core.tracer.trace_rays`` does (raypier/core/tracer.py:28-37 of the
reference): ``all_faces = chain(fs.faces for fs in face_lists)``, ``f.idx = i``.
Objects are read by *duck typing on the class name*, so both this package's own
host mirrors (``raypier_optics_b200.core``) and genuine ``raypier.core`` objects
can be flattened (SURVEY.md section 8b lists which attributes are readable).

No arithmetic of the trace happens here; the only numbers computed on the host
are the ones the reference also computes on the host at set-up time
(ExtrudedPlanarFace.calc_normal cfaces.pyx:657-663, OrientedPolygonFace y-axis
cfaces.pyx:1156, dispersion tables cmaterials.pyx:206-228) and the Zernike
recursion *schedule* (see ``build_zernike_tapes``).
"""
import ctypes as C
import math

import numpy as np

from . import _abi as A


class UnsupportedSceneError(NotImplementedError):
    """A face / material / shape class that librpx has no kernel for."""


def _cls_name(obj, known):
    for k in type(obj).__mro__:
        if k.__name__ in known:
            return k.__name__
    raise UnsupportedSceneError(
        "%s is not a supported type (known: %s)" % (type(obj).__name__, sorted(known)))


def _norm3(v):
    """norm_ of the reference (ctracer.pyx:251-256), same op order, IEEE doubles."""
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    mag = math.sqrt(x * x + y * y + z * z)
    return (x / mag, y / mag, z / mag)


def _cross3(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _transform_rows(t):
    """Transform-like object (``.rotation`` 3x3, ``.translation`` 3) -> (m[9], t[3])."""
    rot = np.asarray(t.rotation, dtype=np.double).reshape(9)
    dt = np.asarray(t.translation, dtype=np.double).reshape(3)
    return rot, dt


# ----------------------------------------------------------------------------
# Zernike tapes
# ----------------------------------------------------------------------------
def _cdiv(a, b):
    """C integer division (truncation toward zero), as under cdivision=True."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


class _ZTapeBuilder:
    """Symbolic execution of the reference's memoised Zernike recursion
    (cdistortions.pyx:149-178 R, 189-219 R', 288-316 R/r).

    Which memo slot is NaN when is independent of the ray, so running the
    recursion once with slots tracked as "filled / not filled" yields the exact
    sequence of floating-point operations the reference performs for every ray,
    including its slot-aliasing quirk (R' uses a stale ``half_n`` for kC).
    Operands: 0 -> 0.0, 1 -> 1.0, 2+3*k+w -> workspace[w][k].
    """

    def __init__(self, prefill_r00):
        self.ops = []
        self.filled = (set(), set(), set())
        if prefill_r00:  # z_offset_c: workspace[0,0] = 1.0 (cdistortions.pyx:433)
            self.ops.append((3, 0, 0, 0, 0, 0, 0))
            self.filled[0].add(0)

    @staticmethod
    def ref(w, k):
        return 2 + 3 * k + w

    def R(self, k, n, m):
        if n < m:
            return 0
        if n == 0:
            return 1
        if k in self.filled[0]:
            return self.ref(0, k)
        nA = n - 1
        mA = abs(m - 1)
        half_n = _cdiv(nA, 2)
        kA = half_n * (half_n + 1) + abs(mA)
        nB = nA
        mB = m + 1
        kB = half_n * (half_n + 1) + abs(mB)
        nC = n - 2
        mC = m
        half_n = _cdiv(nC, 2)
        kC = half_n * (half_n + 1) + abs(mC)
        a = self.R(kA, nA, mA)
        b = self.R(kB, nB, mB)
        c = self.R(kC, nC, mC)
        self.ops.append((0, k, a, b, c, 0, 0))
        self.filled[0].add(k)
        return self.ref(0, k)

    def Rp(self, k, n, m):
        if n < m:
            return 0
        if n == 0:
            return 0
        if k in self.filled[1]:
            return self.ref(1, k)
        nA = n - 1
        mA = abs(m - 1)
        half_n = _cdiv(nA, 2)
        kA = half_n * (half_n + 1) + abs(mA)
        nB = nA
        mB = m + 1
        kB = half_n * (half_n + 1) + abs(mB)
        nC = n - 2
        mC = m
        kC = (half_n - 1) * half_n + abs(mC)  # stale half_n: reference quirk
        a = self.R(kA, nA, mA)
        b = self.R(kB, nB, mB)
        d = self.Rp(kA, nA, mA)
        e = self.Rp(kB, nB, mB)
        c = self.Rp(kC, nC, mC)
        self.ops.append((1, k, a, b, c, d, e))
        self.filled[1].add(k)
        return self.ref(1, k)

    def Rr(self, k, n, m):
        if n < m:
            return 0
        if k in self.filled[2]:
            return self.ref(2, k)
        nA = n - 1
        mA = abs(m - 1)
        half_n = _cdiv(nA, 2)
        kA = half_n * (half_n + 1) + abs(mA)
        nB = nA
        mB = m + 1
        kB = half_n * (half_n + 1) + abs(mB)
        nC = n - 2
        mC = m
        half_n = _cdiv(nC, 2)
        kC = half_n * (half_n + 1) + abs(mC)
        a = self.R(kA, nA, mA)
        b = self.R(kB, nB, mB)
        c = self.Rr(kC, nC, mC)
        self.ops.append((2, k, a, b, c, 0, 0))
        self.filled[2].add(k)
        return self.ref(2, k)


def build_zernike_tapes(coefs):
    """coefs: list of (j, n, m, k, value) in the reference's (sorted-by-j) order.
    Returns (tape_z, tape_g, per-coef operands [(opR, opRp, opRr, opR_z)])."""
    tz = _ZTapeBuilder(prefill_r00=True)
    tg = _ZTapeBuilder(prefill_r00=False)
    operands = []
    for (j, n, m, k, value) in coefs:
        am = abs(m)
        opR_z = tz.R(k, n, am)
        opR = tg.R(k, n, am)
        opRp = tg.Rp(k, n, am)
        opRr = tg.Rr(k, n, am)
        operands.append((opR, opRp, opRr, opR_z))
    return tz.ops, tg.ops, operands


# ----------------------------------------------------------------------------
# Scene builder
# ----------------------------------------------------------------------------
_FACE_CLASSES = {
    "CircularFace", "ShapedPlanarFace", "ImplicitBoundedPlanarFace", "ElipticalPlaneFace",
    "RectangularFace", "SphericalFace", "ShapedSphericalFace", "ExtrudedPlanarFace", "PolygonFace",
    "OrientedPolygonFace", "OffAxisParabolicFace", "EllipsoidalFace", "SaddleFace",
    "CylindericalFace", "AxiconFace", "ConicRevolutionFace", "AsphericFace",
    "ExtendedPolynomialFace", "DistortionFace", "ExtrudedBezierFace", "UVPatchFace", "OBBTreeFace",
}
_MATERIAL_CLASSES = {
    "OpaqueMaterial", "TransparentMaterial", "PECMaterial", "PartiallyReflectiveMaterial",
    "LinearPolarisingMaterial", "WaveplateMaterial", "DielectricMaterial", "FullDielectricMaterial",
    "FullDielectricDispersiveMaterial", "SingleLayerCoatedMaterial", "CoatedDispersiveMaterial",
    "DiffractionGratingMaterial", "CircularApertureMaterial", "RectangularApertureMaterial",
    "ResampleGaussletMaterial",
}
_SHAPE_CLASSES = {"CircleShape", "RectangleShape", "PolygonShape", "InvertShape", "BooleanAND",
                  "BooleanOR", "BooleanXOR", "Shape"}
_IMPLICIT_CLASSES = {"NullSurface", "Plane", "Sphere", "Cylinder", "Invert", "Union", "Intersection",
                     "Difference"}
_DISTORTION_CLASSES = {"SimpleTestZernikeJ7", "ZernikeDistortion"}


class Scene:
    """The flattened scene: numpy tables + a ctypes ``rpx_scene`` pointing at them.

    ``all_faces`` is the reference's ``all_faces`` list (position == ``Face.idx``).
    """

    def __init__(self, face_lists, wavelengths):
        self.wavelengths = np.ascontiguousarray(wavelengths, dtype=np.double).reshape(-1)
        self.face_lists = list(face_lists)
        self.all_faces = [f for fs in self.face_lists for f in fs.faces]
        self._faces = []
        self._extra_faces = []  # DistortionFace bases, appended after the traced faces
        self._face_sets = []
        self._materials = []
        self._mat_index = {}
        self._shape_ops = []
        self._implicit_ops = []
        self._distortions = []
        self._dist_index = {}
        self._zcoefs = []
        self._ztape = []
        self._ntab = []
        self._pool = []      # numpy chunks (a mesh block is millions of doubles: no Python float lists)
        self._pool_len = 0
        self._build()
        self._finalise()

    # -- pool helpers ---------------------------------------------------------
    def _pool_add(self, values):
        off = self._pool_len
        chunk = np.array(values, dtype=np.double).reshape(-1)  # a private copy
        self._pool.append(chunk)
        self._pool_len += chunk.shape[0]
        return off

    def _mesh_block(self, points, cells, tolerance):
        """The pool block of a triangle mesh (include/rpx.h, RPX_FACE_MESH): header, the raw points and
        cells (all the reference object holds, obbtree.pyx:204-223), then triangle records in BVH leaf
        order and the BVH nodes for the device traversal.  Returns (offset, n_cells, n_nodes)."""
        from .core.obbtree import build_bvh, triangle_records
        points = np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)
        cells = np.ascontiguousarray(cells, dtype=np.int64).reshape(-1, 3)
        if len(cells) < 1 or len(points) < 3:
            raise ValueError("a mesh face needs at least one triangle")
        if cells.min() < 0 or cells.max() >= len(points):
            raise IndexError("mesh cell refers to a missing point")
        order, nodes = build_bvh(points, cells)
        p1, v1, v2, n = triangle_records(points, cells)
        tris = np.zeros((len(cells), 16))
        tris[:, 0:3], tris[:, 3:6], tris[:, 6:9], tris[:, 9:12] = p1[order], v1[order], v2[order], n[order]
        tris[:, 12] = order
        n_pts, n_cells, n_nodes = len(points), len(cells), len(nodes)
        off_points = 8
        off_cells = off_points + 3 * n_pts
        off_tris = off_cells + 3 * n_cells
        off_nodes = off_tris + 16 * n_cells
        header = [n_pts, n_cells, n_nodes, off_points, off_cells, off_tris, off_nodes, float(tolerance)]
        off = self._pool_add(header)
        self._pool_add(points)
        self._pool_add(cells.astype(np.double))
        self._pool_add(tris)
        self._pool_add(nodes)
        return off, n_cells, n_nodes

    # -- shapes ---------------------------------------------------------------
    def _emit_shape(self, shape):
        name = _cls_name(shape, _SHAPE_CLASSES)
        op = np.zeros((), dtype=A.shape_op_dtype)
        if name == "Shape":
            op['type'] = A.SHAPE_TRUE
        elif name == "CircleShape":
            cx, cy = shape.centre
            op['type'] = A.SHAPE_CIRCLE
            op['p'][:3] = (cx, cy, shape.radius)
        elif name == "RectangleShape":
            cx, cy = shape.centre
            op['type'] = A.SHAPE_RECT
            op['p'][:4] = (cx, cy, shape.width, shape.height)
        elif name == "PolygonShape":
            pts = np.asarray(shape.coordinates, dtype=np.double).reshape(-1, 2)
            op['type'] = A.SHAPE_POLYGON
            op['aux_off'] = self._pool_add(pts)
            op['aux_n'] = pts.shape[0]
        elif name == "InvertShape":
            self._emit_shape(shape.shape)
            op['type'] = A.SHAPE_NOT
        else:
            self._emit_shape(shape.shape1)
            self._emit_shape(shape.shape2)
            op['type'] = {"BooleanAND": A.SHAPE_AND, "BooleanOR": A.SHAPE_OR,
                          "BooleanXOR": A.SHAPE_XOR}[name]
        self._shape_ops.append(op)

    def _shape_program(self, shape):
        if shape is None:
            return -1, 0
        off = len(self._shape_ops)
        self._emit_shape(shape)
        return off, len(self._shape_ops) - off

    # -- implicit surfaces -------------------------------------------------------
    def _emit_implicit(self, surf):
        name = _cls_name(surf, _IMPLICIT_CLASSES)
        op = np.zeros((), dtype=A.implicit_op_dtype)
        if name == "NullSurface":
            op['type'] = A.IMPL_NULL
        elif name == "Plane":
            op['type'] = A.IMPL_PLANE
            op['p'][0:3] = surf.origin
            op['p'][3:6] = surf.normal  # already normalised by the setter
        elif name == "Sphere":
            op['type'] = A.IMPL_SPHERE
            op['p'][0:3] = surf.centre
            op['p'][3] = surf.radius
        elif name == "Cylinder":
            op['type'] = A.IMPL_CYLINDER
            op['p'][0:3] = surf.origin
            op['p'][3:6] = surf.axis
            op['p'][6] = surf.radius
        elif name == "Invert":
            self._emit_implicit(surf.surf)
            op['type'] = A.IMPL_NEG
        else:
            code = {"Union": A.IMPL_MIN, "Intersection": A.IMPL_MAX, "Difference": A.IMPL_SUB}[name]
            surfaces = list(surf.surfaces)
            if not surfaces:
                raise UnsupportedSceneError("empty %s" % name)
            self._emit_implicit(surfaces[0])
            for s in surfaces[1:]:
                self._emit_implicit(s)
                o2 = np.zeros((), dtype=A.implicit_op_dtype)
                o2['type'] = code
                self._implicit_ops.append(o2)
            return
        self._implicit_ops.append(op)

    # -- distortions ---------------------------------------------------------------
    def _distortion(self, dist):
        key = id(dist)
        if key in self._dist_index:
            return self._dist_index[key]
        name = _cls_name(dist, _DISTORTION_CLASSES)
        d = np.zeros((), dtype=A.distortion_dtype)
        if name == "SimpleTestZernikeJ7":
            d['type'] = A.DIST_ZERNIKE_J7
            d['p'][0] = dist.unit_radius
            d['p'][1] = dist.amplitude
        else:
            coefs = [tuple(dist[i]) for i in range(int(dist.n_coefs))]
            k_max = int(dist.k_max)
            if k_max > A.ZERNIKE_MAX_K:
                raise UnsupportedSceneError("Zernike order too high: k_max=%d > %d"
                                            % (k_max, A.ZERNIKE_MAX_K))
            tz, tg, operands = build_zernike_tapes(coefs)
            d['type'] = A.DIST_ZERNIKE
            d['p'][0] = dist.unit_radius
            d['n_coefs'] = len(coefs)
            d['coef_off'] = len(self._zcoefs)
            d['k_max'] = k_max
            for (j, n, m, k, value), (opR, opRp, opRr, opR_z) in zip(coefs, operands):
                zc = np.zeros((), dtype=A.zcoef_dtype)
                zc['j'], zc['n'], zc['m'], zc['k'], zc['value'] = j, n, m, k, value
                zc['opR'], zc['opRp'], zc['opRr'], zc['opR_z'] = opR, opRp, opRr, opR_z
                self._zcoefs.append(zc)
            for which, tape in (("z", tz), ("g", tg)):
                d['tape_%s_off' % which] = len(self._ztape)
                d['tape_%s_len' % which] = len(tape)
                for (kind, dst, a, b, c, dd, e) in tape:
                    t = np.zeros((), dtype=A.ztape_op_dtype)
                    t['kind'], t['dst'], t['a'], t['b'], t['c'], t['d'], t['e'] = kind, dst, a, b, c, dd, e
                    self._ztape.append(t)
        self._distortions.append(d)
        self._dist_index[key] = len(self._distortions) - 1
        return self._dist_index[key]

    # -- materials --------------------------------------------------------------------
    def _const_ntab(self, n_in, n_out, n_coat=1.0):
        nwl = len(self.wavelengths)
        off = len(self._ntab)
        for v in (n_in, n_out, n_coat):
            self._ntab.extend([complex(v)] * nwl)
        return off

    def _table_ntab(self, rows):
        nwl = len(self.wavelengths)
        off = len(self._ntab)
        for r in rows:
            r = np.asarray(r, dtype=np.complex128).reshape(-1)
            if r.shape[0] != nwl:
                raise ValueError("dispersion table has %d entries for %d wavelengths"
                                 % (r.shape[0], nwl))
            self._ntab.extend(complex(v) for v in r)
        return off

    def _material(self, mat):
        key = id(mat)
        if key in self._mat_index:
            return self._mat_index[key]
        name = _cls_name(mat, _MATERIAL_CLASSES)
        m = np.zeros((), dtype=A.material_dtype)
        m['para_model'] = A.PARA_DEFAULT
        m['ntab_off'] = 0
        p = m['p']
        wl = self.wavelengths
        if name == "OpaqueMaterial":
            m['type'] = A.MAT_OPAQUE
        elif name == "TransparentMaterial":
            m['type'] = A.MAT_TRANSPARENT
        elif name == "PECMaterial":
            m['type'] = A.MAT_PEC
        elif name == "PartiallyReflectiveMaterial":
            m['type'] = A.MAT_PARTIALLY_REFLECTIVE
            p[0] = mat._reflectivity
        elif name == "LinearPolarisingMaterial":
            m['type'] = A.MAT_LINEAR_POLARISING
        elif name == "WaveplateMaterial":
            m['type'] = A.MAT_WAVEPLATE
            if hasattr(mat, "retardance_"):  # host mirror keeps the complex value itself
                r = complex(mat.retardance_)
                p[0], p[1] = r.real, r.imag
            else:  # reference object: only the atan2 round-trip getter is public (quirk Q10)
                val = float(mat.retardance)
                p[0], p[1] = math.cos(val * 2 * math.pi), math.sin(val * 2 * math.pi)
            p[2:5] = mat.fast_axis
        elif name == "DielectricMaterial":
            m['type'] = A.MAT_DIELECTRIC
            m['para_model'] = A.PARA_SNELL
            m['ntab_off'] = self._const_ntab(mat.n_inside, mat.n_outside)
        elif name == "FullDielectricMaterial":
            m['type'] = A.MAT_FULL_DIELECTRIC
            m['para_model'] = A.PARA_SNELL  # inherits DielectricMaterial.eval_parabasal_ray_c
            m['ntab_off'] = self._const_ntab(mat.n_inside, mat.n_outside)
            p[0], p[1] = mat.reflection_threshold, mat.transmission_threshold
        elif name == "FullDielectricDispersiveMaterial":
            m['type'] = A.MAT_FULL_DIELECTRIC
            m['para_model'] = A.PARA_DEFAULT  # no override in the reference
            n_in = mat.dispersion_inside.evaluate_n(wl)
            n_out = mat.dispersion_outside.evaluate_n(wl)
            m['ntab_off'] = self._table_ntab([n_in, n_out, np.ones(len(wl))])
            p[0], p[1] = mat.reflection_threshold, mat.transmission_threshold
        elif name == "SingleLayerCoatedMaterial":
            m['type'] = A.MAT_COATED
            m['para_model'] = A.PARA_SNELL
            m['ntab_off'] = self._const_ntab(mat.n_inside, mat.n_outside, mat.n_coating)
            p[0], p[1], p[2] = mat.reflection_threshold, mat.transmission_threshold, mat.thickness
        elif name == "CoatedDispersiveMaterial":
            m['type'] = A.MAT_COATED
            m['para_model'] = A.PARA_SNELL
            # same values on_set_wavelengths stores (cmaterials.pyx:1220-1225); evaluated here
            # through the public curve API so flattening does not depend on call order
            m['ntab_off'] = self._table_ntab([mat.dispersion_inside.evaluate_n(wl),
                                              mat.dispersion_outside.evaluate_n(wl),
                                              mat.dispersion_coating.evaluate_n(wl)])
            p[0], p[1], p[2] = (mat.reflection_threshold, mat.transmission_threshold,
                                mat.coating_thickness)
        elif name == "DiffractionGratingMaterial":
            m['type'] = A.MAT_GRATING
            m['para_model'] = A.PARA_GRATING
            p[0], p[1], p[2] = mat.lines_per_mm, int(mat.order), mat.efficiency
            p[3:6] = mat.origin
        elif name == "CircularApertureMaterial":
            m['type'] = A.MAT_CIRC_APERTURE
            p[0], p[1], p[2], p[3] = mat.outer_radius, mat.radius, mat.edge_width, int(mat.invert)
            p[4:7] = mat.origin
        elif name == "RectangularApertureMaterial":
            m['type'] = A.MAT_RECT_APERTURE
            p[0], p[1], p[2], p[3] = mat.outer_width, mat.outer_height, mat.width, mat.height
            p[4], p[5] = mat.edge_width, int(mat.invert)
            p[6:9] = mat.origin
        elif name == "ResampleGaussletMaterial":
            # eval_child_ray_c only captures (cmaterials.pyx:1808-1821): no child ray.  The capture, the
            # Python callback and the append happen on the host between generations (core/tracer.py)
            m['type'] = A.MAT_OPAQUE
        else:
            raise UnsupportedSceneError("material class %s is not supported" % name)
        self._materials.append(m)
        self._mat_index[key] = len(self._materials) - 1
        return self._mat_index[key]

    # -- faces ----------------------------------------------------------------------------
    def _face_record(self, face, face_set_idx, traced):
        name = _cls_name(face, _FACE_CLASSES)
        f = np.zeros((), dtype=A.face_dtype)
        f['face_set'] = face_set_idx
        f['material'] = self._material(face.material) if traced else 0
        f['invert_normal'] = int(getattr(face, 'invert_normal', 0))
        f['shape_off'], f['shape_len'] = -1, 0
        f['base_face'] = -1
        f['tolerance'] = face.tolerance
        p = f['p']

        def shaped():
            f['shape_off'], f['shape_len'] = self._shape_program(face.shape)

        if name == "CircularFace":
            f['type'] = A.FACE_CIRCULAR
            p[:4] = (face.diameter, face.offset, face.z_plane, 1.0 if face.invert_normals else 0.0)
        elif name == "ShapedPlanarFace":
            f['type'] = A.FACE_SHAPED_PLANAR
            p[0] = face.z_height
            shaped()
        elif name == "ImplicitBoundedPlanarFace":
            f['type'] = A.FACE_IMPLICIT_PLANAR
            p[0:3] = face.target.origin
            p[3:6] = face.target.normal
            f['aux_off'] = len(self._implicit_ops)
            self._emit_implicit(face.boundary)
            f['aux_n'] = len(self._implicit_ops) - int(f['aux_off'])
        elif name == "ElipticalPlaneFace":
            f['type'] = A.FACE_ELLIPTICAL_PLANE
            p[:3] = (face.g_x, face.g_y, face.diameter)
        elif name == "RectangularFace":
            f['type'] = A.FACE_RECTANGULAR
            p[:4] = (face.length, face.width, face.offset, face.z_plane)
        elif name == "SphericalFace":
            f['type'] = A.FACE_SPHERICAL
            p[:3] = (face.diameter, face.curvature, face.z_height)
        elif name == "ShapedSphericalFace":
            f['type'] = A.FACE_SHAPED_SPHERICAL
            p[:2] = (face.curvature, face.z_height)
            shaped()
        elif name == "ExtrudedPlanarFace":
            f['type'] = A.FACE_EXTRUDED_PLANAR
            x1, y1, x2, y2 = float(face.x1), float(face.y1), float(face.x2), float(face.y2)
            p[:6] = (x1, y1, x2, y2, face.z1, face.z2)
            p[6:9] = _norm3((y1 - y2, x2 - x1, 0.0))  # calc_normal, cfaces.pyx:657-663
        elif name == "PolygonFace":
            f['type'] = A.FACE_POLYGON
            pts = np.asarray(face.xy_points, dtype=np.double).reshape(-1, 2)
            p[0] = face.z_plane
            f['aux_off'] = self._pool_add(pts)
            f['aux_n'] = pts.shape[0]
        elif name == "OrientedPolygonFace":
            f['type'] = A.FACE_ORIENTED_POLYGON
            normal = tuple(float(v) for v in face.normal)
            x_axis = tuple(float(v) for v in face.x_axis)
            p[0:3] = face.origin
            p[3:6] = normal
            p[6:9] = x_axis
            p[9:12] = _cross3(normal, x_axis)  # cfaces.pyx:1156,1169
            pts = np.asarray(face.xy_points, dtype=np.double).reshape(-1, 2)
            f['aux_off'] = self._pool_add(pts)
            f['aux_n'] = pts.shape[0]
        elif name == "OffAxisParabolicFace":
            f['type'] = A.FACE_OFFAXIS_PARABOLIC
            p[:3] = (face.EFL, face.diameter, face.height)
        elif name == "EllipsoidalFace":
            f['type'] = A.FACE_ELLIPSOIDAL
            p[:8] = (face.major, face.minor, face.x1, face.x2, face.y1, face.y2, face.z1, face.z2)
            rot, dt = _transform_rows(face.transform)
            irot, idt = _transform_rows(face.inverse_transform)
            f['aux_off'] = self._pool_add(np.concatenate([rot, dt, irot, idt]))
            f['aux_n'] = 24
        elif name == "SaddleFace":
            f['type'] = A.FACE_SADDLE
            p[:2] = (face.z_height, face.curvature)
            shaped()
        elif name == "CylindericalFace":
            f['type'] = A.FACE_CYLINDRICAL
            p[:2] = (face.z_height, face.radius)
            shaped()
        elif name == "AxiconFace":
            f['type'] = A.FACE_AXICON
            p[:2] = (face.z_height, face.gradient)
            shaped()
        elif name == "ConicRevolutionFace":
            f['type'] = A.FACE_CONIC
            p[:4] = (face.curvature, face.z_height, face.conic_const,
                     1.0 if face.invert_normals else 0.0)
            shaped()
        elif name == "AsphericFace":
            f['type'] = A.FACE_ASPHERIC
            p[:4] = (face.curvature, face.z_height, face.conic_const,
                     1.0 if face.invert_normals else 0.0)
            p[4:11] = (face.A4, face.A6, face.A8, face.A10, face.A12, face.A14, face.A16)
            p[11] = face.atol
            shaped()
        elif name == "ExtendedPolynomialFace":
            f['type'] = A.FACE_EXT_POLY
            if hasattr(face, "ext_poly_R"):  # host mirror stores R and beta as the reference does
                R, beta = face.ext_poly_R, face.ext_poly_beta
            else:  # reference object: only the derived getters are public (cfaces.pyx:2141-2153)
                R, beta = -float(face.curvature), float(face.conic_const) + 1.0
            coefs = np.asarray(face.coefs, dtype=np.double)
            if coefs.ndim != 2:
                raise ValueError("ExtendedPolynomialFace.coefs must be 2-D")
            p[:6] = (R, beta, face.norm_radius, face.z_height, getattr(face, "atol", 1.0e-8),
                     1.0 if face.invert_normals else 0.0)
            f['aux_off'] = self._pool_add(coefs)
            f['aux_n'], f['aux_m'] = coefs.shape
            shaped()
        elif name == "ExtrudedBezierFace":
            f['type'] = A.FACE_EXTRUDED_BEZIER
            if hasattr(face, "curves_array"):
                curves, z1, z2 = face.curves_array, face.z_height_1, face.z_height_2
            else:  # reference object: the members are private cdef attributes; the owner
                # (raypier.splines.Extruded_bezier, splines.py:145-160) passed them in
                o = face.owner
                curves, z1, z2 = o.control_points, o.z_height_1, o.z_height_2
            curves = np.ascontiguousarray(curves, dtype=np.double)
            if curves.ndim != 3 or curves.shape[1:] != (4, 2) or curves.shape[0] < 1:
                raise ValueError("ExtrudedBezierFace needs control points of shape (n, 4, 2)")
            pts = curves.reshape(-1, 2)
            # bounding box exactly as __cinit__ builds it (cfaces.pyx:852-865)
            p[0:6] = (z1, z2, pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max())
            f['aux_off'] = self._pool_add(curves)
            f['aux_n'] = curves.shape[0]
        elif name == "OBBTreeFace":
            f['type'] = A.FACE_MESH
            tree = face.obbtree
            if hasattr(tree, "points") and hasattr(tree, "cells"):
                points, cells = tree.points, tree.cells
            else:  # reference object: OBBTree.points / .cells are private cdef members (obbtree.pxd);
                # the owner that built the tree (raypier.meshes.STLFileMesh, meshes.py:49-67) still has them
                o = face.owner
                if not (hasattr(o, "mesh_points") and hasattr(o, "mesh_cells")):
                    raise UnsupportedSceneError(
                        "a genuine raypier OBBTree does not expose its mesh: give the owner mesh_points (N x 3) "
                        "and mesh_cells (M x 3) attributes, or build the face from raypier_optics_b200.core.obbtree")
                points, cells = o.mesh_points, o.mesh_cells
            p[0] = tree.tolerance
            f['aux_off'], f['aux_n'], f['aux_m'] = self._mesh_block(points, cells, tree.tolerance)
        elif name == "UVPatchFace":
            f['type'] = A.FACE_UVPATCH
            patch = face.patch
            # the tessellation the face traces first (cbezier.pyx:408-418).  u_res / v_res are private cdef
            # members of a genuine raypier UVPatchFace: take them from the face (host mirror) or its owner
            u_res = getattr(face, "u_res", None) or getattr(getattr(face, "owner", None), "u_res", None)
            v_res = getattr(face, "v_res", None) or getattr(getattr(face, "owner", None), "v_res", None)
            if not u_res or not v_res:
                raise UnsupportedSceneError(
                    "a genuine raypier UVPatchFace does not expose its mesh resolution: give its owner u_res / "
                    "v_res attributes, or build the face from raypier_optics_b200.core.cbezier")
            points, cells, uvs = patch.get_mesh(int(u_res), int(v_res))
            tree = face.obbtree
            f['aux_off'], f['aux_n'], f['aux_m'] = self._mesh_block(points, cells, tree.tolerance)
            ctrl = np.ascontiguousarray(patch.control_pts, dtype=np.double)
            N, M = int(patch.order_n), int(patch.order_m)
            if ctrl.shape != (N + 1, M + 1, 3):
                raise ValueError("patch control points must have shape (%d, %d, 3)" % (N + 1, M + 1))
            pname = type(patch).__name__
            if pname == "BezierPatch":
                kind = 0
                fct = math.factorial  # binomial(), cbezier.pyx:120-122
                a = [fct(N) / (fct(i) * fct(N - i)) for i in range(N + 1)]
                b = [fct(M) / (fct(i) * fct(M - i)) for i in range(M + 1)]
                udeg = vdeg = 0
            elif pname == "BSplinePatch":
                kind = 1
                a = np.asarray(patch.u_knots, dtype=np.double)
                b = np.asarray(patch.v_knots, dtype=np.double)
                udeg, vdeg = int(patch.u_degree), int(patch.v_degree)
                if len(a) < N + udeg + 2 or len(b) < M + vdeg + 2:
                    raise ValueError("a B-spline patch needs (degree + n + 2) knots per direction")
                if udeg > 8 or vdeg > 8:
                    raise UnsupportedSceneError("B-spline degree > 8")
            else:
                raise UnsupportedSceneError("patch class %s is not supported" % pname)
            off = self._pool_add(np.asarray(uvs, dtype=np.double).reshape(-1, 2))
            self._pool_add(ctrl)
            self._pool_add(a)
            self._pool_add(b)
            p[0:10] = (face.atol, 1.0 if face.invert_normals else 0.0, kind, N, M, udeg, vdeg, off, len(a), len(b))
        elif name == "DistortionFace":
            f['type'] = A.FACE_DISTORTION
            p[0] = face.accuracy
            f['aux_off'] = self._distortion(face.distortion)
            base = self._face_record(face.base_face, face_set_idx, traced=False)
            if int(base['type']) == A.FACE_DISTORTION:
                raise UnsupportedSceneError("nested DistortionFace")
            self._extra_faces.append(base)
            f['base_face'] = -len(self._extra_faces)  # patched in _finalise
            shaped()
        else:
            raise UnsupportedSceneError(
                "%s is outside the hot-path scope of this build (SURVEY.md section 8a-F)" % name)
        return f

    def _build(self):
        idx = 0
        for si, fs in enumerate(self.face_lists):
            rec = np.zeros((), dtype=A.face_set_dtype)
            rec['trans']['m'], rec['trans']['t'] = _transform_rows(fs.transform)
            rec['inv_trans']['m'], rec['inv_trans']['t'] = _transform_rows(fs.inverse_transform)
            rec['face_begin'] = idx
            for face in fs.faces:
                self._faces.append(self._face_record(face, si, traced=True))
                idx += 1
            rec['face_end'] = idx
            self._face_sets.append(rec)

    @staticmethod
    def _stack(records, dtype):
        if not records:
            return np.zeros(1, dtype=dtype)  # keep a valid pointer
        return np.array(records, dtype=dtype).reshape(-1)

    def _finalise(self):
        n_traced = len(self._faces)
        faces = self._faces + self._extra_faces
        for f in faces:
            if int(f['base_face']) < 0 and int(f['type']) == A.FACE_DISTORTION:
                f['base_face'] = n_traced + (-int(f['base_face']) - 1)
        self.faces = self._stack(faces, A.face_dtype)
        self.face_sets = self._stack(self._face_sets, A.face_set_dtype)
        self.materials = self._stack(self._materials, A.material_dtype)
        self.shape_ops = self._stack(self._shape_ops, A.shape_op_dtype)
        self.implicit_ops = self._stack(self._implicit_ops, A.implicit_op_dtype)
        self.distortions = self._stack(self._distortions, A.distortion_dtype)
        self.zcoefs = self._stack(self._zcoefs, A.zcoef_dtype)
        self.ztape = self._stack(self._ztape, A.ztape_op_dtype)
        self.ntab = (np.array(self._ntab, dtype=np.complex128) if self._ntab
                     else np.zeros(1, dtype=np.complex128))
        self.pool = (np.ascontiguousarray(np.concatenate(self._pool)) if self._pool_len
                     else np.zeros(1, dtype=np.double))
        wl = self.wavelengths if len(self.wavelengths) else np.zeros(1)
        self._wl_buf = np.ascontiguousarray(wl, dtype=np.double)
        s = A.rpx_scene()
        s.abi_version = A.RPX_ABI_VERSION
        s.n_traced_faces = n_traced
        s.n_faces = len(faces)
        s.n_face_sets = len(self._face_sets)
        s.n_materials = len(self._materials)
        s.n_shape_ops = len(self._shape_ops)
        s.n_implicit_ops = len(self._implicit_ops)
        s.n_distortions = len(self._distortions)
        s.n_zcoefs = len(self._zcoefs)
        s.n_ztape = len(self._ztape)
        s.n_wavelengths = len(self.wavelengths)
        s.n_ntab = len(self._ntab)
        s.n_pool = self._pool_len
        for name, arr in (("faces", self.faces), ("face_sets", self.face_sets),
                          ("materials", self.materials), ("shape_ops", self.shape_ops),
                          ("implicit_ops", self.implicit_ops), ("distortions", self.distortions),
                          ("zcoefs", self.zcoefs), ("ztape", self.ztape),
                          ("wavelengths", self._wl_buf), ("ntab", self.ntab), ("pool", self.pool)):
            setattr(s, name, arr.ctypes.data)
        self.c_scene = s

    @property
    def n_traced_faces(self):
        return self.c_scene.n_traced_faces

    def byref(self):
        return C.byref(self.c_scene)

    # -- (de)serialisation for golden fixtures ---------------------------------------
    TABLES = ("faces", "face_sets", "materials", "shape_ops", "implicit_ops", "distortions",
              "zcoefs", "ztape", "ntab", "pool", "wavelengths")

    def to_dict(self):
        d = {name: np.asarray(getattr(self, name)) for name in self.TABLES}
        s = self.c_scene
        d["counts"] = np.array([s.n_traced_faces, s.n_faces, s.n_face_sets, s.n_materials,
                                s.n_shape_ops, s.n_implicit_ops, s.n_distortions, s.n_zcoefs,
                                s.n_ztape, s.n_wavelengths, s.n_ntab, s.n_pool], dtype=np.int64)
        return d

    @classmethod
    def from_dict(cls, d):
        self = cls.__new__(cls)
        self.face_lists, self.all_faces = [], []
        dt = {"faces": A.face_dtype, "face_sets": A.face_set_dtype, "materials": A.material_dtype,
              "shape_ops": A.shape_op_dtype, "implicit_ops": A.implicit_op_dtype,
              "distortions": A.distortion_dtype, "zcoefs": A.zcoef_dtype, "ztape": A.ztape_op_dtype,
              "ntab": np.complex128, "pool": np.double, "wavelengths": np.double}
        for name in cls.TABLES:
            arr = np.ascontiguousarray(d[name])
            if arr.dtype != dt[name]:
                arr = np.ascontiguousarray(arr.view(np.uint8).reshape(-1)).view(dt[name]) \
                    if arr.dtype.itemsize == np.dtype(dt[name]).itemsize else arr.astype(dt[name])
            setattr(self, name, arr)
        c = [int(v) for v in d["counts"]]
        self._wl_buf = self.wavelengths if len(self.wavelengths) else np.zeros(1)
        s = A.rpx_scene()
        s.abi_version = A.RPX_ABI_VERSION
        (s.n_traced_faces, s.n_faces, s.n_face_sets, s.n_materials, s.n_shape_ops, s.n_implicit_ops,
         s.n_distortions, s.n_zcoefs, s.n_ztape, s.n_wavelengths, s.n_ntab, s.n_pool) = c
        for name, arr in (("faces", self.faces), ("face_sets", self.face_sets),
                          ("materials", self.materials), ("shape_ops", self.shape_ops),
                          ("implicit_ops", self.implicit_ops), ("distortions", self.distortions),
                          ("zcoefs", self.zcoefs), ("ztape", self.ztape),
                          ("wavelengths", self._wl_buf), ("ntab", self.ntab), ("pool", self.pool)):
            setattr(s, name, arr.ctypes.data)
        self.c_scene = s
        return self


def flatten_scene(face_lists, wavelengths):
    return Scene(face_lists, wavelengths)
