"""Host mirror of raypier/core/cfaces.pyx: one class per reference Face type, same
constructor keywords and defaults.  Parameter holders only -- ``intersect_c`` /
``compute_normal_c`` of every type are CUDA device functions in librpx
(csrc/rpx_faces.cuh); scene.py flattens these objects into the face table."""
import numpy as np

from .ctracer import Face, Shape, Transform


class ShapedFace(Face):
    """cfaces.pyx:50-57"""

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.shape = kwds.get("shape", Shape())
        self.invert_normals = int(kwds.get('invert_normals', 0))


class CircularFace(Face):
    """cfaces.pyx:140-191"""
    params = ['diameter', 'offset']

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.diameter = kwds.get('diameter', 0.0)
        self.offset = kwds.get('offset', 0.0)
        self.z_plane = kwds.get('z_plane', 0.0)
        self.invert_normals = kwds.get("invert_normals", False)


class ShapedPlanarFace(ShapedFace):
    """cfaces.pyx:194-242"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)


class ImplicitBoundedFace(Face):
    pass


class ImplicitBoundedPlanarFace(ImplicitBoundedFace):
    """cfaces.pyx:251-310"""

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        from .cimplicit_surfs import NullSurface, Plane
        target = kwds.get('target', None)
        if target is None:
            target = Plane()
        if 'origin' in kwds:
            target.origin = kwds['origin']
        if 'normal' in kwds:
            target.normal = kwds['normal']
        self.target = target
        self.boundary = kwds.get('boundary', NullSurface())


class ElipticalPlaneFace(Face):
    """cfaces.pyx:313-351"""
    params = ['diameter']

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.g_x = kwds.get('g_x', 0.0)
        self.g_y = kwds.get('g_y', 0.0)
        self.diameter = kwds.get('diameter', 0.0)


class RectangularFace(Face):
    """cfaces.pyx:354-407"""
    params = ['length', 'width', 'offset']

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.z_plane = kwds.get('z_plane', 0.0)
        self.width = kwds.get("width", 2.0)
        self.length = kwds.get("length", 5.0)
        self.offset = kwds.get("offset", 0.0)


class SphericalFace(Face):
    """cfaces.pyx:410-498"""
    params = ['diameter', ]

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.diameter = kwds.get('diameter', 0.0)
        self.z_height = kwds.get('z_height', 0.0)
        self.curvature = kwds.get('curvature', 25.0)


class ShapedSphericalFace(ShapedFace):
    """cfaces.pyx:501-608"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.curvature = kwds.get("curvature", 100.0)


class ExtrudedPlanarFace(Face):
    """cfaces.pyx:611-709"""

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.x1 = float(kwds.get('x1', 0))
        self.y1 = float(kwds.get('y1', 0))
        self.x2 = float(kwds.get('x2', 0))
        self.y2 = float(kwds.get('y2', 0))
        self.z1 = float(kwds.get('z1', 0))
        self.z2 = float(kwds.get('z2', 0))


class ExtrudedBezierFace(Face):
    """cfaces.pyx:795-1046: an extrusion along z of a chain of cubic Bezier segments
    (``beziercurves`` has shape (n, 4, 2)), as built by raypier.splines.Extruded_bezier
    (splines.py:145-160) for CPCs and dielectric troughs."""

    def __init__(self, beziercurves=None, z_height_1=0, z_height_2=0, **kwds):
        Face.__init__(self, **kwds)
        self.curves_array = np.ascontiguousarray(beziercurves, dtype=np.float64)
        if self.curves_array.ndim != 3 or self.curves_array.shape[1:] != (4, 2):
            raise ValueError("beziercurves must have shape (n, 4, 2)")
        self.z_height_1 = float(z_height_1)
        self.z_height_2 = float(z_height_2)


class PolygonFace(Face):
    """cfaces.pyx:1077-1118"""

    def __init__(self, z_plane=0.0, xy_points=[[]], **kwds):
        Face.__init__(self, **kwds)
        self.z_plane = z_plane
        self.xy_points = xy_points

    @property
    def xy_points(self):
        return self._xy_points

    @xy_points.setter
    def xy_points(self, pts):
        self._xy_points = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)


class OrientedPolygonFace(Face):
    """cfaces.pyx:1121-1219 (normal and x_axis are normalised by their setters)"""

    def __init__(self, **kwds):
        face_kw = {k: kwds.pop(k) for k in ("owner", "tolerance", "max_length", "material",
                                            "invert_normal") if k in kwds}
        Face.__init__(self, **face_kw)
        self._origin = (0.0, 0.0, 0.0)
        self._normal = (0.0, 0.0, 0.0)
        self._x_axis = (0.0, 0.0, 0.0)
        self._xy_points = np.zeros((0, 2))
        for k in kwds:
            setattr(self, k, kwds[k])

    @staticmethod
    def _norm(v):
        import math
        x, y, z = float(v[0]), float(v[1]), float(v[2])
        m = math.sqrt(x * x + y * y + z * z)
        return (x / m, y / m, z / m)

    @property
    def origin(self):
        return self._origin

    @origin.setter
    def origin(self, v):
        self._origin = (float(v[0]), float(v[1]), float(v[2]))

    @property
    def normal(self):
        return self._normal

    @normal.setter
    def normal(self, v):
        self._normal = self._norm(v)

    @property
    def x_axis(self):
        return self._x_axis

    @x_axis.setter
    def x_axis(self, v):
        self._x_axis = self._norm(v)

    @property
    def xy_points(self):
        return self._xy_points

    @xy_points.setter
    def xy_points(self, xy):
        xy = np.asarray(xy).astype(np.double)
        if xy.ndim != 2 or xy.shape[1] != 2:
            raise ValueError("XY points must be an array with shape (N,2)")
        self._xy_points = xy


class OffAxisParabolicFace(Face):
    """cfaces.pyx:1224-1317.  As in the reference, EFL / diameter / height are plain public
    attributes: the class has no __cinit__ of its own, so constructor keywords are swallowed by
    Face and the owner assigns the attributes afterwards (raypier/parabolics.py:43-47)."""

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        self.EFL = 0.0
        self.diameter = 0.0
        self.height = 0.0


class EllipsoidalFace(Face):
    """cfaces.pyx:1320-1426"""

    def __init__(self, **kwds):
        Face.__init__(self, **kwds)
        # public attributes without constructor keywords in the reference (cfaces.pyx:1320-1342)
        self.major = 0.0
        self.minor = 0.0
        for b in ('x1', 'x2', 'y1', 'y2', 'z1', 'z2'):
            setattr(self, b, 0.0)
        self.transform = Transform()
        self.inverse_transform = Transform()

    def update(self):
        Face.update(self)
        owner = self.owner
        self.sync_transform(owner.ellipse_trans)
        self.major, self.minor = owner.axes
        self.x1, self.x2 = owner.X_bounds
        self.y1, self.y2 = owner.Y_bounds
        self.z1, self.z2 = owner.Z_bounds

    def sync_transform(self, vtk_trans):
        m = vtk_trans.matrix
        rot = [[m.get_element(i, j) for j in range(3)] for i in range(3)]
        dt = [m.get_element(i, 3) for i in range(3)]
        self.transform = Transform(rotation=rot, translation=dt)
        m = vtk_trans.linear_inverse.matrix
        rot = [[m.get_element(i, j) for j in range(3)] for i in range(3)]
        dt = [m.get_element(i, 3) for i in range(3)]
        self.inverse_transform = Transform(rotation=rot, translation=dt)


class SaddleFace(ShapedFace):
    """cfaces.pyx:1429-1511"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get("z_height", 0.0)
        self.curvature = kwds.get("curvature", 0.0)


class CylindericalFace(ShapedFace):
    """cfaces.pyx:1516-1604"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.radius = kwds.get("radius", 100.0)


class AxiconFace(ShapedFace):
    """cfaces.pyx:1607-1692"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.gradient = kwds.get('gradient', 0.0)


class ConicRevolutionFace(ShapedFace):
    """cfaces.pyx:1751-1837"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.conic_const = kwds.get('conic_const', 0.0)
        self.curvature = kwds.get('curvature', 10.0)


class AsphericFace(ShapedFace):
    """cfaces.pyx:1885-2025"""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.conic_const = kwds.get('conic_const', 0.0)
        self.curvature = kwds.get('curvature', 25.0)
        for name in ('A4', 'A6', 'A8', 'A10', 'A12', 'A14', 'A16'):
            setattr(self, name, kwds.get(name, 0.0))
        self.atol = kwds.get("atol", 1.0e-8)


class ExtendedPolynomialFace(ShapedFace):
    """cfaces.pyx:2130-2320.  Stores R = -curvature and beta = conic_const + 1 like the
    reference's ``extpoly_t`` does (properties at cfaces.pyx:2141-2153)."""

    def __init__(self, **kwds):
        ShapedFace.__init__(self, **kwds)
        self.z_height = kwds.get('z_height', 0.0)
        self.conic_const = kwds.get('conic_const', 0.0)
        self.curvature = kwds.get('curvature', 0.0)
        self.nterms = kwds.get('nterms', 0.0)
        self.norm_radius = kwds.get('norm_radius', 100.0)
        self.coefs = kwds.get('coefs', np.array([[0.0]]))
        self.atol = kwds.get("atol", 1.0e-8)

    @property
    def curvature(self):
        return -self.ext_poly_R

    @curvature.setter
    def curvature(self, v):
        self.ext_poly_R = -float(v)

    @property
    def conic_const(self):
        return self.ext_poly_beta - 1

    @conic_const.setter
    def conic_const(self, v):
        self.ext_poly_beta = float(v) + 1.0

    @property
    def coefs(self):
        return self._coefs

    @coefs.setter
    def coefs(self, coefs):
        self._coefs = np.ascontiguousarray(coefs, dtype=np.double)


class DistortionFace(ShapedFace):
    """cfaces.pyx:2323-2438"""

    def __init__(self, **kwds):
        face = kwds.get("base_face")
        if "shape" not in kwds:
            kwds = dict(kwds, shape=face.shape)
        ShapedFace.__init__(self, **kwds)
        self.base_face = face
        self.distortion = kwds.get("distortion")
        self.accuracy = kwds.get("accuracy", 1e-6)
