"""Host mirror of raypier/core/cimplicit_surfs.pyx (parameter holders; evaluated on
the GPU as an RPN program, see scene.py)."""
import math

from .ctracer import ImplicitSurface


def _norm(v):
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    m = math.sqrt(x * x + y * y + z * z)
    return (x / m, y / m, z / m)


class NullSurface(ImplicitSurface):
    """cimplicit_surfs.pyx:23-25"""


class Plane(ImplicitSurface):
    """cimplicit_surfs.pyx:27-58 (the normal is normalised by its setter)"""

    def __init__(self, **kwds):
        self.origin = kwds.get('origin', (0.0, 0.0, 0.0))
        self.normal = kwds.get('normal', (1.0, 0.0, 0.0))

    @property
    def origin(self):
        return self._origin

    @origin.setter
    def origin(self, o):
        self._origin = (float(o[0]), float(o[1]), float(o[2]))

    @property
    def normal(self):
        return self._normal

    @normal.setter
    def normal(self, n):
        self._normal = _norm(n)


class Sphere(ImplicitSurface):
    """cimplicit_surfs.pyx:60-81"""

    def __init__(self, **kwds):
        c = kwds.get('centre', (0.0, 0.0, 0.0))
        self.centre = (float(c[0]), float(c[1]), float(c[2]))
        self.radius = kwds.get('radius', 1.0)


class Cylinder(ImplicitSurface):
    """cimplicit_surfs.pyx:83-120 (the axis is normalised by its setter)"""

    def __init__(self, **kwds):
        o = kwds.get('origin', (0.0, 0.0, 0.0))
        self.origin = (float(o[0]), float(o[1]), float(o[2]))
        self.axis = kwds.get('axis', (0., 0., 1.))
        self.radius = kwds.get('radius', 10.0)

    @property
    def axis(self):
        return self._axis

    @axis.setter
    def axis(self, v):
        self._axis = _norm(v)


class Invert(ImplicitSurface):
    def __init__(self, surf):
        self.surf = surf


class Union(ImplicitSurface):
    """cimplicit_surfs.pyx:130-176: min over the member surfaces"""

    def __init__(self, *args):
        self.surfaces = list(args)


class Intersection(Union):
    """max over the member surfaces"""


class Difference(Union):
    """first minus the rest"""
