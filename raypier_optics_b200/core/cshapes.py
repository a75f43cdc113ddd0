"""Host mirror of raypier/core/cshapes.pyx: 2-D aperture shapes (parameter holders;
the point-inside tests run on the GPU, flattened to an RPN program by scene.py)."""
import numpy as np

from .ctracer import Shape


class LogicalOpShape(Shape):
    def __and__(self, other):
        return BooleanAND(self, other)

    def __or__(self, other):
        return BooleanOR(self, other)

    def __xor__(self, other):
        return BooleanXOR(self, other)

    def __invert__(self):
        return InvertShape(self)


class InvertShape(LogicalOpShape):
    """cshapes.pyx:34-44"""

    def __init__(self, shape):
        self.shape = shape


class BooleanShape(LogicalOpShape):
    def __init__(self, shape1, shape2):
        self.shape1 = shape1
        self.shape2 = shape2


class BooleanAND(BooleanShape):
    pass


class BooleanOR(BooleanShape):
    pass


class BooleanXOR(BooleanShape):
    pass


class BasicShape(LogicalOpShape):
    def __init__(self, **kwds):
        if "centre" in kwds:
            self.centre = kwds["centre"]
        else:
            self.centre = (kwds.get("centre_x", 0.0), kwds.get("centre_y", 0.0))

    @property
    def centre(self):
        return (self.centre_x, self.centre_y)

    @centre.setter
    def centre(self, v):
        self.centre_x = float(v[0])
        self.centre_y = float(v[1])


class CircleShape(BasicShape):
    """cshapes.pyx:102-116: inside iff dx^2 + dy^2 < radius^2"""

    def __init__(self, **kwds):
        BasicShape.__init__(self, **kwds)
        self.radius = kwds.get("radius", 1.0)


class RectangleShape(BasicShape):
    """cshapes.pyx:119-136"""

    def __init__(self, **kwds):
        BasicShape.__init__(self, **kwds)
        self.width = kwds.get("width", 5.0)
        self.height = kwds.get("height", 7.0)


class PolygonShape(BasicShape):
    """cshapes.pyx:139-168"""

    def __init__(self, **kwds):
        BasicShape.__init__(self, **kwds)
        if "coordinates" in kwds:
            self.coordinates = kwds["coordinates"]

    @property
    def coordinates(self):
        return self._coordinates

    @coordinates.setter
    def coordinates(self, val):
        self._coordinates = np.ascontiguousarray(val, dtype=np.double).reshape(-1, 2)
