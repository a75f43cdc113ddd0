"""Host-side mirror of the field entry points of ``raypier.core.fields`` (raypier/core/fields.py:50-277):
same names, arguments and return values; projection, neighbour geometry, mode fit and the
N_ray x N_pt summation run on the GPU through librpx."""
import numpy

from .._abi import gausslet_dtype, ray_dtype
from ..engine import get_engine
from .cfields import evaluate_modes as evaluate_modes_c  # noqa: F401
from .cfields import gausslet_modes, sum_gaussian_modes  # noqa: F401


def _ray_array(rays):
    a = rays.copy_as_array() if hasattr(rays, "copy_as_array") else rays
    return numpy.ascontiguousarray(a).view(ray_dtype)


def project_to_sphere(rays, centre=(0, 0, 0), radius=10.0, device=0):
    """fields.py:50-77: project the rays (an array of ray_t dtype) back to their intercept with the sphere at
    ``centre`` / ``radius``; returns the rays that meet it (a new array, like ``rays[selector]``) with
    origin and accumulated_path updated."""
    eng = get_engine(device)
    dev = eng.upload(_ray_array(rays))
    try:
        selector = eng.project_to_sphere(dev, centre, radius)
        return eng.download(dev)[selector]
    finally:
        dev.free()


def evaluate_neighbours(rays, neighbours_idx, device=0):
    """fields.py:80-111: -> (rays[mask], x, y, dx, dy), x .. dy of shape (n_kept, 6): the neighbours of
    every ray projected onto the plane through its origin, in its (E, H) basis, and the changes of
    direction; rays without six neighbours are dropped."""
    eng = get_engine(device)
    r = _ray_array(rays)
    fm, (x, y, dx, dy) = eng.field_prepare_neighbours(r, neighbours_idx, [1.0], want_xy=True)
    fm.free()
    mask = (numpy.asarray(neighbours_idx) >= 0).all(axis=1)
    return r[mask], x, y, dx, dy


def eval_Efield_from_rays(ray_collection, points, wavelengths, blending=1.0, time_ps=0.0, exit_pupil_offset=0.0,
                          exit_pupil_centre=(0.0, 0.0, 0.0), device=0):
    """fields.py:206-229: the E-field of a RayCollection whose rays know their neighbours
    (``ray_collection.neighbours``, ctracer.pyx:1084-1131) at the N x 3 ``points``.  Like the reference,
    a ray that misses the exit-pupil sphere makes the neighbour indices meaningless: IndexError."""
    eng = get_engine(device)
    rays = _ray_array(ray_collection)
    neighbours_idx = ray_collection.neighbours
    dev = eng.upload(rays)
    try:
        if exit_pupil_offset:
            selector = eng.project_to_sphere(dev, exit_pupil_centre, exit_pupil_offset)
            if not selector.all():  # numpy raises on rays[selector] indexed with the full-length neighbour mask
                raise IndexError("boolean index did not match indexed array: %d of %d rays miss the exit pupil sphere"
                                 % (int((~selector).sum()), len(selector)))
        fm = eng.field_prepare_neighbours(dev, neighbours_idx, wavelengths, blending=blending)
    finally:
        dev.free()
    try:
        return fm.evaluate(numpy.ascontiguousarray(points, dtype=numpy.double).reshape(-1, 3), time_ps)
    finally:
        fm.free()


def _gausslet_array(gausslet_collection):
    g = gausslet_collection.copy_as_array() if hasattr(gausslet_collection, "copy_as_array") \
        else gausslet_collection
    return numpy.ascontiguousarray(g).view(gausslet_dtype)


def ExtractGamma(gausslet_collection, blending=1.0):
    """fields.py:196-203 ("used in testing"): the fitted (A, B, C) of every gausslet."""
    return gausslet_modes(_gausslet_array(gausslet_collection), blending=blending)


class EFieldSummation(object):
    """fields.py:206-249: convert the gausslets to Gaussian-mode parameters once (on the device),
    then evaluate the field for as many point sets as needed."""

    def __init__(self, gausslet_collection, wavelengths=None, blending=1.0, device=0):
        if wavelengths is None:
            wavelengths = numpy.asarray(gausslet_collection.wavelengths)
        if wavelengths is None:
            raise ValueError("No wavelengths supplied")
        self.wavelengths = wavelengths
        self.gc = _gausslet_array(gausslet_collection)
        self._fm = get_engine(device).field_prepare(self.gc, wavelengths, blending=blending)

    @property
    def modes(self):
        return self._fm.modes

    def evaluate(self, points, time_ps=0.0):
        """E-field at ``points`` (any shape ending in 3); returns the same shape, complex128."""
        points = numpy.ascontiguousarray(points)
        shape = points.shape
        E = self._fm.evaluate(points.reshape(-1, 3), time_ps)
        E.shape = shape
        return E


def eval_Efield_from_gausslets(gausslet_collection, points, wavelengths=None, blending=1.0, time_ps=0.0,
                               device=0, **kwds):
    """fields.py:252-277: the vector E-field (N x 3 complex128) of a GaussletCollection at the
    N x 3 ``points``."""
    if wavelengths is None:
        wavelengths = numpy.asarray(gausslet_collection.wavelengths)
    fm = get_engine(device).field_prepare(_gausslet_array(gausslet_collection), wavelengths, blending=blending)
    try:
        return fm.evaluate(numpy.ascontiguousarray(points, dtype=numpy.double).reshape(-1, 3), time_ps)
    finally:
        fm.free()
