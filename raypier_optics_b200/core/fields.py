"""Host-side mirror of the gausslet entry points of ``raypier.core.fields``
(raypier/core/fields.py:196-277): same names, arguments and return values; the mode fit and the
N_ray x N_pt summation run on the GPU through librpx."""
import numpy

from .._abi import gausslet_dtype
from ..engine import get_engine
from .cfields import gausslet_modes, sum_gaussian_modes  # noqa: F401


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
