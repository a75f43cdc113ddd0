"""Host-side mirror of ``raypier.core.cfields`` (raypier/core/cfields.pyx): the E-field at a set
of points as a sum of general astigmatic Gaussian modes.  The arithmetic runs in librpx
(``k_field_prepare`` / ``k_field_sum``); there is no CPU fallback."""
import numpy as np

from .._abi import gausslet_dtype, ray_dtype
from ..engine import get_engine


def _as_ray_array(rays):
    a = rays.copy_as_array() if hasattr(rays, "copy_as_array") else rays
    a = np.ascontiguousarray(a)
    if a.dtype.itemsize == gausslet_dtype.itemsize:
        return np.ascontiguousarray(a.view(gausslet_dtype)['base_ray'])
    return a.view(ray_dtype)


def sum_gaussian_modes(rays, modes, wavelengths, points, time_ps=0.0, device=0):
    """cfields.pyx:51-115: ``rays`` a RayCollection (or ray_dtype array) of N rays, ``modes`` an
    N x 3 complex array of (A, B, C), ``wavelengths`` in microns, ``points`` M x 3.
    Returns the M x 3 complex128 field."""
    eng = get_engine(device)
    fm = eng.field_prepare(_as_ray_array(rays), wavelengths, modes=modes)
    try:
        return fm.evaluate(points, time_ps)
    finally:
        fm.free()


def gausslet_modes(gausslets, blending=1.0, device=0):
    """The chain ``evaluate_neighbours_gc`` (fields.py:114-137) -> ``evaluate_modes``
    (cfields.pyx:217-228) for a gausslet_dtype array: N x 3 complex (A, B, C)."""
    eng = get_engine(device)
    g = gausslets.copy_as_array() if hasattr(gausslets, "copy_as_array") else gausslets
    g = np.ascontiguousarray(g).view(gausslet_dtype)
    fm = eng.field_prepare(g, [1.0], blending=blending)
    try:
        return fm.modes
    finally:
        fm.free()


def evaluate_modes(neighbour_x, neighbour_y, dx, dy, blending=1.0, device=0):
    """cfields.pyx:217-228: for N rays with the (x, y) of their six neighbours and the direction
    differences (dx, dy), all N x 6, the N x 3 complex (A, B, C) of the fitted astigmatic Gaussian modes."""
    return get_engine(device).evaluate_modes(neighbour_x, neighbour_y, dx, dy, blending)
