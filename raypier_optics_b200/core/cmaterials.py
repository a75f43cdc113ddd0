"""Host mirror of raypier/core/cmaterials.pyx: dispersion curves (evaluated on the
host at set-up time, exactly as the reference does in ``on_set_wavelengths``) and one
parameter-holder class per reference InterfaceMaterial.  ``eval_child_ray_c`` /
``eval_parabasal_ray_c`` of every class are CUDA device functions in librpx
(csrc/rpx_materials.cuh)."""
import math

import numpy as np

from .ctracer import InterfaceMaterial


# ---- dispersion formulas, cmaterials.pyx:107-149 -----------------------------------
def _nondispersive_0(wavelen, coefs):
    return coefs[0]


def _sellmeier_1(wavelen, coefs):
    n2 = 1.0
    wl2 = wavelen * wavelen
    n2 += coefs[0]
    for i in range((len(coefs) - 1) // 2):
        n2 += coefs[2 * i + 1] * wl2 / (wl2 - coefs[2 * i + 2] ** 2)
    return math.sqrt(n2)


def _sellmeier_2(wavelen, coefs):
    n2 = 1.0
    wl2 = wavelen * wavelen
    n2 += coefs[0]
    for i in range((len(coefs) - 1) // 2):
        n2 += coefs[2 * i + 1] * wl2 / (wl2 - coefs[2 * i + 2])
    return math.sqrt(n2)


def _sellmeier_3(wavelen, coefs):
    n2 = coefs[0]
    for i in range(1, len(coefs) - 1, 2):
        n2 += coefs[i] * (wavelen ** (coefs[i + 1]))
    return math.sqrt(n2)


def _sellmeier_5(wavelen, coefs):
    # The reference's sellmeier_5 has no ``return`` (cmaterials.pyx:143-149): a cdef
    # double function falling off the end yields 0.0.  Reproduced (quirk Q5).
    return 0.0


_CURVES = {0: _nondispersive_0, 1: _sellmeier_1, 2: _sellmeier_2, 3: _sellmeier_3, 5: _sellmeier_5}


class BaseDispersionCurve(object):
    """cmaterials.pyx:152-239"""

    def __init__(self, formula_id, coefs, absorption=0.0, wavelength_min=0.1,
                 wavelength_max=1000000.0):
        if formula_id not in _CURVES:
            raise ValueError("Unknown formula id (%d)" % (formula_id,))
        self.coefs = [float(c) for c in np.asarray(coefs, dtype=np.double).reshape(-1)]
        self.formula_id = formula_id
        self.absorption = absorption
        self.wavelength_min = wavelength_min
        self.wavelength_max = wavelength_max

    def evaluate_n(self, wavelen):
        """Complex refractive index at the given wavelengths in microns
        (c_evaluate_n, cmaterials.pyx:206-228)."""
        wavelen = np.asarray(wavelen, dtype=np.double).reshape(-1)
        curve = _CURVES[self.formula_id]
        n_imag = 0.0001 * self.absorption / (4 * math.pi)
        out = np.empty(wavelen.shape[0], dtype=np.complex128)
        for i, wvl in enumerate(wavelen):
            wvl = float(wvl)
            if (wvl < self.wavelength_min) or (wvl > self.wavelength_max):
                raise ValueError("Wavelength (%f) outside range of dispersion curve (%f -> %f)"
                                 % (wvl, self.wavelength_min, self.wavelength_max))
            out[i] = complex(curve(wvl, self.coefs), n_imag * wvl)
        return out


vacuum = BaseDispersionCurve(0, np.array([1.0, ]))


class OpaqueMaterial(InterfaceMaterial):
    """cmaterials.pyx:242-251"""


class TransparentMaterial(InterfaceMaterial):
    """cmaterials.pyx:254-278"""


class PECMaterial(InterfaceMaterial):
    """cmaterials.pyx:281-319"""


class PartiallyReflectiveMaterial(InterfaceMaterial):
    """cmaterials.pyx:322-397"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.reflectivity = kwds.get("reflectivity", 0.5)

    @property
    def reflectivity(self):
        return self._reflectivity

    @reflectivity.setter
    def reflectivity(self, val):
        if val < 0.0 or val > 1.0:
            raise ValueError("Reflectivity must be in range 0.0 to 1.0 (%s given)." % (val,))
        self._reflectivity = float(val)


class LinearPolarisingMaterial(InterfaceMaterial):
    """cmaterials.pyx:400-455"""


class WaveplateMaterial(InterfaceMaterial):
    """cmaterials.pyx:458-551"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.retardance = kwds.get("retardance", 0.25)
        self.fast_axis = kwds.get("fast_axis", (1.0, 0, 0))

    @property
    def retardance(self):
        val = math.atan2(self.retardance_.imag, self.retardance_.real) / (2 * math.pi)
        if val < 0:
            val += 1.0
        return val

    @retardance.setter
    def retardance(self, val):
        self.retardance_ = complex(math.cos(val * 2 * math.pi), math.sin(val * 2 * math.pi))

    @property
    def fast_axis(self):
        return self._fast_axis

    @fast_axis.setter
    def fast_axis(self, ax):
        self._fast_axis = (float(ax[0]), float(ax[1]), float(ax[2]))


class DielectricMaterial(InterfaceMaterial):
    """cmaterials.pyx:554-724"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.n_inside = kwds.get('n_inside', 1.5)
        self.n_outside = kwds.get('n_outside', 1.0)

    @property
    def n_inside(self):
        return self._n_inside

    @n_inside.setter
    def n_inside(self, v):
        self._n_inside = complex(v)

    @property
    def n_outside(self):
        return self._n_outside

    @n_outside.setter
    def n_outside(self, v):
        self._n_outside = complex(v)


class FullDielectricMaterial(DielectricMaterial):
    """cmaterials.pyx:727-872"""

    def __init__(self, **kwds):
        DielectricMaterial.__init__(self, **kwds)
        self.n_coating = kwds.get("n_coating", 1.0)
        self.thickness = kwds.get("thickness", 0.1)
        self.reflection_threshold = kwds.get('reflection_threshold', 0.1)
        self.transmission_threshold = kwds.get('transmission_threshold', 0.1)

    @property
    def n_coating(self):
        return self._n_coating

    @n_coating.setter
    def n_coating(self, v):
        self._n_coating = complex(v)


class FullDielectricDispersiveMaterial(InterfaceMaterial):
    """cmaterials.pyx:875-1013"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.dispersion_inside = kwds.get("dispersion_inside", vacuum)
        self.dispersion_outside = kwds.get("dispersion_outside", vacuum)
        self.reflection_threshold = kwds.get('reflection_threshold', 0.1)
        self.transmission_threshold = kwds.get('transmission_threshold', 0.1)

    def on_set_wavelengths(self):
        self.n_inside = self.dispersion_inside.evaluate_n(self._wavelengths)
        self.n_outside = self.dispersion_outside.evaluate_n(self._wavelengths)


class SingleLayerCoatedMaterial(FullDielectricMaterial):
    """cmaterials.pyx:1017-1182"""


class CoatedDispersiveMaterial(InterfaceMaterial):
    """cmaterials.pyx:1187-1434"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.reflection_threshold = kwds.get('reflection_threshold', 0.1)
        self.transmission_threshold = kwds.get('transmission_threshold', 0.1)
        self.dispersion_inside = kwds.get("dispersion_inside", vacuum)
        self.dispersion_outside = kwds.get("dispersion_outside", vacuum)
        self.dispersion_coating = kwds.get("dispersion_coating", vacuum)
        self.coating_thickness = kwds.get("coating_thickness", 0.1)
        self.n_inside = self.n_outside = self.n_coating = np.zeros(0, dtype=np.complex128)

    def on_set_wavelengths(self):
        self.n_inside = self.dispersion_inside.evaluate_n(self._wavelengths)
        self.n_outside = self.dispersion_outside.evaluate_n(self._wavelengths)
        self.n_coating = self.dispersion_coating.evaluate_n(self._wavelengths)


class DiffractionGratingMaterial(InterfaceMaterial):
    """cmaterials.pyx:1437-1599"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.lines_per_mm = kwds.get("lines_per_mm", 1000)
        self.order = int(kwds.get("order", 1))
        self.efficiency = kwds.get("efficiency", 1.0)
        o = kwds.get("origin", (0.0, 0.0, 0.0))
        self.origin = (float(o[0]), float(o[1]), float(o[2]))


class CircularApertureMaterial(InterfaceMaterial):
    """cmaterials.pyx:1602-1674"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.outer_radius = kwds.get("outer_radius", 25.0)
        self.radius = kwds.get("radius", 15.0)
        self.edge_width = kwds.get("edge_width", 1.0)
        o = kwds.get("origin", (0.0, 0.0, 0.0))
        self.origin = (float(o[0]), float(o[1]), float(o[2]))
        self.invert = int(kwds.get("invert", 0))


class RectangularApertureMaterial(InterfaceMaterial):
    """cmaterials.pyx:1677-1763"""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        self.outer_width = kwds.get("outer_width", 15.0)
        self.outer_height = kwds.get("outer_height", 20.0)
        self.width = kwds.get("width", 5.0)
        self.height = kwds.get("height", 10.0)
        self.edge_width = kwds.get("edge_width", 1.0)
        o = kwds.get("origin", (0.0, 0.0, 0.0))
        self.origin = (float(o[0]), float(o[1]), float(o[2]))
        self.invert = int(kwds.get("invert", 0))


class ResampleGaussletMaterial(InterfaceMaterial):
    """cmaterials.pyx:1766-1831 -- a pseudo-material: gausslets that reach it are captured and, once the
    generation is complete, handed to ``eval_func`` (a callable taking a GaussletCollection and returning
    the new outgoing GaussletCollection), whose result is appended to the new generation
    (ctracer.pyx:2274-2278).  On the device the face absorbs; the capture, the callback and the append
    are done by the host between generations (core/tracer.py)."""

    def __init__(self, **kwds):
        InterfaceMaterial.__init__(self)
        from .ctracer import GaussletCollection
        self.capture_count = 0
        self.captured_rays = GaussletCollection(kwds.get("size", 2))
        self._evaluation_func = None
        func = kwds.get("eval_func", None)
        if func is not None:
            self.eval_func = func

    @property
    def eval_func(self):
        return self._evaluation_func

    @eval_func.setter
    def eval_func(self, obj):
        if not callable(obj):
            raise ValueError("The eval_func property must be a callable.")
        self._evaluation_func = obj

    def is_decomp_material(self):
        return True
