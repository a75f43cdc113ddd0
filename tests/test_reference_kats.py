"""The reference's own known-answer tests for the hot path, restated against the C ABI
(CUDA, ``-m gpu``) and against the plain-C oracle (CPU): test/test_cfaces.py,
test/test_cmaterials.py, test/test_ctracer.py, test/test_cdistortions.py and
test/test_gratings.py of the reference."""
import math

import numpy as np
import pytest

from raypier_optics_b200 import _abi as A

from backends import CudaBackend, OracleBackend, ray_power, ray_record, unit_scene


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request):
    if request.param == "oracle":
        return OracleBackend()
    return CudaBackend(request.getfixturevalue("engine"))


class AnOwner(object):
    def __init__(self, **kwds):
        self.__dict__.update(kwds)


# ---------------------------------------------------------------- test/test_cfaces.py
def test_circular_face_params_update(core):
    o = AnOwner(diameter=5.5, offset=6.6)
    c = core.cfaces.CircularFace(owner=o)
    assert c.params == ['diameter', 'offset']
    c.update()
    assert (c.diameter, c.offset) == (o.diameter, o.offset)


def test_circular_face_intersection(be, core):
    o = AnOwner(diameter=5.5, offset=6.6)
    c = core.cfaces.CircularFace(owner=o)
    c.update()
    sc = unit_scene(core, [c])
    # is_base_ray=0: the reference test predates the aperture argument (offset 6.6 would miss)
    assert be.face_intersect(sc, 0, (0, 0, -2), (0, 0, 2), 0) == 2.0            # test_cfaces.py:29-33
    assert be.face_intersect(sc, 0, (-1, 0, -1), (1, 0, 1), 0) == pytest.approx(math.sqrt(2.0), abs=1e-15)
    assert be.face_intersect(sc, 0, (6, 0, -2), (6, 0, 2), 1) == 2.0            # x - offset = -0.6: inside
    assert be.face_intersect(sc, 0, (0, 0, -2), (0, 0, 2), 1) <= 0.0            # outside the aperture: miss (:52-57)


def test_extruded_face(be, core):
    f = core.cfaces.ExtrudedPlanarFace(owner=AnOwner(), z1=-1, z2=3, x1=-2.0, y1=2.0, x2=2.0, y2=-2.0)
    sc = unit_scene(core, [f])
    assert be.face_intersect(sc, 0, (-5, 0, 0), (5, 0, 0), 1) == 5.0             # test_cfaces.py:66-68
    assert be.face_intersect(sc, 0, (-5, 0, 4), (5, 0, 4), 1) <= 0.0             # :70-72
    assert be.face_intersect(sc, 0, (-5, 0, -2), (5, 0, -0.1), 1) <= 0.0         # :74-76
    assert be.face_intersect(sc, 0, (-5, 2.1, -1), (5, 2.1, 1), 1) <= 0.0        # :78-80


# ---------------------------------------------------------------- test/test_ctracer.py
def test_struct_sizes(core):
    assert core.ctracer.get_ray_size() == core.ctracer.ray_dtype.itemsize == 188   # test_ctracer.py:15-17
    assert A.gausslet_dtype.itemsize == 668


def test_pec_reflection(be, core):
    f = core.cfaces.CircularFace(owner=AnOwner(diameter=10.0, offset=0.0), material=core.cmaterials.PECMaterial())
    sc = unit_scene(core, [f])
    ray = ray_record(origin=(-1, 0, -1), direction=(1, 0, 1), E_vector=(0, 1, 0), E1_amp=1.0, length=math.sqrt(2))
    kids = be.material_eval(sc, 0, ray, 7, (0, 0, 0), (0, 0, -1))
    assert len(kids) == 1
    assert np.allclose(kids[0]['direction'], (1, 0, -1), atol=1e-15)              # test_ctracer.py:101-108
    assert kids[0]['ray_type_id'] & A.REFL_RAY
    assert kids[0]['parent_idx'] == 7


# ---------------------------------------------------------------- test/test_cmaterials.py
def _P(z):
    return z.real ** 2 + z.imag ** 2


def test_convert_to_sp_conserves_power(core):
    from oracle import oracle as O
    ray = ray_record(origin=(-1., -2., -3.), direction=(1., 2., 3.), E_vector=(1., -2., 0.),
                     E1_amp=(1.0 + 2.0j), E2_amp=(3.0 + 4.0j))
    out = O.convert_to_sp(ray, (0.2, 0.3, -1.1))                                 # test_cmaterials.py:99-118
    assert _P(out['E1_amp']) + _P(out['E2_amp']) == pytest.approx(_P(1 + 2j) + _P(3 + 4j), abs=1e-12)
    same = O.convert_to_sp(ray, (1.0, 2.0, 3.0))                                 # normal incidence: unchanged
    assert same['E1_amp'] == ray[0]['E1_amp'] and same['E2_amp'] == ray[0]['E2_amp']


def _full_dielectric_scene(core, n_in, n_out, cls="FullDielectricMaterial", **kw):
    mat = getattr(core.cmaterials, cls)(n_inside=n_in, n_outside=n_out, reflection_threshold=-0.01,
                                        transmission_threshold=-0.01, **kw)
    f = core.cfaces.CircularFace(owner=AnOwner(diameter=10.0, offset=0.0), material=mat)
    return unit_scene(core, [f])


@pytest.mark.parametrize("cls,kw", [("FullDielectricMaterial", {}),
                                    ("SingleLayerCoatedMaterial", dict(n_coating=1.0, thickness=0.1))])
def test_fresnel_normal_incidence(be, core, cls, kw):
    n_out, n_in = 1.3, 2.0                                                       # test_cmaterials.py:256-301, 436-489
    if cls == "SingleLayerCoatedMaterial":
        kw = dict(kw, n_coating=n_out)  # a coating of the outside index is no coating
    sc = _full_dielectric_scene(core, n_in, n_out, cls, **kw)
    ray = ray_record(origin=(0., 0., -1.), direction=(0., 0., 3.), E_vector=(5., 0., 0.), E1_amp=(1 + 1j),
                     E2_amp=(2 + 0j), refractive_index=n_out, length=1.0)
    P_in = ray_power(ray[0])
    kids = be.material_eval(sc, 0, ray, 123, (0, 0, 0), (0, 0, -1), (0, -1, 0))
    assert len(kids) == 2
    R = ((n_in - n_out) / (n_in + n_out)) ** 2
    T = (n_in / n_out) * ((2 * n_out / (n_out + n_in)) ** 2)
    assert T + R == pytest.approx(1.0, abs=1e-12)
    assert ray_power(kids[0]) / P_in == pytest.approx(R, abs=1e-10)
    assert ray_power(kids[1]) / P_in == pytest.approx(T, abs=1e-10)
    assert kids[0]['refractive_index'] == n_out and kids[1]['refractive_index'] == n_in   # :483-484
    assert kids[0]['ray_type_id'] & 1 and not (kids[1]['ray_type_id'] & 1)


def test_brewster_angle(be, core):
    n_out, n_in = 1.2, 2.0                                                       # test_cmaterials.py:372-432
    theta_B = math.atan2(n_in, n_out)
    y, z = math.sin(theta_B), math.cos(theta_B)
    theta_inside = math.asin(n_out * math.sin(theta_B) / n_in)
    trans_direction = np.array((0.0, math.sin(theta_inside), math.cos(theta_inside)))
    sc = _full_dielectric_scene(core, n_in, n_out)
    ray = ray_record(origin=(0., -y, -z), direction=(0., y, z), E_vector=(0., 1., 0.), E1_amp=(1 + 1j),
                     E2_amp=0j, refractive_index=n_out, length=1.0)
    P_in = ray_power(ray[0])
    kids = be.material_eval(sc, 0, ray, 123, (0, 0, 0), (0, 0, -1), (0, -1, 0))
    assert len(kids) == 2
    assert ray_power(kids[0]) == pytest.approx(0.0, abs=1e-12)     # no reflected P-polarised power
    assert ray_power(kids[1]) == pytest.approx(P_in, abs=1e-12)
    d = np.asarray(ray[0]['direction']) - kids[0]['direction']
    assert abs(np.dot(d / np.linalg.norm(d), (0, 0, -1))) == pytest.approx(1.0, abs=1e-12)
    assert abs(np.dot(kids[1]['direction'], trans_direction)) == pytest.approx(1.0, abs=1e-12)


def test_snell_with_dispersion_curves(be, core):
    M = core.cmaterials                                                          # test_cmaterials.py:566-634
    n_in = 1.7
    mat = M.CoatedDispersiveMaterial(dispersion_inside=M.BaseDispersionCurve(0, np.array([n_in])),
                                     dispersion_outside=M.BaseDispersionCurve(0, np.array([1.0])),
                                     dispersion_coating=M.BaseDispersionCurve(0, np.array([1.3])),
                                     coating_thickness=0.05, reflection_threshold=-1.0,
                                     transmission_threshold=-1.0)
    f = core.cfaces.CircularFace(owner=AnOwner(diameter=10.0, offset=0.0), material=mat)
    sc = unit_scene(core, [f], wavelengths=(0.6, 0.8))
    th = math.radians(35.0)
    ray = ray_record(origin=(0., -math.sin(th), -math.cos(th)), direction=(0., math.sin(th), math.cos(th)),
                     E_vector=(1., 0., 0.), E1_amp=1.0, E2_amp=0.5, wavelength_idx=1, length=1.0)
    kids = be.material_eval(sc, 0, ray, 0, (0, 0, 0), (0, 0, -1))
    assert len(kids) == 2
    t = kids[1]['direction']
    sin_out = math.hypot(t[0], t[1]) / np.linalg.norm(t)
    assert n_in * sin_out == pytest.approx(1.0 * math.sin(th), abs=1e-12)        # Snell
    P_in = ray_power(ray[0])
    # power conservation with the obliquity factor the reference folds into the amplitudes
    assert ray_power(kids[0]) + ray_power(kids[1]) == pytest.approx(P_in, rel=1e-9)
    assert all(k['wavelength_idx'] == 1 for k in kids)


def test_waveplate_retardance(be, core):
    M = core.cmaterials                                                          # test_cmaterials.py:145-211
    for retard, expect in ((0.0, 1.0 + 0j), (0.25, 1j), (0.5, -1.0 + 0j)):
        mat = M.WaveplateMaterial(retardance=retard, fast_axis=(1.0, 0.0, 0.0))
        f = core.cfaces.CircularFace(owner=AnOwner(diameter=10.0, offset=0.0), material=mat)
        sc = unit_scene(core, [f])
        # E_vector along y: after convert_to_sp on the fast axis E1 is the component across it
        ray = ray_record(origin=(0., 0., -1.), direction=(0., 0., 1.), E_vector=(0., 1., 0.), E1_amp=1.0,
                         E2_amp=1.0, length=1.0)
        kids = be.material_eval(sc, 0, ray, 0, (0, 0, 0), (0, 0, -1))
        assert len(kids) == 1
        ratio = kids[0]['E1_amp'] / kids[0]['E2_amp']
        # equal magnitudes; the retardance shows up as the relative phase of E1 against E2
        assert abs(abs(ratio) - 1.0) < 1e-12
        assert abs(abs(np.angle(ratio)) % math.pi - (2 * math.pi * retard) % math.pi) < 1e-9
        assert abs(ratio) * abs(expect) == pytest.approx(1.0, abs=1e-12)


# ---------------------------------------------------------------- test/test_cdistortions.py
def test_ansi_index_table(core):
    # j -> (n, m) for j < 20 (test_cdistortions.py:54-93): OSA/ANSI single index
    D = core.cdistortions
    for j in range(20):
        n, m, k = D.eval_nmk(j)
        assert (n * (n + 2) + m) // 2 == j
        assert (n - abs(m)) % 2 == 0 and abs(m) <= n


def test_zernike_radial_closed_forms():
    from oracle import oracle as O                                              # test_cdistortions.py:34-51,164-180
    closed = {(2, 0): lambda r: 2 * r ** 2 - 1, (2, 2): lambda r: r ** 2, (3, 1): lambda r: 3 * r ** 3 - 2 * r,
              (3, 3): lambda r: r ** 3, (4, 0): lambda r: 6 * r ** 4 - 6 * r ** 2 + 1,
              (4, 2): lambda r: 4 * r ** 4 - 3 * r ** 2, (4, 4): lambda r: r ** 4,
              (5, 1): lambda r: 10 * r ** 5 - 12 * r ** 3 + 3 * r, (6, 0): lambda r: 20 * r ** 6 - 30 * r ** 4 + 12 * r ** 2 - 1}
    for (n, m), f in closed.items():
        half = n // 2
        k = half * (half + 1) + m
        for r in (0.0, 0.3, 0.77, 1.0):
            assert O.zernike("R", r, k, n, m, 64) == pytest.approx(f(r), abs=1e-12)
            if r > 0 and m > 0:                                                 # R == r * (R / r), :183-205
                assert r * O.zernike("R_over_r", r, k, n, m, 64) == pytest.approx(f(r), abs=1e-12)


def test_zernike_j7_matches_simple(be, core):
    D = core.cdistortions                                                        # test_cdistortions.py:100-142
    z = D.ZernikeDistortion(unit_radius=2.5, j7=0.3)
    s = D.SimpleTestZernikeJ7(unit_radius=2.5, amplitude=0.3)
    F, S = core.cfaces, core.cshapes
    shape = S.CircleShape(radius=5.0)
    faces = [F.DistortionFace(base_face=F.ShapedPlanarFace(shape=shape), distortion=d, shape=shape) for d in (z, s)]
    sc = unit_scene(core, faces)
    for (x, y) in ((0.0, 0.0), (0.4, -1.1), (1.7, 0.2), (-2.0, 1.0)):
        z0, g0 = be.distortion(sc, 0, x, y)
        z1, g1 = be.distortion(sc, 1, x, y)
        assert z0 == pytest.approx(z1, abs=1e-12)
        assert g0[2] == pytest.approx(z1, abs=1e-12)
        if (x, y) != (0.0, 0.0):
            assert np.allclose(g0, g1, atol=1e-12)
        else:  # at the origin the reference's R/r recursion is 0/0-free only for the sag
            assert np.isfinite(g0[2])


# ---------------------------------------------------------------- test/test_gratings.py
def test_grating_equation(be, core):
    M, F = core.cmaterials, core.cfaces                                          # test_gratings.py:16-44
    lines, order, wl = 600.0, 1, 0.8
    D = 1000.0 / lines
    mat = M.DiffractionGratingMaterial(lines_per_mm=lines, order=order)
    f = F.RectangularFace(owner=AnOwner(length=50.0, width=50.0, offset=0.0), material=mat, length=50.0, width=50.0)
    sc = unit_scene(core, [f], wavelengths=(wl,))
    for deg in (5.0, 20.0, 35.0):
        th = math.radians(deg)
        ray = ray_record(origin=(-math.sin(th), 0., -math.cos(th)), direction=(math.sin(th), 0., math.cos(th)),
                         E_vector=(0., 1., 0.), E1_amp=1.0, length=1.0)
        kids = be.material_eval(sc, 0, ray, 0, (0, 0, 0), (0, 0, -1), (1, 0, 0))
        assert len(kids) == 1
        d = kids[0]['direction']
        sin_in, sin_out = math.sin(th), d[0] / np.linalg.norm(d)
        # m*lambda = D*(sin_in - sin_out) with the reference's sign convention (k_x -= m*lambda/D)
        assert sin_out == pytest.approx(sin_in - order * wl / D, abs=1e-12)
