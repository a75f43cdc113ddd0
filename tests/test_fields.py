"""E-field summation: raypier/core/cfields.pyx sum_gaussian_modes / calc_mode_U /
evaluate_modes and the gausslet front end of raypier/core/fields.py.

CPU: the oracle restatement is bit-exact with the compiled reference (and with fields.py
imported in place from /root/reference); golden vectors made from the reference are committed for
the GPU box.  GPU: the CUDA path against the oracle and the golden vectors.

Tolerance: |dE| <= 1e-10 * max|E| (the north star's E-field tolerance); measured on B200: 2e-15.
The optical phase of a mode is ~1e6 rad (k = 2000 pi / lambda per mm times ~100 mm of path), so
ONE ulp of the path term is 1.2e-10 rad: the CUDA kernel therefore keeps the reference's
association and roundings for exactly the terms that carry that magnitude, and reduces the
argument of sin/cos with an exact-product FMA step.
"""
import os

import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import configs, scene as SC

TOL_FIELD_SUM = 1e-10
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fields_michelson.npz")


def michelson_output(core, n=600, seed=5):
    """Gausslets leaving the Michelson towards the output port (last generation) + a detector
    line across the fringes and a small grid off-axis."""
    from oracle import oracle as O
    cfg = configs.build(core, "config5", n=n, gausslets=True, seed=seed)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    gens, _ = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    g = gens[-1]
    out = g[g['base_ray']['direction'][:, 1] < -0.5]  # the beams travelling to -y (output port)
    assert len(out) > n
    xs = np.linspace(-4.0, 4.0, 41)
    line = np.stack([xs, np.full_like(xs, -14.0), np.zeros_like(xs)], axis=1)
    gx, gz = np.meshgrid(np.linspace(-2, 2, 5), np.linspace(-1.5, 1.5, 4))
    grid = np.stack([gx.ravel(), np.full(gx.size, -25.0), gz.ravel()], axis=1)
    return cfg, np.ascontiguousarray(out), np.concatenate([line, grid])


GOLDEN_HEX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fields_hexgrid.npz")


def hexgrid_points():
    """A line through the focus of configs.hex_grid_source and a few points beside it."""
    xs = np.linspace(-0.5, 0.5, 25)
    line = np.stack([xs, np.zeros_like(xs), np.full_like(xs, 79.0)], axis=1)
    return np.concatenate([line, np.array([[0.1, 0.2, 60.0], [-0.3, 0.1, 85.0], [0.0, 0.0, 80.0]])])


def rel_err(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


def test_oracle_fields_bit_exact_with_reference(refcore):
    from oracle import oracle as O
    F = O.reference_fields(refcore)
    if F is None:
        pytest.skip("raypier/core/fields.py not importable here")
    from raypier.core import cfields
    cfg, g, pts = michelson_output(refcore)
    base, rx, ry, rdx, rdy = F.evaluate_neighbours_gc(g)
    x, y, dx, dy = O.evaluate_neighbours_gc(g)
    for a, b in ((rx, x), (ry, y), (rdx, dx), (rdy, dy)):
        assert a.tobytes() == b.tobytes()
    for blending in (1.0, 0.7):
        rm = cfields.evaluate_modes(rx, ry, rdx, rdy, blending=blending)
        om = O.evaluate_modes(x, y, dx, dy, blending)
        assert rm.tobytes() == om.tobytes()
    rc = O.reference_collection(refcore, np.ascontiguousarray(g['base_ray']), cfg['wavelengths'])
    for t in (0.0, 3.5):
        want = cfields.sum_gaussian_modes(rc, rm, np.asarray(cfg['wavelengths']), pts, t)
        got = O.sum_gaussian_modes(np.ascontiguousarray(g['base_ray']), om, cfg['wavelengths'], pts, t)
        assert np.abs(want).max() > 1e-3
        assert got.tobytes() == want.tobytes(), "sum_gaussian_modes not bit-identical (time_ps=%g)" % t
    # the whole chain as the reference's public function runs it
    gc = refcore.ctracer.GaussletCollection.from_array(g.view(refcore.ctracer.gausslet_dtype))
    gc.wavelengths = np.asarray(cfg['wavelengths'])
    want = F.eval_Efield_from_gausslets(gc, pts, blending=0.7, time_ps=3.5)
    got = O.eval_Efield_from_gausslets(g, pts, cfg['wavelengths'], blending=0.7, time_ps=3.5)
    assert got.tobytes() == want.tobytes()


def test_oracle_reproduces_golden_fields():
    from oracle import oracle as O
    z = np.load(GOLDEN)
    g = z['gausslets'].view(A.gausslet_dtype).reshape(-1)
    modes = O.evaluate_modes(*O.evaluate_neighbours_gc(g), blending=float(z['blending']))
    assert modes.tobytes() == z['modes'].tobytes()
    E = O.eval_Efield_from_gausslets(g, z['points'], z['wavelengths'], float(z['blending']), float(z['time_ps']))
    assert E.tobytes() == z['E'].tobytes()


def test_field_is_linear_in_amplitude_and_additive_over_rays():
    """Size-independent properties of the summation (also used at full size on the GPU)."""
    from oracle import oracle as O
    z = np.load(GOLDEN)
    g = z['gausslets'].view(A.gausslet_dtype).reshape(-1)
    pts, wl = z['points'][:16], z['wavelengths']
    E = O.eval_Efield_from_gausslets(g, pts, wl)
    h = len(g) // 2
    E2 = O.eval_Efield_from_gausslets(g[:h], pts, wl) + O.eval_Efield_from_gausslets(g[h:], pts, wl)
    assert rel_err(E2, E) < 1e-13
    g3 = g.copy()
    g3['base_ray']['E1_amp'] *= 3.0
    g3['base_ray']['E2_amp'] *= 3.0
    assert rel_err(O.eval_Efield_from_gausslets(g3, pts, wl), 3.0 * E) < 1e-14


def test_oracle_ray_front_end_bit_exact_with_reference(refcore):
    """project_to_sphere / evaluate_neighbours / eval_Efield_from_rays (fields.py:50-111, 206-229): the
    oracle's loops against the reference's numpy, bit for bit."""
    from oracle import oracle as O
    F = O.reference_fields(refcore)
    if F is None:
        pytest.skip("raypier/core/fields.py not importable here")
    rays, nb = configs.hex_grid_source(n_side=15)
    wl, pts = np.array([1.0]), hexgrid_points()
    assert (nb < 0).any() and (nb >= 0).all(axis=1).sum() > 100
    for centre, radius in (((0, 0, 0), 0.0), ((0.3, -0.2, 80.0), 90.0)):
        proj = rays
        if radius:
            proj = F.project_to_sphere(rays.copy(), centre, radius)
            assert O.project_to_sphere(rays, centre, radius).tobytes() == proj.tobytes()
            assert np.abs(proj['origin'] - rays['origin']).max() > 1.0  # really moved
        want = F.evaluate_neighbours(proj.copy(), nb)
        got = O.evaluate_neighbours(proj, nb)
        for a, b in zip(want, got):
            assert a.tobytes() == b.tobytes()
        rc = O.reference_collection(refcore, rays.copy(), wl)
        rc.neighbours = nb
        E = F.eval_Efield_from_rays(rc, pts, wl, blending=0.8, time_ps=1.5, exit_pupil_offset=radius,
                                    exit_pupil_centre=centre)
        assert np.abs(E).max() > 1.0
        assert O.eval_Efield_from_rays(rays, nb, pts, wl, 0.8, 1.5, radius, centre).tobytes() == E.tobytes()
    # a ray that misses the sphere: the reference's own indexing raises, and so does the restatement
    far = rays.copy()
    far['origin'][3] = (500.0, 0.0, 0.0)
    far['direction'][3] = (0.0, 0.0, 1.0)
    rc = O.reference_collection(refcore, far.copy(), wl)
    rc.neighbours = nb
    with pytest.raises(IndexError):
        F.eval_Efield_from_rays(rc, pts, wl, exit_pupil_offset=90.0, exit_pupil_centre=(0.3, -0.2, 80.0))
    with pytest.raises(IndexError):
        O.eval_Efield_from_rays(far, nb, pts, wl, exit_pupil_offset=90.0, exit_pupil_centre=(0.3, -0.2, 80.0))


def test_oracle_reproduces_golden_hexgrid_fields():
    from oracle import oracle as O
    z = np.load(GOLDEN_HEX)
    rays = z['rays'].view(A.ray_dtype).reshape(-1)
    nb, wl, pts = z['neighbours'], z['wavelengths'], z['points']
    centre, radius, blending, time_ps = z['centre'], float(z['radius']), float(z['blending']), float(z['time_ps'])
    proj = O.project_to_sphere(rays, centre, radius)
    assert proj.tobytes() == z['projected'].tobytes()
    kept, x, y, dx, dy = O.evaluate_neighbours(proj, nb)
    for a, k in ((x, 'x'), (y, 'y'), (dx, 'dx'), (dy, 'dy')):
        assert a.tobytes() == z[k].tobytes()
    assert O.evaluate_modes(x, y, dx, dy, blending).tobytes() == z['modes'].tobytes()
    assert O.eval_Efield_from_rays(rays, nb, pts, wl, blending, time_ps).tobytes() == z['E_plain'].tobytes()
    assert O.eval_Efield_from_rays(rays, nb, pts, wl, blending, time_ps, radius, centre).tobytes() == z['E_sphere'].tobytes()


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_ray_front_end_matches_golden_reference(core, engine):
    """The mirrors of fields.project_to_sphere / evaluate_neighbours / cfields.evaluate_modes /
    fields.eval_Efield_from_rays (CUDA through the C ABI) against vectors made by the reference."""
    from raypier_optics_b200.core import cfields as CF, fields as FD
    z = np.load(GOLDEN_HEX)
    rays = z['rays'].view(A.ray_dtype).reshape(-1)
    nb, wl, pts = z['neighbours'], z['wavelengths'], z['points']
    centre, radius, blending, time_ps = z['centre'], float(z['radius']), float(z['blending']), float(z['time_ps'])
    want_proj = z['projected'].view(A.ray_dtype).reshape(-1)
    proj = FD.project_to_sphere(rays, centre, radius)
    assert len(proj) == len(want_proj)
    for name in ('direction', 'E_vector', 'E1_amp', 'E2_amp', 'wavelength_idx', 'parent_idx', 'ray_type_id'):
        assert np.array_equal(proj[name], want_proj[name])
    assert np.abs(proj['origin'] - want_proj['origin']).max() <= 1e-9 * np.abs(want_proj['origin']).max()
    assert np.abs(proj['accumulated_path'] - want_proj['accumulated_path']).max() <= 1e-9 * 100.0
    kept, x, y, dx, dy = FD.evaluate_neighbours(want_proj, nb)
    assert kept.tobytes() == want_proj[(nb >= 0).all(axis=1)].tobytes()
    for got, k in ((x, 'x'), (y, 'y'), (dx, 'dx'), (dy, 'dy')):
        assert got.shape == z[k].shape
        assert np.abs(got - z[k]).max() <= 1e-10 * max(np.abs(z[k]).max(), 1.0)
    modes = CF.evaluate_modes(z['x'], z['y'], z['dx'], z['dy'], blending=blending)
    scale = np.abs(z['modes']).max(axis=1, keepdims=True)
    assert float((np.abs(modes - z['modes']) / scale).max()) < 1e-10
    rc = core.ctracer.RayCollection.from_array(rays)
    rc.neighbours = nb
    E = FD.eval_Efield_from_rays(rc, pts, wl, blending=blending, time_ps=time_ps)
    assert rel_err(E, z['E_plain']) <= TOL_FIELD_SUM
    E = FD.eval_Efield_from_rays(rc, pts, wl, blending=blending, time_ps=time_ps, exit_pupil_offset=radius,
                                 exit_pupil_centre=centre)
    assert rel_err(E, z['E_sphere']) <= TOL_FIELD_SUM
    # the reference raises when a ray misses the exit-pupil sphere (its neighbour indexing breaks)
    far = rays.copy()
    far['origin'][3] = (500.0, 0.0, 0.0)
    far['direction'][3] = (0.0, 0.0, 1.0)
    rc = core.ctracer.RayCollection.from_array(far)
    rc.neighbours = nb
    with pytest.raises(IndexError):
        FD.eval_Efield_from_rays(rc, pts, wl, exit_pupil_offset=radius, exit_pupil_centre=centre)


@pytest.mark.gpu
def test_cuda_ray_front_end_matches_oracle_at_size(core, engine):
    """201 x 201 rays (38k with six neighbours) x 28 points against the oracle."""
    from oracle import oracle as O
    from raypier_optics_b200.core import fields as FD
    rays, nb = configs.hex_grid_source(n_side=201, pitch=0.04)
    wl, pts = np.array([0.8, 1.0]), hexgrid_points()
    rays['wavelength_idx'] = np.arange(len(rays)) % 2
    rc = core.ctracer.RayCollection.from_array(rays)
    rc.neighbours = nb
    for radius in (0.0, 95.0):
        want = O.eval_Efield_from_rays(rays, nb, pts, wl, 0.9, 0.5, radius, (0.0, 0.1, 80.0))
        got = FD.eval_Efield_from_rays(rc, pts, wl, blending=0.9, time_ps=0.5, exit_pupil_offset=radius,
                                       exit_pupil_centre=(0.0, 0.1, 80.0))
        assert np.abs(want).max() > 1.0
        assert rel_err(got, want) <= TOL_FIELD_SUM


@pytest.mark.gpu
def test_cuda_fields_match_golden_reference(engine):
    z = np.load(GOLDEN)
    g = z['gausslets'].view(A.gausslet_dtype).reshape(-1)
    fm = engine.field_prepare(g, z['wavelengths'], blending=float(z['blending']))
    try:
        modes = fm.modes
        scale = np.abs(z['modes']).max(axis=1, keepdims=True)
        assert float((np.abs(modes - z['modes']) / scale).max()) < 1e-10
        E = fm.evaluate(z['points'], float(z['time_ps']))
    finally:
        fm.free()
    err = rel_err(E, z['E'])
    assert err <= TOL_FIELD_SUM, "E-field rel err %.3e" % err


@pytest.mark.gpu
@pytest.mark.parametrize("time_ps", [0.0, 3.5])
def test_cuda_fields_match_oracle(core, engine, time_ps):
    from oracle import oracle as O
    from raypier_optics_b200.core import cfields as CF, fields as FD
    cfg, g, pts = michelson_output(core, n=900, seed=11)
    want = O.eval_Efield_from_gausslets(g, pts, cfg['wavelengths'], 0.8, time_ps)
    gc = core.ctracer.GaussletCollection.from_array(g)
    gc.wavelengths = cfg['wavelengths']
    got = FD.eval_Efield_from_gausslets(gc, pts, blending=0.8, time_ps=time_ps)
    assert got.shape == want.shape and got.dtype == np.complex128
    assert rel_err(got, want) <= TOL_FIELD_SUM
    # the exact sum_gaussian_modes signature: explicit modes with the base rays
    modes = O.evaluate_modes(*O.evaluate_neighbours_gc(g), blending=0.8)
    got2 = CF.sum_gaussian_modes(np.ascontiguousarray(g['base_ray']), modes, cfg['wavelengths'], pts, time_ps)
    assert rel_err(got2, want) <= TOL_FIELD_SUM
    # EFieldSummation: prepare once, evaluate point sets of any shape
    s = FD.EFieldSummation(gc, blending=0.8)
    E3 = s.evaluate(pts.reshape(-1, 1, 3), time_ps)
    assert E3.shape == (len(pts), 1, 3)
    assert rel_err(E3.reshape(-1, 3), want) <= TOL_FIELD_SUM


@pytest.mark.gpu
def test_cuda_field_properties_at_size(core, engine):
    """50k gausslets x 20k points (1e9 pairs; the oracle would need minutes): additivity over ray
    shards -- what the multi-GPU all-reduce relies on -- linearity, and a spot check of 8 points
    against the oracle."""
    from oracle import oracle as O
    cfg, g, _ = michelson_output(core, n=25000, seed=3)
    g = g[:50000]
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-4, 4, 20000), np.full(20000, -14.0), rng.uniform(-3, 3, 20000)], axis=1)
    wl = cfg['wavelengths']
    fm = engine.field_prepare(g, wl)
    E = fm.evaluate(pts)
    ms = fm.last_ms
    fm.free()
    assert ms > 0
    h = len(g) // 3
    parts = []
    for part in (g[:h], g[h:]):
        f = engine.field_prepare(part, wl)
        parts.append(f.evaluate(pts))
        f.free()
    assert rel_err(parts[0] + parts[1], E) < 1e-12
    want = O.eval_Efield_from_gausslets(g, pts[:8], wl)
    assert np.abs(E[:8] - want).max() / np.abs(E).max() <= TOL_FIELD_SUM
