"""Capture planes: select_ray_intersections / select_gausslet_intersections
(raypier/core/ctracer.pyx:1981-2058), the filter behind probes.py RayCapturePlane /
GaussletCapturePlane (:119-143).

CPU (-m "not gpu"): the oracle's restatement is bit-exact with the real reference functions.
GPU (-m gpu): rpx_capture (device filter over resident generations, and the uploaded-collection
mirror) agrees with the oracle -- order and integer fields exact, fp64 within tolerance.
"""
import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import configs, scene as SC

from util import build_case, compare_generation

# (config, kwargs, recursion limit, capture plane centre, direction, (length, width))
CAPTURE_CASES = {
    # behind the doublet, square smaller than the beam: only part of the transmitted rays, and
    # the weak reflected generations (thresholds 1e-3) that come back through it
    "achromat": ("config2", dict(n=6000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6,
                 (0.0, 25.0, 0.0), (0.0, 1.0, 0.0), (12.0, 9.0)),
    # Michelson: a plane across the return arm sees several generations from both directions
    "michelson": ("config5", dict(n=2000, gausslets=False), None,
                  (0.0, 12.0, 0.0), (0.0, 1.0, 0.0), (6.0, 6.0)),
    "michelson_gausslets": ("config5", dict(n=1500, gausslets=True), None,
                            (12.0, 0.0, 0.0), (1.0, 0.0, 0.0), (6.0, 5.0)),
}


def capture_face_list(core, centre, direction, size):
    """BaseCapturePlane._face_list_default (probes.py:97-101): one RectangularFace in a FaceList
    whose owner carries the plane's pose."""
    face = core.cfaces.RectangularFace(length=size[0], width=size[1], offset=0.0, z_plane=0.0)
    fl = core.ctracer.FaceList(owner=configs.Pose(centre=centre, direction=direction))
    fl.faces = [face]
    fl.sync_transforms()
    return fl


def staggered_wavelengths(wavelengths, n_gens):
    """Different (overlapping) wavelength tables per collection so the np.unique merge and the
    running wl_offset are really exercised (one trace alone would repeat the same table)."""
    wl = np.asarray(wavelengths, dtype=np.double)
    return [wl if g % 2 == 0 else np.concatenate([wl[::-1], [wl.max() + 0.1 * (g + 1)]]) for g in range(n_gens)]


def oracle_trace(core, case):
    from oracle import oracle as O
    name, kw, rl, centre, direction, size = CAPTURE_CASES[case]
    cfg = build_case(core, name, kw, rl)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    gens, _ = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    return cfg, gens, capture_face_list(core, centre, direction, size)


@pytest.mark.parametrize("case", sorted(CAPTURE_CASES))
def test_oracle_capture_bit_exact_with_reference(refcore, case):
    from oracle import oracle as O
    cfg, gens, fl = oracle_trace(refcore, case)
    wls = staggered_wavelengths(cfg['wavelengths'], len(gens))
    # for odd collections the staggered table is longer and reversed: re-point the indices so
    # they stay valid (index i of the reversed table = index n-1-i of the original)
    cols = []
    for g, (arr, wl) in enumerate(zip(gens, wls)):
        arr = arr.copy()
        if g % 2 == 1:
            b = arr['base_ray'] if arr.dtype == A.gausslet_dtype else arr
            b['wavelength_idx'] = len(cfg['wavelengths']) - 1 - b['wavelength_idx']
        cols.append(arr)
    ref_cols = [O.reference_collection(refcore, a, w) for a, w in zip(cols, wls)]
    want, want_wl, _ = O.reference_select_intersections(refcore, fl, ref_cols)
    got, got_wl, counts = O.select_intersections(SC.Scene([fl], np.asarray([1.0])), cols, wls,
                                                 face_ids=[f.idx for f in fl.faces])
    assert len(want) > 0 and sum(1 for c in counts if c) >= 2, "capture plane should see several generations"
    assert 0 < len(want) < sum(len(a) for a in cols)
    assert np.array_equal(got_wl, want_wl)
    assert got.tobytes() == want.tobytes(), "%s: captured rays not bit-identical to the reference" % case


def test_capture_leaves_inputs_untouched_and_cuts_length(core):
    from oracle import oracle as O
    cfg, gens, fl = oracle_trace(core, "michelson")
    before = [g.copy() for g in gens]
    got, wl, counts = O.select_intersections(SC.Scene([fl], np.asarray([1.0])), gens,
                                             [cfg['wavelengths']] * len(gens))
    assert all(a.tobytes() == b.tobytes() for a, b in zip(gens, before))
    assert sum(counts) == len(got) and np.array_equal(wl, np.unique(cfg['wavelengths']))
    # every captured ray ends ON the capture plane (y = 12 in this case) and is shorter than traced
    end = got['origin'] + got['direction'] * got['length'][:, None]
    assert np.allclose(end[:, 1], 12.0, atol=1e-9)
    assert np.all(got['end_face_idx'] == 0)


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CAPTURE_CASES))
def test_cuda_capture_resident_generations_match_oracle(core, engine, case):
    """Trace on the device, filter the still-resident generations (rpx_capture): nothing but the
    captured rays crosses PCIe."""
    from oracle import oracle as O
    cfg, gens, fl = oracle_trace(core, case)
    wls = [cfg['wavelengths']] * len(gens)
    want, want_wl, want_counts = O.select_intersections(SC.Scene([fl], np.asarray([1.0])), gens, wls)
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    engine.set_capture_scene(SC.Scene([fl], np.asarray([1.0])))
    rays = np.ascontiguousarray(cfg['rays'])
    res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'])
    try:
        assert res.counts == [len(g) for g in gens]
        got, got_wl, counts = res.capture(cfg['wavelengths'])
    finally:
        res.free()
    assert counts == want_counts
    assert np.array_equal(got_wl, want_wl)
    assert len(got) == len(want) > 0
    compare_generation(got, want, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["achromat", "michelson_gausslets"])
def test_cuda_select_intersections_mirror(core, engine, case):
    """The drop-in functions: host collections in, one merged collection out, wavelength tables
    merged like np.unique does in the reference."""
    from oracle import oracle as O
    from raypier_optics_b200.core import ctracer as CT
    cfg, gens, fl = oracle_trace(core, case)
    wls = staggered_wavelengths(cfg['wavelengths'], len(gens))
    cols = []
    for g, arr in enumerate(gens):
        arr = arr.copy()
        if g % 2 == 1:
            b = arr['base_ray'] if arr.dtype == A.gausslet_dtype else arr
            b['wavelength_idx'] = len(cfg['wavelengths']) - 1 - b['wavelength_idx']
        cols.append(arr)
    for f in fl.faces:
        f.idx = 7  # whatever the caller left in Face.idx ends up in end_face_idx
    want, want_wl, _ = O.select_intersections(SC.Scene([fl], np.asarray([1.0])), cols, wls, face_ids=[7])
    is_g = cols[0].dtype == A.gausslet_dtype
    cls = CT.GaussletCollection if is_g else CT.RayCollection
    host_cols = []
    for a, w in zip(cols, wls):
        c = cls.from_array(a)
        c.wavelengths = w
        host_cols.append(c)
    fn = CT.select_gausslet_intersections if is_g else CT.select_ray_intersections
    out = fn(fl, host_cols)
    assert isinstance(out, cls)
    assert np.array_equal(np.asarray(out.wavelengths), want_wl)
    got = out.copy_as_array()
    assert len(got) == len(want) > 0
    compare_generation(got, want, case)
    base = got['base_ray'] if is_g else got
    assert np.all(base['end_face_idx'] == 7)


@pytest.mark.gpu
def test_cuda_capture_handles_empty_and_missing_plane(core, engine):
    cfg, gens, fl = oracle_trace(core, "michelson")
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    # a plane far away from everything captures nothing
    far = capture_face_list(core, (500.0, 500.0, 500.0), (0.0, 0.0, 1.0), (1.0, 1.0))
    engine.set_capture_scene(SC.Scene([far], np.asarray([1.0])))
    res = engine.trace(np.ascontiguousarray(cfg['rays']), cfg['max_length'], cfg['recursion_limit'])
    try:
        got, wl, counts = res.capture(cfg['wavelengths'])
    finally:
        res.free()
    assert len(got) == 0 and sum(counts) == 0
