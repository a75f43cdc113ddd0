"""GPU parity: the CUDA path (through the C ABI) against the plain-C oracle on the same
seeded inputs.  Integer fields bit-exact; fp64 fields within the north-star tolerances."""
import numpy as np
import pytest

from raypier_optics_b200 import scene as SC

from util import (PARITY_CASES, UNPINNED_CASES, build_case, check_noisy_faces_on_surface, compare_traces,
                  noisy_face_scales)

pytestmark = pytest.mark.gpu


ALL_CASES = PARITY_CASES + UNPINNED_CASES


@pytest.mark.parametrize("name,kw,rl", ALL_CASES, ids=[c[0] + "-" + str(i) for i, c in enumerate(ALL_CASES)])
def test_cuda_matches_oracle(engine, core, name, kw, rl):
    from oracle import oracle as O
    cfg = build_case(core, name, kw, rl)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want, want_counts = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    engine.set_scene(sc)
    res = engine.trace(cfg['rays'], cfg['max_length'], cfg['recursion_limit'])
    got = res.generations()
    assert res.counts == [len(g) for g in want]
    keep = None
    if name == "zoo":  # two face types whose reference formulae are rounding-noise limited: see util.py
        keep = noisy_face_scales(sc, want)
        n_checked = check_noisy_faces_on_surface(sc, got, want, name)
        assert n_checked > 100
        print("zoo: %d rays left out of the fp64 comparison, %d hit points checked against the analytic surface"
              % (sum(int(np.isinf(k).sum()) for k in keep), n_checked))
    worst = compare_traces(got, want, name, keep=keep)
    assert np.array_equal(res.face_counts, want_counts)
    assert res.launches >= len(want) + 1  # k_intersect for generation 0 + one k_shade per generation
    print("%s: %d generations, %d segments, worst rel err %.2e" % (name, len(got), res.segments, worst))


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,rl,chunk", [
    ("config2", dict(n=20000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6, 4096),
    ("config5", dict(n=3000, gausslets=True), None, 1000),
    ("config4_prisms", dict(n=5000), 12, 777),
])
def test_streamed_trace_is_identical_to_one_shot(core, engine, name, kw, rl, chunk):
    """rpx_trace_streamed (chunked source, overlapped copies) returns byte-for-byte what one
    rpx_trace call returns: same generations in the same order, parent_idx renumbered globally,
    same Face.count."""
    from util import build_case
    from raypier_optics_b200 import scene as SC
    cfg = build_case(core, name, kw, rl)
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    rays = np.ascontiguousarray(cfg['rays'])
    res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'])
    want = res.generations()
    want_fc = res.face_counts.copy()
    res.free()
    out = [engine.pinned_empty(len(g) + 64, rays.dtype) for g in want] + [engine.pinned_empty(64, rays.dtype)]
    gens, fc, ms = engine.trace_streamed(rays, cfg['max_length'], cfg['recursion_limit'], out, chunk_rays=chunk)
    assert [len(g) for g in gens] == [len(g) for g in want]
    assert len(rays) > 2 * chunk  # really chunked
    for g, (a, b) in enumerate(zip(gens, want)):
        assert a.tobytes() == b.tobytes(), "generation %d differs from the one-shot trace" % g
    assert fc.tolist() == want_fc.tolist() and ms > 0
    # too few / too small output buffers are reported, not overrun
    from raypier_optics_b200._lib import RpxError
    with pytest.raises(RpxError):
        engine.trace_streamed(rays, cfg['max_length'], cfg['recursion_limit'], out[:1], chunk_rays=chunk)
    small = [engine.pinned_empty(8, rays.dtype) for _ in out]
    with pytest.raises(RpxError):
        engine.trace_streamed(rays, cfg['max_length'], cfg['recursion_limit'], small, chunk_rays=chunk)


@pytest.mark.parametrize("name,kw,rl,chunk", [
    ("config2", dict(n=20000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6, 4096),
    ("config4_prisms", dict(n=5000), 12, 777),
])
def test_streamed_trace_in_place_returns_the_same_generation_0(core, engine, name, kw, rl, chunk):
    """rpx_trace_streamed with out[0] aliasing the source (the reference's in-place convention,
    traced_rays[0] is input_rays): every generation byte-identical to the one-shot trace."""
    cfg = build_case(core, name, kw, rl)
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    rays = np.ascontiguousarray(cfg['rays'])
    res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'])
    want = res.generations()
    res.free()
    src = engine.pinned_empty(len(rays), rays.dtype)
    src[:] = rays
    out = [src] + [engine.pinned_empty(len(g) + 64, rays.dtype) for g in want[1:]] + [engine.pinned_empty(64, rays.dtype)]
    gens, fc, ms = engine.trace_streamed(src, cfg['max_length'], cfg['recursion_limit'], out, chunk_rays=chunk)
    assert [len(g) for g in gens] == [len(g) for g in want]
    assert gens[0].ctypes.data == src.ctypes.data
    for g, (a, b) in enumerate(zip(gens, want)):
        assert a.tobytes() == b.tobytes(), "generation %d differs from the one-shot trace" % g
