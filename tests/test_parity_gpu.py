"""GPU parity: the CUDA path (through the C ABI) against the plain-C oracle on the same
seeded inputs.  Integer fields bit-exact; fp64 fields within the north-star tolerances."""
import numpy as np
import pytest

from raypier_optics_b200 import scene as SC

from util import PARITY_CASES, UNPINNED_CASES, build_case, compare_traces

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
    worst = compare_traces(got, want, name)
    assert np.array_equal(res.face_counts, want_counts)
    assert res.launches >= len(want) + 1  # k_intersect for generation 0 + one k_shade per generation
    print("%s: %d generations, %d segments, worst rel err %.2e" % (name, len(got), res.segments, worst))
