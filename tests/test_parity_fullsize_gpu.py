"""GPU parity at BASELINE sizes: the CUDA path against the plain-C oracle on >= 1e6-ray sources.

Only at this size do the grouped look-back's multi-window path (> 1024 tiles), the exact-sync
generation loop (traces too large for the pipelined loop's 4 GB bound, rpx_api.cu trace_loop) and
knife-edge comparisons under the non-IEEE rcp / rsqrt / fdiv primitives see enough rays to matter.

The oracle runs the same trace as independent contiguous blocks on the host threads
(oracle.trace_rays_blocks); the slice of CUDA generation g that descends from a block is located
with the CUDA parent chain (root ancestor index, non-decreasing because emission is parent-ordered)
and compared field by field: integer fields bit-exact, fp64 within the north-star tolerances
(tests/util.py).  Reference semantics: core/tracer.py:39-45, ctracer.pyx:2062-2118, 2214-2281.
"""
import time

import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import scene as SC

from util import build_case, compare_generation

pytestmark = pytest.mark.gpu


def _base(arr):
    return arr['base_ray'] if arr.dtype == A.gausslet_dtype else arr


def _roots(gens):
    """root[g][i] = index of the source ray that ray i of generation g descends from."""
    roots = [np.arange(len(gens[0]), dtype=np.int64)]
    for g in range(1, len(gens)):
        p = _base(gens[g])['parent_idx'].astype(np.int64)
        assert np.all(np.diff(p) >= 0), "generation %d is not in parent order" % g
        roots.append(roots[g - 1][p])
    return roots


def check_against_blocked_oracle(engine, cfg, label, block=32768, flags=0, got=None):
    from oracle import oracle as O
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    rays = np.ascontiguousarray(cfg['rays'])
    if got is None:
        engine.set_scene(sc)
        res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'], flags)
        got = res.generations()
        got_fc = res.face_counts.copy()
        res.free()
    else:
        got, got_fc = got
    roots = _roots(got)
    t0 = time.time()
    fc = np.zeros_like(got_fc, dtype=np.int64)
    worst, segs, n_blocks = 0.0, 0, 0
    for lo, hi, want, wfc in O.trace_rays_blocks(sc, rays, cfg['recursion_limit'], cfg['max_length'], block=block):
        fc += wfc
        n_blocks += 1
        start = [int(np.searchsorted(r, lo, side='left')) for r in roots]
        stop = [int(np.searchsorted(r, hi, side='left')) for r in roots]
        for g in range(len(got)):
            w = want[g] if g < len(want) else rays[:0]
            assert stop[g] - start[g] == len(w), "%s block %d..%d gen %d: %d rays vs oracle %d" % (
                label, lo, hi, g, stop[g] - start[g], len(w))
            if len(w) == 0:
                continue
            if g > 0:
                _base(w)['parent_idx'] += np.uint32(start[g - 1])
            worst = max(worst, compare_generation(got[g][start[g]:stop[g]], w, "%s block %d gen %d" % (label, lo, g)))
            segs += len(w)
        assert len(want) <= len(got)
    assert segs == sum(len(g) for g in got)
    assert np.array_equal(fc, got_fc.astype(np.int64)), "%s: Face.count differs" % label
    print("%s: %d generations %s, %d segments in %d oracle blocks (%.1f s), worst rel err %.2e"
          % (label, len(got), [len(g) for g in got], segs, n_blocks, time.time() - t0, worst))
    return worst


FULL_CASES = [
    # BASELINE configs[1] at its quoted size
    ("config2", dict(n=1000000), None),
    # configs[2] (Newton + Zernike secant), 1e6 of its 1e7
    ("config3", dict(n=1000000), None),
    # configs[3], TIR prism chain: branches to ~4e6 rays per generation
    ("config4_prisms", dict(n=1000000), 12),
    # configs[4], Michelson gausslets: branches to 8e5 gausslets (668-byte records)
    ("config5", dict(n=200000, gausslets=True), None),
    # configs[4] as plain rays at 1e6: 1.8e7 segments
    ("config5", dict(n=1000000, gausslets=False), None),
]


@pytest.mark.parametrize("name,kw,rl", FULL_CASES, ids=["%s-%d" % (c[0], c[1]['n']) for c in FULL_CASES])
def test_cuda_matches_oracle_at_baseline_size(engine, core, name, kw, rl):
    cfg = build_case(core, name, kw, rl)
    check_against_blocked_oracle(engine, cfg, "%s@%d" % (name, kw['n']))


def test_exact_sync_loop_beyond_the_pipelined_bound(engine, core):
    """5.5e6 achromat rays x 188 B x kids^2 (= 4) > 4e9: rpx_trace leaves the pipelined loop and
    runs the host-synchronised generation loop (rpx_api.cu trace_loop); same answers."""
    cfg = build_case(core, "config2", dict(n=5500000), None)
    assert len(cfg['rays']) * 188 * 4 > 4.0e9
    check_against_blocked_oracle(engine, cfg, "config2@5.5e6 (exact-sync loop)", block=65536)


def test_streamed_trace_at_full_size_matches_oracle(engine, core):
    """rpx_trace_streamed (chunked source, overlapped copies, parent renumbering across chunks)
    at 1e6 rays with branching, against the oracle."""
    cfg = build_case(core, "config2", dict(n=1000000, reflection_threshold=0.02, transmission_threshold=0.02), 5)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    engine.set_scene(sc)
    rays = np.ascontiguousarray(cfg['rays'])
    res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'])
    counts = list(res.counts)
    res.free()
    out = [engine.pinned_empty(c + 64, rays.dtype) for c in counts] + [engine.pinned_empty(64, rays.dtype)]
    gens, fc, _ = engine.trace_streamed(rays, cfg['max_length'], cfg['recursion_limit'], out, chunk_rays=131072)
    assert [len(g) for g in gens] == counts
    check_against_blocked_oracle(engine, cfg, "config2 branching, streamed", got=([g for g in gens], fc))
