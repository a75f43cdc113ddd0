"""Full-size GPU checks through size-independent properties (the oracle would take minutes at
these sizes): determinism, the sharding identity trace(A ++ B) == trace(A) ++ trace(B) (what
multi-GPU tracing relies on), ordering invariants of the emission, energy bookkeeping, and the
streaming mode."""
import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import configs, distributed as rd, scene as SC

pytestmark = pytest.mark.gpu


def _trace(engine, cfg, rays=None, flags=0):
    res = engine.trace(cfg['rays'] if rays is None else rays, cfg['max_length'], cfg['recursion_limit'], flags)
    return res


@pytest.fixture(scope="module")
def achromat(core):
    cfg = configs.build(core, "config2", n=1000000, reflection_threshold=0.02, transmission_threshold=0.02)
    cfg['recursion_limit'] = 5
    return cfg


def test_full_size_determinism_and_invariants(engine, achromat):
    cfg = achromat
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    r1 = _trace(engine, cfg)
    g1 = r1.generations()
    r2 = _trace(engine, cfg)
    g2 = r2.generations()
    assert r1.counts == r2.counts and r1.counts[0] == 1000000 and len(r1.counts) == 5
    for a, b in zip(g1, g2):
        assert a.tobytes() == b.tobytes()          # bit-reproducible run to run
    for g in range(1, len(g1)):
        child, parent = g1[g], g1[g - 1]
        p = child['parent_idx'].astype(np.int64)
        assert np.all(np.diff(p) >= 0)                                   # parents in order
        same = p[1:] == p[:-1]
        refl = (child['ray_type_id'] & A.REFL_RAY).astype(bool)
        assert np.all(refl[:-1][same] & ~refl[1:][same])                 # reflected before transmitted
        assert np.all(parent['end_face_idx'][p] != A.NO_FACE)            # only terminated parents have children
        hit = parent['end_face_idx'] != A.NO_FACE
        assert np.all(np.bincount(p, minlength=len(parent))[~hit] == 0)
        assert np.array_equal(child['wavelength_idx'], parent['wavelength_idx'][p])
        assert np.array_equal(child['ray_ident'], parent['ray_ident'][p])
        end = parent['origin'][p] + parent['direction'][p] * parent['length'][p][:, None]
        assert np.allclose(child['origin'], end, rtol=0, atol=1e-9)      # children start at the hit point
        # power bookkeeping of the Fresnel split: children never carry more power than the parent
        def power(x):
            return (np.abs(x['E1_amp']) ** 2 + np.abs(x['E2_amp']) ** 2) * x['refractive_index'].real
        csum = np.bincount(p, weights=power(child), minlength=len(parent))
        assert np.all(csum[hit] <= power(parent)[hit] * (1 + 1e-9))
    assert int(r1.face_counts.sum()) == sum(int((g['end_face_idx'] != A.NO_FACE).sum()) for g in g1)
    r1.free()
    r2.free()


def test_sharding_identity_at_full_size(engine, achromat):
    """trace(A ++ B) == trace(A) ++ trace(B) after parent renumbering: the property the
    multi-GPU path (distributed.trace_sharded) is built on."""
    cfg = achromat
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    whole = _trace(engine, cfg)
    gw = whole.generations()
    lo, hi = rd.shard_bounds(len(cfg['rays']), 2, 0), rd.shard_bounds(len(cfg['rays']), 2, 1)
    parts, counts = [], []
    for a, b in (lo, hi):
        r = _trace(engine, cfg, np.ascontiguousarray(cfg['rays'][a:b]))
        parts.append(r.generations())
        counts.append(r.counts + [0] * (len(gw) - len(r.counts)))
        r.free()
    counts_all = np.array(counts)
    for g in range(len(gw)):
        pieces = []
        for rank in range(2):
            arr = parts[rank][g] if g < len(parts[rank]) else gw[g][:0]
            pieces.append(rd.globalize_generation(arr, g, rd.parent_offsets(counts_all, rank)))
        assert np.concatenate(pieces).tobytes() == gw[g].tobytes()
    whole.free()


def test_streaming_mode_keeps_counts(engine, core):
    cfg = configs.build(core, "config5", n=200000, gausslets=True)
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    full = _trace(engine, cfg)
    lean = _trace(engine, cfg, flags=A.TRACE_KEEP_LAST_ONLY)
    assert full.counts == lean.counts == [200000, 200000, 400000, 400000, 400000, 400000, 800000, 800000]
    assert np.array_equal(full.face_counts, lean.face_counts)
    from raypier_optics_b200._lib import RpxError
    with pytest.raises(RpxError):
        lean.generation(0)                      # dropped generations are reported, not faked
    last = full.generation(7)
    assert last.dtype == A.gausslet_dtype and len(last) == 800000
    assert np.all(np.isfinite(last['para_rays']['direction']))
    full.free()
    lean.free()


def test_empty_and_ragged_inputs(engine, core):
    cfg = configs.build(core, "config1", n=1000)
    engine.set_scene(SC.Scene(cfg['face_lists'], cfg['wavelengths']))
    empty = _trace(engine, cfg, cfg['rays'][:0])
    assert empty.counts == [] and empty.n_generations == 0
    for n in (1, 31, 33, 127, 129, 1000):       # tile-boundary sizes
        r = _trace(engine, cfg, np.ascontiguousarray(cfg['rays'][:n]))
        assert r.counts == [n, n, n]
        r.free()
    miss = cfg['rays'][:200].copy()
    miss['origin'][:, 0] += 500.0               # every ray misses the lens
    r = _trace(engine, cfg, miss)
    assert r.counts == [200] and np.all(r.generation(0)['end_face_idx'] == A.NO_FACE)
    assert np.all(r.generation(0)['length'] == np.float32(cfg['max_length']))
    r.free()
    limited = engine.trace(cfg['rays'], cfg['max_length'], 2)   # recursion limit cuts the trace
    assert limited.counts == [1000, 1000]
    limited.free()
