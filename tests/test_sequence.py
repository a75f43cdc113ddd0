"""Sequential mode: trace_ray_sequence (raypier/core/tracer.py:50-99) over
trace_one_face_segment_c / trace_one_face_gausslet_c (ctracer.pyx:2121-2170, 2284-2347)."""
import numpy as np
import pytest

from raypier_optics_b200 import configs, scene as SC
from raypier_optics_b200.core import tracer as T

from util import compare_traces


def achromat_sequence(core, n, gausslets=False):
    """The doublet traced front to back: one FaceList listed three times (so all_faces chains its
    faces three times and f.idx keeps the LAST occurrence, as in the reference)."""
    cfg = configs.build(core, "config2", n=n)
    fl = cfg['face_lists'][0]
    rays = cfg['rays']
    if gausslets:
        rays = rays.copy()
        rays['length'] = cfg['max_length']
        rays['ray_type_id'] = 2
        gc = core.ctracer.GaussletCollection.from_rays(rays)
        gc.config_parabasal_rays(cfg['wavelengths'], 0.3, 0.0)
        rays = gc.copy_as_array()
    return cfg, [(fl, 0), (fl, 1), (fl, 2)], rays


def mirror_sequence(core, n):
    """Michelson arm: cube entrance face, diagonal splitter, cube exit, mirror, back."""
    cfg = configs.build(core, "config5", n=n, gausslets=False)
    cube, m1, m2 = cfg['face_lists']
    return cfg, [(cube, 0), (cube, 6), (cube, 1), (m1, 0), (cube, 1)], cfg['rays']


@pytest.mark.parametrize("which", ["achromat", "achromat_gausslets", "michelson"])
def test_oracle_sequence_bit_exact_with_reference(refcore, which):
    from oracle import oracle as O
    if which == "michelson":
        cfg, seq, rays = mirror_sequence(refcore, 1500)
    else:
        cfg, seq, rays = achromat_sequence(refcore, 1500, gausslets=which.endswith("gausslets"))
    rc = O.reference_collection(refcore, rays, cfg['wavelengths'])
    traced, all_faces = O.reference_trace_ray_sequence(refcore, rc, seq, cfg['recursion_limit'], cfg['max_length'])
    ref = [t.copy_as_array() for t in traced]
    sc = SC.Scene([fl for fl, _ in seq], cfg['wavelengths'])
    gidx = T.sequence_face_indices(seq)
    gens, counts = O.trace_ray_sequence(sc, rays, gidx, cfg['recursion_limit'], cfg['max_length'])
    assert [len(g) for g in gens] == [len(r) for r in ref]
    assert len(gens) >= 3
    for gi, (g, r) in enumerate(zip(gens, ref)):
        assert g.tobytes() == r.tobytes(), "%s generation %d not bit-identical" % (which, gi)
    # a FaceList listed k times appears k times in all_faces (same objects): the hit counts
    # land on the index of the LAST occurrence, which is the one Face.idx keeps
    last = {id(f): i for i, f in enumerate(all_faces)}
    expect = [f.count if last[id(f)] == i else 0 for i, f in enumerate(all_faces)]
    assert counts.tolist() == expect


def test_sequence_indices_follow_last_occurrence(core):
    cfg, seq, _ = achromat_sequence(core, 10)
    assert T.sequence_face_indices(seq) == [6, 7, 8]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["achromat", "achromat_gausslets", "michelson"])
def test_cuda_sequence_matches_oracle(engine, core, which):
    from oracle import oracle as O
    if which == "michelson":
        cfg, seq, rays = mirror_sequence(core, 20000)
    else:
        cfg, seq, rays = achromat_sequence(core, 20000, gausslets=which.endswith("gausslets"))
    sc = SC.Scene([fl for fl, _ in seq], cfg['wavelengths'])
    gidx = T.sequence_face_indices(seq)
    want, want_counts = O.trace_ray_sequence(sc, rays, gidx, cfg['recursion_limit'], cfg['max_length'])
    engine.set_scene(sc)
    res = engine.trace_sequence(rays, gidx, cfg['max_length'], cfg['recursion_limit'])
    got = res.generations()
    compare_traces(got, want, which)
    assert np.array_equal(res.face_counts, want_counts)
    res.free()


@pytest.mark.gpu
def test_drop_in_trace_functions(engine, core):
    """The reference-facing functions with the reference's container classes."""
    from oracle import oracle as O
    ct = core.ctracer
    cfg = configs.build(core, "config1", n=3000)
    rc = ct.RayCollection.from_array(cfg['rays'])
    rc.wavelengths = cfg['wavelengths']
    traced, all_faces = T.trace_rays(rc, cfg['face_lists'], recursion_limit=cfg['recursion_limit'],
                                     max_length=cfg['max_length'])
    assert traced[0] is rc and [len(t) for t in traced] == [3000, 3000, 3000]
    assert traced[1].parent is traced[0] and np.array_equal(traced[2].wavelengths, cfg['wavelengths'])
    assert [f.idx for f in all_faces] == [0, 1] and [f.count for f in all_faces] == [3000, 3000]
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want, _ = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    compare_traces([t.copy_as_array() for t in traced], want, "trace_rays drop-in")
    # sequential drop-in
    cfg2, seq, rays = achromat_sequence(core, 2000)
    rc2 = ct.RayCollection.from_array(rays)
    rc2.wavelengths = cfg2['wavelengths']
    traced2, faces2 = T.trace_ray_sequence(rc2, seq, recursion_limit=100, max_length=cfg2['max_length'])
    assert len(traced2) == 4 and len(faces2) == 9
    assert np.all(np.isinf(traced2[-1].length))            # the last generation is returned untraced


@pytest.mark.gpu
@pytest.mark.parametrize("gauss", [False, True])
def test_drop_in_retraces_the_same_collection(engine, core, gauss):
    """A model re-traces ONE source collection over and over (raypier/tracer.py calls trace_rays on every trait
    change): trace_rays mutates it in place -- after the first call it holds generation 0 as the device left it
    (lengths / end_face_idx written back), in a pooled page-locked block that the next call lends straight to the
    upload.  The second and third trace must reproduce the first byte for byte, results large enough for the
    pool (> 4 MB per generation) included, and arrays handed out earlier must stay intact while the pool recycles."""
    from raypier_optics_b200._hostpool import get_pool
    ct = core.ctracer
    cfg = configs.build(core, "config5", n=30000, gausslets=gauss)
    cls = ct.GaussletCollection if gauss else ct.RayCollection
    rc = cls.from_array(cfg['rays'])
    rc.wavelengths = cfg['wavelengths']
    kept = []
    for rep in range(3):
        traced, _ = T.trace_rays(rc, cfg['face_lists'], recursion_limit=cfg['recursion_limit'], max_length=cfg['max_length'])
        assert traced[0] is rc
        arrays = [t.copy_as_array() for t in traced]
        if rep == 0:
            first = arrays
            kept = traced                      # holds the pooled blocks of the first trace
            snapshot = [a.copy() for a in arrays]
        else:
            assert len(arrays) == len(first)
            for g, (a, b) in enumerate(zip(arrays, first)):
                assert a.tobytes() == b.tobytes(), "generation %d differs on re-trace %d" % (g, rep)
    # the first trace's generations (still referenced) were not overwritten by the later ones
    for t, snap in zip(kept[1:], snapshot[1:]):
        assert t.copy_as_array().tobytes() == snap.tobytes()
    st = get_pool(None).stats()
    assert st["hits"] + st["misses"] + st["fallbacks"] > 0   # the generations did go through the result pool


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("config2", dict(n=20000)), ("config5", dict(n=3000, gausslets=True)),
                                     ("config4_prisms", dict(n=4000))])
def test_drop_in_trace_rays_with_genuine_reference_objects(engine, name, kw):
    """trace_rays fed with GENUINE raypier.core objects (RayCollection / GaussletCollection, FaceList,
    faces and materials built from oracle/_ref, which travels to the GPU box): the scene is flattened
    from the reference's own attributes, the result comes back in the reference's own container
    classes, traced_rays[0] is input_rays (mutated in place) and every generation matches the
    reference's own Cython trace of the same objects (core/tracer.py:9-47)."""
    from oracle import oracle as O
    refcore = O.import_reference("parity")
    if refcore is None:
        pytest.skip("oracle/_ref is not present on this machine")
    ct = refcore.ctracer
    gauss = bool(kw.get("gausslets"))
    cls = ct.GaussletCollection if gauss else ct.RayCollection
    rl = 12 if name == "config4_prisms" else None

    def build():
        cfg = configs.build(refcore, name, **kw)
        if rl:
            cfg['recursion_limit'] = rl
        rc = O.reference_collection(refcore, cfg['rays'], cfg['wavelengths'])
        assert type(rc) is cls
        return cfg, rc

    cfg, rc = build()
    want, want_faces = O.reference_trace_rays(refcore, rc, cfg['face_lists'], cfg['recursion_limit'], cfg['max_length'])
    want = [w.copy_as_array() for w in want]
    want_counts = [f.count for f in want_faces]
    cfg, rc = build()   # fresh objects: the reference trace mutated the first set
    traced, all_faces = T.trace_rays(rc, cfg['face_lists'], recursion_limit=cfg['recursion_limit'],
                                     max_length=cfg['max_length'])
    assert traced[0] is rc
    assert all(type(t) is cls for t in traced)
    assert [len(t) for t in traced] == [len(w) for w in want]
    assert all(traced[g].parent is traced[g - 1] for g in range(1, len(traced)))
    assert [f.idx for f in all_faces] == list(range(len(all_faces)))
    assert [f.count for f in all_faces] == want_counts
    from raypier_optics_b200 import _abi as A
    native = A.gausslet_dtype if gauss else A.ray_dtype
    got = [np.ascontiguousarray(t.copy_as_array()).view(native) for t in traced]
    worst = compare_traces(got, [np.ascontiguousarray(w).view(native) for w in want], name + " (genuine reference objects)")
    print("%s with genuine raypier.core objects: %s, worst rel err %.2e" % (name, [len(t) for t in traced], worst))


@pytest.mark.gpu
@pytest.mark.parametrize("genuine", [False, True])
def test_drop_in_trace_rays_with_a_decomposition_plane(engine, core, genuine):
    """ResampleGaussletMaterial (cmaterials.pyx:1766-1831): the face absorbs on the device, and between
    generations the host hands the captured gausslets to the material's Python callback and appends the
    result (ctracer.pyx:2274-2280).  Against the oracle's decomposition loop, which is pinned bit-exact to
    the reference (test_oracle_decomposition_loop_bit_exact_with_reference); with the host mirrors and with
    genuine raypier.core objects."""
    from oracle import oracle as O
    from test_oracle_vs_reference import _relaunch_for
    lib = core
    if genuine:
        lib = O.import_reference("parity")
        if lib is None:
            pytest.skip("oracle/_ref is not present on this machine")
    cfg = configs.build(lib, "resample", n=3000)
    mat = cfg['decomp_material']
    mat.eval_func = _relaunch_for(lib, cfg['max_length'])
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want, want_counts = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'],
                                     decomp={1: lambda a: configs.resample_relaunch(a, cfg['max_length'])})
    rc = O.reference_collection(lib, cfg['rays'], cfg['wavelengths']) if genuine else \
        lib.ctracer.GaussletCollection.from_array(cfg['rays'])
    rc.wavelengths = cfg['wavelengths']
    traced, all_faces = T.trace_rays(rc, cfg['face_lists'], recursion_limit=cfg['recursion_limit'],
                                     max_length=cfg['max_length'])
    assert traced[0] is rc and [len(t) for t in traced] == [len(w) for w in want] and len(traced) == 8
    assert [f.count for f in all_faces] == want_counts.tolist() and all_faces[1].count == 0
    assert mat.capture_count == sum(int((w['base_ray']['end_face_idx'] == 1).sum()) for w in want)
    from raypier_optics_b200 import _abi as A
    got = [np.ascontiguousarray(t.copy_as_array()).view(A.gausslet_dtype) for t in traced]
    worst = compare_traces(got, want, "decomposition plane")
    print("decomposition plane (%s objects): %s, worst rel err %.2e"
          % ("genuine" if genuine else "mirror", [len(t) for t in traced], worst))
