"""Host-side logic that needs no GPU: Zernike tape construction, scene validation errors,
collection containers, sharding arithmetic."""
import os
import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import configs, distributed, scene as SC


def test_zernike_tape_matches_recursive_oracle(core):
    """The host-built straight-line tape must evaluate to exactly what the reference's
    memoised recursion gives (the oracle runs the recursion itself)."""
    from oracle import oracle as O
    D, F, S = core.cdistortions, core.cfaces, core.cshapes
    coefs = {"j%d" % j: 0.001 * ((-1) ** j) * (j + 1) for j in (1, 2, 3, 4, 5, 7, 8, 11, 12, 13, 17, 22, 24, 28)}
    dist = D.ZernikeDistortion(unit_radius=7.5, **coefs)
    shape = S.CircleShape(radius=10.0)
    face = F.DistortionFace(base_face=F.ShapedPlanarFace(shape=shape), distortion=dist, shape=shape)
    fl = core.ctracer.FaceList()
    fl.faces = [face]
    sc = SC.Scene([fl], np.array([1.0]))
    d = sc.distortions[0]
    K = A.ZERNIKE_MAX_K

    def run_tape(off, length, r):
        ws = np.full((3, K), np.nan)

        def op(o):
            if o < 2:
                return float(o)
            k, w = divmod(o - 2, 3)
            return ws[w, k]
        for t in sc.ztape[off:off + length]:
            a, b, c = op(t['a']), op(t['b']), op(t['c'])
            if t['kind'] == 0:
                ws[0, t['dst']] = r * (a + b) - c
            elif t['kind'] == 1:
                v = a + b
                v += r * (op(t['d']) + op(t['e']))
                ws[1, t['dst']] = v - c
            elif t['kind'] == 2:
                ws[2, t['dst']] = (a + b) - c
            else:
                ws[0, t['dst']] = 1.0
        return op

    rng = np.random.default_rng(0)
    for _ in range(20):
        x, y = rng.uniform(-7, 7, 2)
        xn, yn = x / 7.5, y / 7.5
        r, theta = np.sqrt(xn * xn + yn * yn), np.arctan2(yn, xn)
        opz = run_tape(d['tape_z_off'], d['tape_z_len'], r)
        opg = run_tape(d['tape_g_off'], d['tape_g_len'], r)
        Z, Zg = 0.0, np.zeros(3)
        for c in sc.zcoefs[d['coef_off']:d['coef_off'] + d['n_coefs']]:
            n, m = int(c['n']), int(c['m'])
            N = (np.sqrt(n + 1) if m == 0 else np.sqrt(2 * (n + 1))) * c['value']
            PH = np.cos(m * theta) if m >= 0 else -np.sin(m * theta)
            PHp = -m * np.sin(m * theta) if m >= 0 else -m * np.cos(m * theta)
            Z += N * opz(c['opR_z']) * PH
            R, Rp, Rr = opg(c['opR']), opg(c['opRp']), opg(c['opRr'])
            Zg[2] += N * R * PH
            Zg[0] += N * (Rp * np.cos(theta) * PH + Rr * (-np.sin(theta)) * PHp)
            Zg[1] += N * (Rp * np.sin(theta) * PH + Rr * (np.cos(theta)) * PHp)
        Zg[:2] /= 7.5
        assert Z == pytest.approx(O.distortion_z(sc, 0, x, y), rel=1e-13, abs=1e-15)
        assert np.allclose(Zg, O.distortion_zgrad(sc, 0, x, y), rtol=1e-12, atol=1e-15)


def test_unsupported_types_are_reported(core):
    """Content the flattener does not know is reported, never skipped."""
    class HolographicFace(core.ctracer.Face):          # no such class in raypier.core
        pass
    fl = core.ctracer.FaceList()
    fl.faces = [HolographicFace()]
    with pytest.raises(SC.UnsupportedSceneError):
        SC.Scene([fl], np.array([1.0]))

    class MagicMaterial(core.ctracer.InterfaceMaterial):
        pass
    fl.faces = [core.cfaces.CircularFace(material=MagicMaterial())]
    with pytest.raises(SC.UnsupportedSceneError):
        SC.Scene([fl], np.array([1.0]))

    # a UVPatchFace whose mesh resolution cannot be read (a genuine raypier one keeps u_res / v_res private)
    cfg = configs.build(core, "uvpatch", n=4)
    face = cfg['face_lists'][0].faces[0]
    face.u_res = None
    face.owner.u_res = None
    with pytest.raises(SC.UnsupportedSceneError):
        SC.Scene(cfg['face_lists'], cfg['wavelengths'])


def test_decomposition_material_flattens_as_an_absorber(core):
    """ResampleGaussletMaterial: eval_child_ray_c only captures (cmaterials.pyx:1808-1821), so the device
    sees an absorber; the callback runs on the host between generations (core/tracer.py)."""
    fl = core.ctracer.FaceList()
    fl.faces = [core.cfaces.CircularFace(material=core.cmaterials.ResampleGaussletMaterial(eval_func=lambda gc: gc))]
    sc = SC.Scene([fl], np.array([1.0]))
    assert sc.materials[0]['type'] == A.MAT_OPAQUE
    assert fl.faces[0].material.is_decomp_material() and fl.faces[0].material.capture_count == 0


def test_subclasses_resolve_through_mro(core):
    class MyLensFace(core.cfaces.SphericalFace):
        pass
    fl = core.ctracer.FaceList()
    fl.faces = [MyLensFace(diameter=10.0, curvature=30.0)]
    sc = SC.Scene([fl], np.array([1.0]))
    assert sc.faces[0]['type'] == A.FACE_SPHERICAL


def test_scene_roundtrip_through_dict(core):
    cfg = configs.build(core, "config3", n=10)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    back = SC.Scene.from_dict(sc.to_dict())
    for t in SC.Scene.TABLES:
        assert np.asarray(getattr(sc, t)).tobytes() == np.asarray(getattr(back, t)).tobytes()
    assert back.c_scene.n_traced_faces == 2 and back.c_scene.n_faces == 3


def test_collections_behave_like_the_reference(core):
    ct = core.ctracer
    rays = configs.disc_source(10, (0, 0, 0), (0, 0, 1), 1.0, seed=1)
    rc = ct.RayCollection.from_array(rays)
    rc.wavelengths = [0.5]
    assert len(rc) == rc.n_rays == 10
    a = rc.copy_as_array()
    a['length'] = 3.0
    assert np.all(np.isinf(rc.length))          # copy_as_array always copies
    rc.reset_length(7.0)
    assert np.all(rc.length == 7.0)
    child = ct.RayCollection.from_array(rays[:4])
    child.parent = rc
    assert child.parent is rc and np.array_equal(child.wavelengths, rc.wavelengths)
    with pytest.raises(ValueError):
        ct.RayCollection.from_array(np.zeros(3))
    gc = ct.GaussletCollection.from_rays(rays)
    assert gc.copy_as_array().dtype == A.gausslet_dtype
    assert np.array_equal(gc.para_origin[:, 3], rays['origin'])


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 8):
            spans = [distributed.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_parent_offsets():
    counts_all = np.array([[4, 4, 7], [3, 5, 5]])
    assert distributed.parent_offsets(counts_all, 0).tolist() == [0, 0, 0]
    assert distributed.parent_offsets(counts_all, 1).tolist() == [4, 4, 7]


def test_mesh_bvh_covers_every_triangle_once():
    """The BVH the host builds for OBBTreeFace meshes (core/obbtree.py::build_bvh): every cell in exactly
    one leaf, children numbered after their parent (what rpx_scene_set validates), every box contains
    the triangles below it."""
    import numpy as np
    from raypier_optics_b200 import configs
    from raypier_optics_b200.core.obbtree import LEAF_CELLS, build_bvh
    import itertools
    meshes = (configs.icosphere(5.0, 2), configs.bowl_mesh(60.0, 17, 30.0), configs.icosphere(1.0, 0))
    for (pts, cells), sah in itertools.product(meshes, (False, True)):   # median split / binned surface-area heuristic
        order, nodes = build_bvh(pts, cells, sah=sah)
        height = np.ones(len(nodes), dtype=int)
        for k in range(len(nodes) - 1, -1, -1):
            if nodes[k, 6] >= 0:
                height[k] = 1 + max(height[int(nodes[k, 6])], height[int(nodes[k, 7])])
        assert height[0] <= 46                                             # what the packed device nodes hold
        assert sorted(order.tolist()) == list(range(len(cells)))
        cover = np.zeros(len(cells), dtype=int)

        def check(k, lo, hi):
            n = nodes[k]
            assert (n[0:3] >= lo - 1e-12).all() and (n[3:6] <= hi + 1e-12).all()  # nested boxes
            if n[6] >= 0:
                assert n[6] > k and n[7] > k
                check(int(n[6]), n[0:3], n[3:6])
                check(int(n[7]), n[0:3], n[3:6])
            else:
                first, count = int(-n[6] - 1), int(n[7])
                assert 1 <= count <= LEAF_CELLS
                cover[first:first + count] += 1
                tri = pts[cells[order[first:first + count]]].reshape(-1, 3)
                assert (tri >= n[0:3]).all() and (tri <= n[3:6]).all()
        check(0, np.full(3, -np.inf), np.full(3, np.inf))
        assert (cover == 1).all()


def test_face_sequence_inference_works_on_mirror_collections(core):
    """BaseRaySource.set_face_sequence (raypier/sources.py:138-155) reads ``rays.base_rays.end_face_idx`` of
    every traced generation and takes the most common face; the host collections must offer that API for
    rays AND gausslets (GaussletBaseRayView is a RayArrayView in the reference, ctracer.pyx:1157)."""
    import numpy as np
    from oracle import oracle as O
    from raypier_optics_b200 import configs, scene as SC
    for gausslets in (False, True):
        cfg = configs.build(core, "config5", n=200, gausslets=gausslets)
        sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
        gens, _ = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
        cls = core.ctracer.GaussletCollection if gausslets else core.ctracer.RayCollection
        traced = [cls.from_array(np.ascontiguousarray(g)) for g in gens]
        all_faces = [f for fl in cfg['face_lists'] for f in fl.faces]
        # the reference's own lines
        global_map = {face: i for i, face in enumerate(all_faces)}
        face_map = {global_map[face]: (fl, fi) for fl in cfg['face_lists'] for fi, face in enumerate(fl.faces)}
        seq = []
        for rays in traced:
            ids, counts = np.unique(rays.base_rays.end_face_idx, return_counts=True)
            most_common = ids[counts.argmax()]
            if most_common not in face_map:
                break
            seq.append(face_map[most_common])
        assert len(seq) >= 4 and seq[0][0] is cfg['face_lists'][0]
        view = traced[1].base_rays
        assert len(view) == len(traced[1]) and view.origin.shape == (len(view), 3)
        assert view.termination.shape == (len(view), 3)


def test_blocked_oracle_equals_whole_oracle(core):
    """The harness of tests/test_parity_fullsize_gpu.py (blocks of the source traced on host threads,
    slices of the whole trace located through the parent chain) on the oracle itself."""
    import numpy as np
    from oracle import oracle as O
    from raypier_optics_b200 import scene as SC
    from util import build_case
    from test_parity_fullsize_gpu import check_against_blocked_oracle
    for name, kw, rl in (("config4_prisms", dict(n=6000), 12), ("config5", dict(n=1500, gausslets=True), None)):
        cfg = build_case(core, name, kw, rl)
        sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
        gens, fc = O.trace_rays(sc, np.ascontiguousarray(cfg['rays']), cfg['recursion_limit'], cfg['max_length'])
        worst = check_against_blocked_oracle(None, cfg, name, block=701, got=(gens, fc))
        assert worst == 0.0


def test_wrap_generations_mutates_genuine_reference_collections_in_place(refcore):
    """traced_rays[0] IS input_rays, with the write-back visible through it, also when the
    caller hands in genuine raypier.core collections (which own malloc'd memory)."""
    import numpy as np
    from raypier_optics_b200 import _abi as A, configs
    from raypier_optics_b200.core.tracer import _wrap_generations
    ct = refcore.ctracer
    for gauss in (False, True):
        cfg = configs.build(refcore, "config5", n=257, gausslets=gauss)
        rays = np.ascontiguousarray(cfg['rays'])
        cls = ct.GaussletCollection if gauss else ct.RayCollection
        rc = cls.from_array(rays.view(ct.gausslet_dtype if gauss else ct.ray_dtype).copy())
        rc.wavelengths = np.asarray(cfg['wavelengths'])
        g0 = rays.copy()
        b = g0['base_ray'] if gauss else g0
        b['length'] = np.linspace(1.0, 2.0, len(g0))
        b['end_face_idx'] = np.arange(len(g0)) % 7
        if gauss:
            g0['para_rays']['length'][:] = 3.25
        g1 = rays[:100].copy()
        out = _wrap_generations(rc, [g0, g1], np.asarray(cfg['wavelengths']))
        assert out[0] is rc and len(rc) == len(g0)
        assert rc.copy_as_array().tobytes() == g0.tobytes()
        assert type(out[1]) is cls and out[1].parent is rc and len(out[1]) == 100
        assert np.array_equal(out[1].wavelengths, cfg['wavelengths'])


class _FakeHostLib(object):
    """malloc / free behind the names of the page-locked allocator (no CUDA on the CPU suite)."""

    def __init__(self, refuse_above=None):
        import ctypes as C
        self._libc = C.CDLL(None)
        self._libc.malloc.restype = C.c_void_p
        self._libc.malloc.argtypes = [C.c_size_t]
        self._libc.free.argtypes = [C.c_void_p]
        self.live = {}
        self.refuse_above = refuse_above

    def rpx_host_alloc(self, nbytes):
        if self.refuse_above is not None and nbytes > self.refuse_above:
            return None
        p = self._libc.malloc(nbytes)
        self.live[p] = nbytes
        return p

    def rpx_host_free(self, p):
        del self.live[p]
        self._libc.free(p)


def test_host_pool_recycles_blocks_when_the_last_view_dies():
    import gc
    from raypier_optics_b200 import _abi as A
    from raypier_optics_b200._hostpool import HostPool, MIN_BYTES
    lib = _FakeHostLib()
    pool = HostPool(lib)
    pool.idle_cap, pool.max_bytes = 64 << 20, 256 << 20
    n = 20000  # 13 MB of gausslets
    a = pool.empty(n, A.gausslet_dtype)
    assert a.shape == (n,) and a.dtype == A.gausslet_dtype and a.flags.writeable and a.flags.c_contiguous
    a['base_ray']['length'] = 3.0
    view = a[5:10]
    ptr = a.ctypes.data
    assert ptr in lib.live and pool.stats()["out_bytes"] >= a.nbytes
    del a
    gc.collect()
    assert pool.stats()["idle_bytes"] == 0  # a view is still alive
    assert float(view['base_ray']['length'][0]) == 3.0
    del view
    gc.collect()
    st = pool.stats()
    assert st["out_bytes"] == 0 and st["idle_bytes"] >= n * 668
    b = pool.empty(n - 7, A.gausslet_dtype)  # nearly the same size: the idle block is reused
    assert b.ctypes.data == ptr and pool.stats()["hits"] == 1
    small = pool.empty(10, A.ray_dtype)      # small results stay plain numpy
    assert small.nbytes < MIN_BYTES and small.ctypes.data not in lib.live
    del b
    gc.collect()
    pool.trim()
    assert not lib.live and pool.stats()["idle_bytes"] == 0


def test_host_pool_falls_back_to_numpy_when_page_locking_is_refused_or_capped():
    import gc
    from raypier_optics_b200 import _abi as A
    from raypier_optics_b200._hostpool import HostPool
    lib = _FakeHostLib(refuse_above=0)
    pool = HostPool(lib)
    a = pool.empty(100000, A.ray_dtype)
    assert a.shape == (100000,) and not lib.live and pool.stats()["fallbacks"] == 1 and pool.stats()["out_bytes"] == 0
    lib2 = _FakeHostLib()
    pool2 = HostPool(lib2)
    pool2.idle_cap, pool2.max_bytes = 64 << 20, 40 << 20
    x = pool2.empty(100000, A.ray_dtype)   # 18.8 MB page-locked
    y = pool2.empty(100000, A.ray_dtype)   # would exceed the 40 MB cap together with its size-class slack? no: fits
    z = pool2.empty(100000, A.ray_dtype)   # this one does not
    assert len(lib2.live) == 2 and pool2.stats()["fallbacks"] == 1
    del x, y, z
    gc.collect()
    # over the cap with only IDLE blocks in the way: they are given back to make room
    w = pool2.empty(200000, A.ray_dtype)   # 37.6 MB
    assert w.ctypes.data in lib2.live and len(lib2.live) == 1
    del w
    gc.collect()
    pool2.trim()
    assert not lib2.live
    os.environ["RPX_PINNED_RESULTS"] = "0"
    try:
        pool3 = HostPool(_FakeHostLib())
        assert pool3.empty(100000, A.ray_dtype).shape == (100000,) and pool3.stats()["misses"] == 0
    finally:
        del os.environ["RPX_PINNED_RESULTS"]


def _emulated_walk(nodes, order, P1, V1, V2, N, o, d, tol=1e-9):
    """The device's ordered segment / BVH walk (rpx_faces.cuh::mesh_intersect) in Python on the fp64 tree: near child
    first, far child stacked with its entry parameter, entries behind the best hit dropped on pop.  Returns
    (best alpha or 1.0, cell id or -1, inner nodes visited, triangles tested)."""
    def slab(k, best):
        with np.errstate(divide='ignore', invalid='ignore'):
            a, b = (nodes[k, 0:3] - o) / d, (nodes[k, 3:6] - o) / d
        tmin = max(np.nanmin([a[0], b[0]]), np.nanmin([a[1], b[1]]), np.nanmin([a[2], b[2]]), 0.0)
        tmax = min(np.nanmax([a[0], b[0]]), np.nanmax([a[1], b[1]]), np.nanmax([a[2], b[2]]), best)
        return tmin <= tmax, tmin
    best, best_id, visits, tests, stack = 1.0, -1, 0, 0, []
    cur = 0
    while True:
        while cur is not None and nodes[cur, 6] >= 0:
            visits += 1
            a, b = int(nodes[cur, 6]), int(nodes[cur, 7])
            (ha, ta), (hb, tb) = slab(a, best), slab(b, best)
            if ha and hb:
                near, far, tf = (a, b, tb) if ta <= tb else (b, a, ta)
                stack.append((far, tf))
                cur = near
            elif ha or hb:
                cur = a if ha else b
            else:
                cur = None
        if cur is not None:
            first, count = int(-nodes[cur, 6] - 1), int(nodes[cur, 7])
            for c in order[first:first + count]:
                tests += 1
                det = -np.dot(d, N[c])
                if det == 0.0:
                    continue
                a0 = o - P1[c]
                da0 = np.cross(a0, d)
                u, v, al = np.dot(V2[c], da0) / det, -np.dot(V1[c], da0) / det, np.dot(a0, N[c]) / det
                if u + v > 1.0 or u < 0 or v < 0 or al < 0:
                    continue
                if al >= tol and (al < best or (al == best and c < best_id)):
                    best, best_id = al, int(c)
        cur = None
        while stack:
            k, t = stack.pop()
            if t <= best:
                cur = k
                break
        if cur is None:
            return best, best_id, visits, tests


def test_sah_tree_gives_the_same_hits_with_fewer_node_visits(core):
    """The BVH only prunes: the median tree, the binned-SAH tree (what ships) and brute force over all triangles
    find the same nearest triangle for the rays of a real mesh trace (generations 0 - 2 of the oracle's trace of
    the mesh scene), and the SAH tree gets there with fewer inner-node visits (the device measurement behind the
    default: profiles/r02_notes.md section 9.8)."""
    from oracle import oracle as O
    from raypier_optics_b200.core.obbtree import build_bvh, triangle_records
    cfg = configs.build(core, "mesh", n=60, ball_subdiv=3, mesh_n=40)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    gens, _ = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    visits = {False: 0, True: 0}
    n_walks = 0
    for si, fl in enumerate(cfg['face_lists']):
        owner = fl.faces[0].owner
        pts, cells = np.asarray(owner.mesh_points, dtype=float), np.asarray(owner.mesh_cells, dtype=np.int32)
        P1, V1, V2, N = triangle_records(pts, cells)
        trees = {sah: build_bvh(pts, cells, sah=sah) for sah in (False, True)}
        it = sc.face_sets[si]['inv_trans']
        R, T = np.array(it['m']).reshape(3, 3), np.array(it['t']).reshape(3)
        for g in gens[:3]:
            o = g['origin'] @ R.T + T
            e = (g['origin'] + g['direction'] * cfg['max_length']) @ R.T + T
            for i in range(0, len(g), 2):
                d = e[i] - o[i]
                res = {}
                for sah, (order, nodes) in trees.items():
                    al, cid, v, _t = _emulated_walk(nodes, order, P1, V1, V2, N, o[i], d)
                    res[sah] = (al, cid)
                    visits[sah] += v
                assert res[False] == res[True]
                # brute force over every triangle
                det = -(N @ d)
                with np.errstate(divide='ignore', invalid='ignore'):
                    a0 = o[i] - P1
                    da0 = np.cross(a0, d)
                    u = np.einsum('ij,ij->i', V2, da0) / det
                    v = -np.einsum('ij,ij->i', V1, da0) / det
                    al = np.einsum('ij,ij->i', a0, N) / det
                ok = (det != 0) & ~(u + v > 1.0) & ~(u < 0) & ~(v < 0) & ~(al < 0) & (al >= 1e-9) & (al < 1.0)
                want = (1.0, -1)
                if ok.any():
                    amin = al[ok].min()
                    want = (float(amin), int(np.flatnonzero(ok & (al == amin))[0]))
                assert res[True][1] == want[1] and abs(res[True][0] - want[0]) <= 1e-12 * abs(want[0])  # (summation order)
                n_walks += 1
    assert n_walks > 100
    assert visits[True] < 0.97 * visits[False], visits
