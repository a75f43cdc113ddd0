"""The device-side consumers of a trace and the chunked trace that feeds them
(rpx_trace_consume): terminal-ray selection, device-resident AoS export / import, the accumulating
detector and the capture plane, each against the oracle's trace of the same seeded source."""
import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import configs, scene as SC
from raypier_optics_b200.engine import ConsumeResult

from util import build_case, compare_generation

pytestmark = pytest.mark.gpu


def _base(a):
    return a['base_ray'] if a.dtype == A.gausslet_dtype else a


def _oracle(cfg):
    from oracle import oracle as O
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    gens, fc = O.trace_rays(sc, np.ascontiguousarray(cfg['rays']), cfg['recursion_limit'], cfg['max_length'])
    return sc, gens, fc


def _capture_scene(core, centre, direction, size):
    face = core.cfaces.RectangularFace(length=size[0], width=size[1], offset=0.0, z_plane=0.0)
    fl = core.ctracer.FaceList(owner=configs.Pose(centre=centre, direction=direction))
    fl.faces = [face]
    fl.sync_transforms()
    return SC.Scene([fl], np.asarray([1.0]))


def _terminal_of(gens, unterminated, faces):
    """The reference-side definition: rays whose end_face_idx stayed (unsigned)-1 after the write-back
    (ctracer.pyx:2086-2087), and / or rays that end on one of ``faces``; (generation, ray) order."""
    parts = []
    for g in gens:
        ef = _base(g)['end_face_idx']
        sel = np.zeros(len(g), dtype=bool)
        if unterminated:
            sel |= ef == A.NO_FACE
        if faces:
            sel |= np.isin(ef, list(faces))
        parts.append(g[sel])
    return parts


@pytest.mark.parametrize("name,kw,rl,faces", [
    ("config4_prisms", dict(n=6000), 12, None),
    ("config2", dict(n=8000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6, (1,)),
    ("config5", dict(n=2000, gausslets=True), None, (0, 3)),
])
def test_terminal_selection_matches_oracle(engine, core, name, kw, rl, faces):
    cfg = build_case(core, name, kw, rl)
    sc, want_gens, _ = _oracle(cfg)
    engine.set_scene(sc)
    res = engine.trace(np.ascontiguousarray(cfg['rays']), cfg['max_length'], cfg['recursion_limit'])
    is_g = cfg['rays'].dtype == A.gausslet_dtype
    for unterminated in (True, False):
        if not unterminated and not faces:
            continue
        dev, counts = engine.select_terminal(res.device_generations(), is_g, unterminated=unterminated, faces=faces)
        want = _terminal_of(want_gens, unterminated, faces)
        assert counts == [len(w) for w in want]
        got = engine.download(dev)
        assert len(got) == sum(counts) and sum(counts) > 0
        compare_generation(got, np.concatenate(want), "%s terminal rays" % name)
        # device-resident AoS round trip (the NCCL send / receive buffers of distributed.gather_terminal)
        import torch
        t = torch.empty(len(got) * got.dtype.itemsize + 4, dtype=torch.uint8, device="cuda")
        engine.export_device(dev, t.data_ptr(), len(got))
        assert t[:len(got) * got.dtype.itemsize].cpu().numpy().tobytes() == got.tobytes()
        back = engine.import_device(t.data_ptr(), len(got), is_g)
        assert engine.download(back).tobytes() == got.tobytes()
        back.free()
        dev.free()
    res.free()


def test_detector_accumulates_like_one_field_evaluation(engine, core):
    """rpx_detector: collections summed one after the other equal one EFieldSummation over all of them
    (fields.py:206-249), and the oracle's sum_gaussian_modes."""
    from oracle import oracle as O
    cfg = configs.build(core, "config5", n=1500, gausslets=True, seed=3)
    sc, gens, _ = _oracle(cfg)
    last = np.ascontiguousarray(gens[-1][:3000])
    xs = np.linspace(-3.0, 3.0, 12)
    gx, gz = np.meshgrid(xs, xs)
    pts = np.ascontiguousarray(np.stack([gx.ravel(), np.full(gx.size, -14.0), gz.ravel()], axis=1))
    want = O.eval_Efield_from_gausslets(last, pts, cfg['wavelengths'])
    det = engine.detector(pts, cfg['wavelengths'])
    for part in np.array_split(last, 4):
        d = engine.upload(np.ascontiguousarray(part))
        det.accumulate(d)
        d.free()
    got = det.read()
    assert det.modes == len(last) and det.ms > 0
    scale = np.abs(want).max()
    assert np.abs(got - want).max() / scale < 1e-10
    det.reset()
    assert det.modes == 0 and np.all(det.read() == 0)
    det.free()


@pytest.mark.parametrize("name,kw,rl,chunk,plane,source_on_device", [
    ("config5", dict(n=3000, gausslets=True), None, 700,
     dict(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0), size=(12.0, 12.0)), False),
    ("config5", dict(n=3000, gausslets=True), None, 1024,
     dict(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0), size=(12.0, 12.0)), True),
    ("config2", dict(n=20000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6, 4096,
     dict(centre=(-4.86, -31.3, 0.07), direction=(0.2, 1.0, 0.1), size=(30.0, 30.0)), False),
    ("config4_prisms", dict(n=5000), 12, 777,
     dict(centre=(0.0, 0.0, 0.0), direction=(1.0, 0.3, 0.0), size=(400.0, 400.0)), True),
])
def test_trace_consume_matches_oracle(engine, core, name, kw, rl, chunk, plane, source_on_device):
    """rpx_trace_consume: counts, Face.count, terminal rays, captured rays (restored to the reference's
    (generation, ray) order) and the detector field of a chunked trace whose generations never leave the
    GPU, against the oracle's one-shot trace + select_*_intersections + sum_gaussian_modes."""
    from oracle import oracle as O
    cfg = build_case(core, name, kw, rl)
    sc, want_gens, want_fc = _oracle(cfg)
    rays = np.ascontiguousarray(cfg['rays'])
    is_g = rays.dtype == A.gausslet_dtype
    cap_scene = _capture_scene(core, plane['centre'], plane['direction'], plane['size'])
    want_cap, _, want_cap_counts = O.select_intersections(cap_scene, want_gens, [cfg['wavelengths']] * len(want_gens))
    want_term = _terminal_of(want_gens, True, None)
    assert len(rays) > 2 * chunk and len(want_cap) > 0
    engine.set_scene(sc)
    engine.set_capture_scene(cap_scene)
    det = None
    if is_g:
        xs = np.linspace(-2.0, 2.0, 8)
        gx, gz = np.meshgrid(xs, xs)
        pts = np.ascontiguousarray(np.stack([gx.ravel(), np.full(gx.size, -14.0), gz.ravel()], axis=1))
        det = engine.detector(pts, cfg['wavelengths'])
    src, keep = rays, None
    if source_on_device:
        import torch
        keep = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
        torch.cuda.synchronize()
        src = keep.data_ptr()
    out = engine.trace_consume(src, cfg['max_length'], cfg['recursion_limit'], n=len(rays), is_gausslet=is_g,
                               chunk_rays=chunk, terminal=True, terminal_capacity=sum(len(t) for t in want_term) + 16,
                               capture=True, captured_capacity=len(want_cap) + 16, detector=det, per_chunk=True)
    assert out.counts == [len(g) for g in want_gens]
    assert out.n_chunks == -(-len(rays) // chunk)
    assert np.array_equal(out.face_counts, want_fc)
    assert out.device_ms > 0 and out.trace_ms > 0 and out.launches > out.n_chunks * (len(want_gens) + 1)
    # terminal rays
    assert out.n_terminal == sum(len(t) for t in want_term) == len(out.terminal)
    got_term = ConsumeResult.reference_order(engine.download(out.terminal), out.per_chunk_terminal)
    compare_generation(got_term, np.concatenate(want_term), name + " terminal (chunked)")
    # captured rays
    assert out.n_captured == len(want_cap) == len(out.captured)
    assert out.per_chunk_captured.sum(axis=0).tolist() == list(want_cap_counts)
    got_cap = ConsumeResult.reference_order(engine.download(out.captured), out.per_chunk_captured)
    compare_generation(got_cap, want_cap, name + " captured (chunked)")
    # detector field
    if det is not None:
        got_E = det.read()
        assert det.modes == len(want_cap)
        # (a) the summation itself: the oracle's sum over the SAME (CUDA-traced) gausslets, 1e-10
        same_E = O.eval_Efield_from_gausslets(np.ascontiguousarray(got_cap), pts, cfg['wavelengths'])
        assert np.abs(got_E - same_E).max() / np.abs(same_E).max() < 1e-10
        # (b) the whole pipeline against the oracle's own trace: the optical phase is k * path with
        # k = 2 pi / 1e-3 mm, so the 1e-9 relative tolerance of traced positions / paths (~50 mm) allows
        # ~3e-4 rad of phase per gausslet; measured 3e-10 of the field maximum
        want_E = O.eval_Efield_from_gausslets(want_cap, pts, cfg['wavelengths'])
        assert np.abs(got_E - want_E).max() / np.abs(want_E).max() < 1e-8
        det.free()
    out.free()
    # counting without keeping: capacities of 0
    lean = engine.trace_consume(src, cfg['max_length'], cfg['recursion_limit'], n=len(rays), is_gausslet=is_g,
                                chunk_rays=chunk, terminal=True, capture=True)
    assert lean.terminal is None and lean.captured is None
    assert (lean.n_terminal, lean.n_captured, lean.counts) == (out.n_terminal, out.n_captured, out.counts)
    # a capacity that is too small is reported, never overrun
    from raypier_optics_b200._lib import RpxError
    with pytest.raises(RpxError):
        engine.trace_consume(src, cfg['max_length'], cfg['recursion_limit'], n=len(rays), is_gausslet=is_g,
                             chunk_rays=chunk, capture=True, captured_capacity=max(len(want_cap) // 3, 1))


def test_trace_consume_edge_cases(engine, core):
    """Empty and ragged sources, a chunk larger than the source, and the misuse errors of rpx_trace_consume."""
    from raypier_optics_b200._lib import RpxError
    cfg = configs.build(core, "config1", n=1000)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    engine.set_scene(sc)
    rays = np.ascontiguousarray(cfg['rays'])
    empty = engine.trace_consume(rays[:0], cfg['max_length'], cfg['recursion_limit'], terminal=True)
    assert empty.counts == [] and empty.n_chunks == 0 and empty.n_terminal == 0 and empty.segments == 0
    for n, chunk in ((1, 0), (129, 128), (1000, 4096), (1000, 1)):
        if chunk == 1 and n > 64:
            n = 64  # one ray per chunk: 64 chunks
        r = engine.trace_consume(np.ascontiguousarray(rays[:n]), cfg['max_length'], cfg['recursion_limit'], chunk_rays=chunk,
                                 terminal=True, terminal_capacity=n + 8)
        assert r.counts == [n, n, n] and r.n_terminal == n == len(r.terminal)   # every ray leaves through the lens
        got = engine.download(r.terminal)
        assert np.all(got['end_face_idx'] == A.NO_FACE) and np.array_equal(np.sort(got['ray_ident']), np.arange(n))
        r.free()
    # terminal faces: rays ending on face 1 (the second lens surface) are generation 1
    r = engine.trace_consume(rays, cfg['max_length'], cfg['recursion_limit'], terminal_faces=[1], terminal_capacity=2000)
    assert r.n_terminal == 1000 and np.all(engine.download(r.terminal)['end_face_idx'] == 1)
    r.free()
    with pytest.raises(RpxError):   # a detector needs gausslets
        det = engine.detector(np.zeros((4, 3)), cfg['wavelengths'])
        try:
            engine.trace_consume(rays, cfg['max_length'], cfg['recursion_limit'], detector=det)
        finally:
            det.free()
    # zero-length export / import
    import torch
    res = engine.trace(rays, cfg['max_length'], cfg['recursion_limit'])
    dev, counts = engine.select_terminal(res.device_generations()[:1], False, unterminated=True)  # all of generation 0 hit the lens
    res.free()
    assert counts == [0] and len(dev) == 0
    t = torch.empty(8, dtype=torch.uint8, device="cuda")
    engine.export_device(dev, t.data_ptr(), 0)
    back = engine.import_device(t.data_ptr(), 0, False)
    assert len(back) == 0
    back.free()
    dev.free()
