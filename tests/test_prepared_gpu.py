"""GPU tests of code paths that were written after round 1's GPU minutes had run out and have therefore
NEVER run on a GPU.  They carry the marker ``gpu_prepared`` (not ``gpu``), so neither ``-m gpu`` nor
``-m "not gpu"`` on a GPU-less machine runs them (the latter selects them, and they skip themselves).
Run ``pytest -m gpu_prepared`` on a B200 first; once green, move each test next to its siblings under the
``gpu`` marker."""
import numpy as np
import pytest


def _engine_or_skip():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from raypier_optics_b200.engine import get_engine
    return get_engine(0)


@pytest.mark.gpu_prepared
@pytest.mark.parametrize("name,kw,rl,chunk", [
    ("config2", dict(n=20000, reflection_threshold=1e-3, transmission_threshold=1e-3), 6, 4096),
    ("config4_prisms", dict(n=5000), 12, 777),
])
def test_streamed_trace_in_place_returns_the_same_generation_0(core, name, kw, rl, chunk):
    """rpx_trace_streamed with out[0] aliasing the source (the reference's in-place convention): every
    generation byte-identical to the call with a separate generation-0 buffer."""
    from util import build_case
    from raypier_optics_b200 import scene as SC
    engine = _engine_or_skip()
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
