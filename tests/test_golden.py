"""Committed golden vectors (tests/golden/*.npz, generated from the REAL reference by
tests/golden/make_golden.py): the oracle must reproduce them on the CPU, the CUDA path
through the C ABI must reproduce them on the GPU, and the host mirrors must flatten to the
same scene tables the genuine reference objects flattened to."""
import glob
import os

import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import scene as SC

from util import build_case, compare_traces

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not os.path.basename(f).startswith("fields_"))  # fields_*: tests/test_fields.py
IDS = [os.path.basename(f)[:-4] for f in FILES]


def load(path):
    z = np.load(path)
    rays = z["input"]
    dtype = A.gausslet_dtype if rays.dtype.itemsize == 668 else A.ray_dtype
    rays = np.ascontiguousarray(rays).view(dtype) if rays.dtype != dtype else rays
    gens = [np.frombuffer(z["gen%02d" % g].tobytes(), dtype=dtype) for g in range(int(z["n_generations"]))]
    scene = SC.Scene.from_dict({k[len("scene_"):]: z[k] for k in z.files if k.startswith("scene_")})
    return dict(rays=rays, gens=gens, scene=scene, max_length=float(z["max_length"]),
                recursion_limit=int(z["recursion_limit"]), face_counts=z["face_counts"])


def test_golden_files_present():
    assert len(FILES) >= 8


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_reproduces_reference_golden(path):
    from oracle import oracle as O
    g = load(path)
    gens, counts = O.trace_rays(g["scene"], g["rays"], g["recursion_limit"], g["max_length"])
    worst = compare_traces(gens, g["gens"], os.path.basename(path))
    assert np.array_equal(counts, g["face_counts"])
    # same compiler flags and libm as the reference build: expect bit-identical floats too
    assert worst <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_reproduces_reference_golden(engine, path):
    g = load(path)
    engine.set_scene(g["scene"])
    res = engine.trace(g["rays"], g["max_length"], g["recursion_limit"])
    got = res.generations()
    keep = None
    if os.path.basename(path).startswith("zoo"):  # see util.noisy_face_scales
        from util import check_noisy_faces_on_surface, noisy_face_scales
        keep = noisy_face_scales(g["scene"], g["gens"])
        check_noisy_faces_on_surface(g["scene"], got, g["gens"], os.path.basename(path))
    compare_traces(got, g["gens"], os.path.basename(path), keep=keep)
    assert np.array_equal(res.face_counts, g["face_counts"])
    res.free()


MIRROR_CASES = {
    "config1": ("config1", dict(n=160), None),
    "config2": ("config2", dict(n=160), None),
    "config2_lowthr": ("config2", dict(n=96, reflection_threshold=1e-3, transmission_threshold=1e-3), 5),
    "config3": ("config3", dict(n=160), None),
    "config4_prisms": ("config4_prisms", dict(n=96), 10),
    "config4_grating": ("config4_grating", dict(n=160), None),
    "config5_rays": ("config5", dict(n=96, gausslets=False), None),
    "config5": ("config5", dict(n=40, gausslets=True), None),
    "uvpatch_rays": ("uvpatch", dict(n=200, gausslets=False), None),
}


@pytest.mark.parametrize("tag", sorted(MIRROR_CASES))
def test_host_mirrors_flatten_like_reference_objects(core, tag):
    """The golden scene tables came from genuine raypier.core objects; this package's host
    mirrors must flatten to byte-identical tables and generate identical source rays."""
    name, kw, rl = MIRROR_CASES[tag]
    g = load(os.path.join(GOLDEN_DIR, tag + ".npz"))
    cfg = build_case(core, name, kw, rl)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want = g["scene"]
    for t in SC.Scene.TABLES:
        a, b = np.asarray(getattr(sc, t)), np.asarray(getattr(want, t))
        assert a.tobytes() == b.tobytes(), "scene table %s differs" % t
    assert cfg['rays'].tobytes() == g["rays"].tobytes()
