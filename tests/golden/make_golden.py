#!/usr/bin/env python
"""Generate the committed golden vectors from the REAL reference (the unmodified
raypier/core built into oracle/_ref by oracle/build_ref.sh).  Run in the build container:

    python tests/golden/make_golden.py

For every parity case: the seeded input rays, the flattened scene tables (built from the
genuine reference objects), every generation the reference's own trace returns and its
Face.count values.  Small N so the fixtures stay small; the large-N parity runs compare
CUDA against the oracle, which test_oracle_vs_reference.py pins bit-exact on the same cases.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from raypier_optics_b200 import scene as SC  # noqa: E402

GOLDEN_CASES = [
    ("config1", dict(n=160), None),
    ("config2", dict(n=160), None),
    ("config2_lowthr", dict(n=96, reflection_threshold=1e-3, transmission_threshold=1e-3), 5),
    ("config3", dict(n=160), None),
    ("config4_prisms", dict(n=96), 10),
    ("config4_grating", dict(n=160), None),
    ("config5_rays", dict(n=96, gausslets=False), None),
    ("config5", dict(n=40, gausslets=True), None),
    ("zoo_rays", dict(n=1560, gausslets=False), None),
    ("zoo", dict(n=390, gausslets=True), None),
    # > 232 faces: scene tables beyond the shared-memory staging budget (SS=false kernels)
    ("big_scene_rays", dict(n=96, gausslets=False), None),
    ("big_scene", dict(n=40, gausslets=True), None),
    # OBBTreeFace triangle meshes (coarser mirror mesh: the scene tables travel with the fixture)
    ("mesh_rays", dict(n=160, gausslets=False, mesh_n=14), None),
    ("mesh", dict(n=40, gausslets=True, mesh_n=14), None),
    # UVPatchFace over a Bezier and a B-spline patch (cbezier.pyx; imported with the numpy.math alias)
    ("uvpatch_rays", dict(n=200, gausslets=False), None),
    ("uvpatch", dict(n=60, gausslets=True), None),
]


def main():
    from util import build_case
    core = O.import_reference("parity")
    if core is None:
        raise SystemExit("reference not built: run oracle/build_ref.sh first")
    only = set(sys.argv[1:])  # optional: regenerate just these tags
    for tag, kw, rl in GOLDEN_CASES:
        if only and tag not in only:
            continue
        name = tag.replace("_lowthr", "").replace("_rays", "")
        cfg = build_case(core, name, kw, rl)
        sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
        rc = O.reference_collection(core, cfg['rays'], cfg['wavelengths'])
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):  # UVPatchFace.intersect_c prints a line per hit (:524)
            traced, all_faces = O.reference_trace_rays(core, rc, cfg['face_lists'], cfg['recursion_limit'],
                                                       cfg['max_length'])
        out = {"input": cfg['rays'], "max_length": np.array(cfg['max_length']),
               "recursion_limit": np.array(cfg['recursion_limit']),
               "face_counts": np.array([f.count for f in all_faces], dtype=np.uint32),
               "n_generations": np.array(len(traced))}
        for g, t in enumerate(traced):
            a = t.copy_as_array()
            out["gen%02d" % g] = np.frombuffer(a.tobytes(), dtype=np.uint8)
        for k, v in sc.to_dict().items():
            out["scene_" + k] = v
        path = os.path.join(HERE, tag + ".npz")
        np.savez_compressed(path, **out)
        print("%-18s generations %s -> %s (%.0f kB)" % (tag, [len(t) for t in traced], os.path.basename(path),
                                                       os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
