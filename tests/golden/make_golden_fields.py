"""Golden vectors for the E-field summation, produced by the REAL reference:
raypier/core/fields.py (imported in place from /root/reference) over the compiled cfields of
oracle/_ref.  Run here (the GPU box has neither):  python tests/golden/make_golden_fields.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402


def main():
    core = O.import_reference("parity")
    F = O.reference_fields(core)
    assert core is not None and F is not None, "build the reference first: oracle/build_ref.sh"
    from test_fields import michelson_output
    cfg, g, pts = michelson_output(core, n=500, seed=21)
    gc = core.ctracer.GaussletCollection.from_array(g.view(core.ctracer.gausslet_dtype))
    gc.wavelengths = np.asarray(cfg['wavelengths'])
    blending, time_ps = 0.9, 1.25
    modes = F.ExtractGamma(gc, blending=blending)
    E = F.eval_Efield_from_gausslets(gc, pts, blending=blending, time_ps=time_ps)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fields_michelson.npz")
    np.savez_compressed(out, gausslets=g.view(np.uint8), points=pts, wavelengths=np.asarray(cfg['wavelengths']),
                        blending=blending, time_ps=time_ps, modes=modes, E=E)
    print("wrote", out, len(g), "gausslets", len(pts), "points", "max|E| %.4g" % np.abs(E).max())


if __name__ == "__main__":
    main()
