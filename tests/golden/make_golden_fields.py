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
    make_hexgrid(core, F)


def make_hexgrid(core, F):
    """The plain-ray front end: project_to_sphere / evaluate_neighbours / eval_Efield_from_rays
    (fields.py:50-111, 206-229) on a hexagonal grid of rays with neighbour lists."""
    from raypier_optics_b200 import configs
    from test_fields import hexgrid_points
    rays, nb = configs.hex_grid_source(n_side=17)
    wl = np.array([1.0])
    pts = hexgrid_points()
    centre, radius, blending, time_ps = (0.3, -0.2, 80.0), 90.0, 0.8, 1.5
    projected = F.project_to_sphere(rays.copy(), centre, radius)
    kept, x, y, dx, dy = F.evaluate_neighbours(projected.copy(), nb)
    from raypier.core import cfields
    modes = cfields.evaluate_modes(x, y, dx, dy, blending=blending)
    rc = O.reference_collection(core, rays.copy(), wl)
    rc.neighbours = nb
    E_plain = F.eval_Efield_from_rays(rc, pts, wl, blending=blending, time_ps=time_ps)
    E_sphere = F.eval_Efield_from_rays(rc, pts, wl, blending=blending, time_ps=time_ps, exit_pupil_offset=radius,
                                       exit_pupil_centre=centre)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fields_hexgrid.npz")
    np.savez_compressed(out, rays=rays.view(np.uint8), neighbours=nb, points=pts, wavelengths=wl, centre=np.array(centre),
                        radius=radius, blending=blending, time_ps=time_ps, projected=projected.view(np.uint8),
                        x=x, y=y, dx=dx, dy=dy, modes=modes, E_plain=E_plain, E_sphere=E_sphere)
    print("wrote", out, len(rays), "rays", len(kept), "with six neighbours", "max|E| %.4g / %.4g"
          % (np.abs(E_plain).max(), np.abs(E_sphere).max()))


if __name__ == "__main__":
    main()
