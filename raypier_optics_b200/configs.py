"""Synthetic restatements of the BASELINE.json configs.

Every builder takes ``core`` -- a namespace exposing ``cfaces, cmaterials, ctracer,
cshapes, cdistortions, cimplicit_surfs`` -- so the *same* model can be assembled
from this package's host mirrors (``raypier_optics_b200.core``) or from the genuine
reference (``raypier.core`` built into oracle/_ref), which is how the parity pins
and the reference arm of bench.py get identical scenes.

Glass coefficients are the Schott Sellmeier (formula 2) rows of the reference's
sqlite glass database (raypier/material_data/glass_dispersion_database.db), copied
here as data because that file does not exist on the GPU box.
"""
import math

import numpy as np

from ._abi import GAUSSLET, gausslet_dtype, ray_dtype

# name -> (formula_id, wavelength_min, wavelength_max, coefs)
GLASS = {
    "N-LAK22": (2, 0.31, 2.5, [0.0, 1.14229781, 0.00585778594, 0.535138441, 0.0198546147,
                               1.04088385, 100.834017]),
    "N-SF6": (2, 0.37, 2.5, [0.0, 1.77931763, 0.0133714182, 0.338149866, 0.0617533621,
                             2.08734474, 174.01759]),
    "N-BK7": (2, 0.3, 2.5, [0.0, 1.03961212, 0.00600069867, 0.231792344, 0.0200179144,
                            1.01046945, 103.560653]),
    "N-SF11": (2, 0.37, 2.5, [0.0, 1.73759695, 0.013188707, 0.313747346, 0.0623068142,
                              1.89878101, 155.23629]),
}


def glass_curve(core, name, absorption=0.0):
    """NamedDispersionCurve(name) of raypier/dispersion.py:66-99."""
    fid, wmin, wmax, coefs = GLASS[name]
    return core.cmaterials.BaseDispersionCurve(fid, np.array(coefs, dtype=np.double), absorption,
                                               wmin, wmax)


def nondispersive(core, n=1.0, absorption=0.0):
    """NondispersiveCurve of raypier/dispersion.py:21-40."""
    return core.cmaterials.BaseDispersionCurve(0, np.array([n], dtype=np.double), absorption, 0.0,
                                               1000000.0)


# ---------------------------------------------------------------------------------
# Optic pose: what raypier.bases.Traceable builds in a tvtk.Transform
# (bases.py:195-203: translate(centre) . rotate_z(o) . rotate_x(e) . rotate_z(rotation),
#  with (o, e) from the direction vector, bases.py:246-252)
# ---------------------------------------------------------------------------------
class _Matrix(object):
    def __init__(self, m):
        self._m = np.asarray(m, dtype=np.double)

    def get_element(self, i, j):
        return float(self._m[i, j])


class _VtkLikeTransform(object):
    def __init__(self, m):
        self.matrix = _Matrix(m)
        self._m = np.asarray(m, dtype=np.double)

    @property
    def linear_inverse(self):
        R = self._m[:3, :3]
        t = self._m[:3, 3]
        inv = np.eye(4)
        inv[:3, :3] = R.T
        inv[:3, 3] = -(R.T @ t)
        return _VtkLikeTransform(inv)


class Pose(object):
    """Duck-typed FaceList owner: exposes ``.transform.matrix.get_element(i, j)`` and
    ``.transform.linear_inverse.matrix`` -- all FaceList.sync_transforms touches
    (ctracer.pyx:1820-1837).  Extra keyword arguments become attributes (what
    Face.update() copies from its owner, e.g. diameter / offset)."""

    def __init__(self, centre=(0., 0., 0.), direction=(0., 0., 1.), rotation=0.0, **attrs):
        x, y, z = np.asarray(direction, dtype=np.double) / np.linalg.norm(direction)
        theta = 180 * math.acos(z) / math.pi
        phi = 180 * math.atan2(x, y) / math.pi
        o, e = -phi, -theta

        def rz(deg):
            a = math.radians(deg)
            c, s = math.cos(a), math.sin(a)
            return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])

        def rx(deg):
            a = math.radians(deg)
            c, s = math.cos(a), math.sin(a)
            return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])

        T = np.eye(4)
        T[:3, 3] = centre
        self.transform = _VtkLikeTransform(T @ rz(o) @ rx(e) @ rz(rotation))
        self.centre = tuple(centre)
        self.direction = (x, y, z)
        for k, v in attrs.items():
            setattr(self, k, v)


# ---------------------------------------------------------------------------------
# Synthetic sources
# ---------------------------------------------------------------------------------
def _frame(axis):
    w = np.asarray(axis, dtype=np.double)
    w = w / np.linalg.norm(w)
    a = np.array([1.0, 0.0, 0.0]) if abs(w[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    u = np.cross(w, a)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    return u, v, w


def disc_source(n, centre, axis, radius, seed, n_wavelengths=1, E_vector=None, E1=1.0, E2=0.0,
                jitter=0.0, ray_type_id=0, gaussian_sigma=None):
    """Collimated disc of ``n`` rays, uniformly random in area (numpy default_rng(seed)),
    travelling along ``axis``; the synthetic stand-in for ParallelRaySource
    (raypier/sources.py:572-608).  Returns a ``ray_dtype`` array."""
    rng = np.random.default_rng(seed)
    u, v, w = _frame(axis)
    r = radius * np.sqrt(rng.random(n))
    phi = 2 * np.pi * rng.random(n)
    rays = np.zeros(n, dtype=ray_dtype)
    rays['origin'] = (np.asarray(centre, dtype=np.double)[None, :] + (r * np.cos(phi))[:, None] * u[None, :]
                      + (r * np.sin(phi))[:, None] * v[None, :])
    d = np.repeat(w[None, :], n, axis=0)
    if jitter:
        d = d + rng.normal(0.0, jitter, size=(n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
    rays['direction'] = d
    rays['E_vector'] = u if E_vector is None else np.asarray(E_vector, dtype=np.double)
    amp = 1.0
    if gaussian_sigma is not None:
        amp = np.exp(-(r * r) / (gaussian_sigma * gaussian_sigma))
    rays['E1_amp'] = E1 * amp
    rays['E2_amp'] = E2 * amp
    rays['refractive_index'] = 1.0
    rays['normal'] = (0.0, 1.0, 0.0)
    rays['length'] = np.inf
    if n_wavelengths > 1:
        rays['wavelength_idx'] = rng.integers(0, n_wavelengths, size=n, dtype=np.uint32)
    rays['ray_ident'] = np.arange(n, dtype=np.uint32)
    rays['ray_type_id'] = ray_type_id
    return rays


def hex_grid_source(n_side=21, pitch=0.4, focus=80.0, wavelength_idx=0):
    """Plain rays on a hexagonal lattice with their six-neighbour lists -- the input of
    eval_Efield_from_rays (core/fields.py:206-229; HexagonalRayFieldSource-like, sources.py): a slightly
    converging spherical wavefront (focus at z = `focus`), elliptical polarisation, rim rays with
    missing neighbours (-1).  Returns (ray_dtype array, int32 N x 6 neighbour indices)."""
    from ._abi import ray_dtype
    idx = np.arange(n_side * n_side).reshape(n_side, n_side)
    n = n_side * n_side
    rays = np.zeros(n, dtype=ray_dtype)
    nb = -np.ones((n, 6), dtype=np.int32)
    offs = [(1, 0), (0, 1), (-1, 1), (-1, 0), (0, -1), (1, -1)]
    for i in range(n_side):
        for j in range(n_side):
            k = idx[i, j]
            x = pitch * (i + 0.5 * j - 0.75 * n_side)
            y = pitch * (np.sqrt(3) / 2 * j - 0.43 * n_side)
            rays['origin'][k] = (x, y, 0.0)
            d = np.array([-x, -y, focus])
            d /= np.sqrt((d ** 2).sum())
            rays['direction'][k] = d
            e = np.cross(d, np.cross([1.0, 0.0, 0.0], d))
            rays['E_vector'][k] = e / np.sqrt((e ** 2).sum())
            for m, (di, dj) in enumerate(offs):
                ii, jj = i + di, j + dj
                if 0 <= ii < n_side and 0 <= jj < n_side:
                    nb[k, m] = idx[ii, jj]
    rays['refractive_index'] = 1.0 + 0j
    rays['E1_amp'] = 1.0
    rays['E2_amp'] = 0.3j
    rays['length'] = 100.0
    rays['wavelength_idx'] = wavelength_idx
    rays['accumulated_path'] = np.linspace(0.0, 0.01, n)
    rays['end_face_idx'] = 0xFFFFFFFF
    return rays, nb


# ---------------------------------------------------------------------------------
# Config builders.  Each returns a dict:
#   face_lists, rays (ray_dtype | gausslet_dtype array), wavelengths, max_length,
#   recursion_limit, name
# ---------------------------------------------------------------------------------
def _facelist(core, owner, faces):
    fl = core.ctracer.FaceList(owner=owner)
    fl.faces = faces
    fl.sync_transforms()
    for f in faces:
        f.update()  # owner -> face parameter copy that trace_rays performs (core/tracer.py:32)
    return fl


def config1_singlet(core, n=10000, seed=0):
    """Config 1: PlanoConvexLens (raypier/lenses.py:50-103) = CircularFace +
    SphericalFace with one SingleLayerCoatedMaterial (coating thickness stays at the
    0.1 um default, quirk Q9), collimated disc source, one wavelength."""
    F, M = core.cfaces, core.cmaterials
    owner = Pose(centre=(0., -20., 0.), direction=(0., 1., 0.), diameter=25.4, offset=0.0)
    mat = M.SingleLayerCoatedMaterial(n_inside=1.5, n_outside=1.0, n_coating=1.3)
    f1 = F.CircularFace(owner=owner, diameter=25.4, material=mat)
    f2 = F.SphericalFace(owner=owner, diameter=25.4, material=mat, z_height=6.0, curvature=20.0)
    fl = _facelist(core, owner, [f1, f2])
    rays = disc_source(n, centre=(0., -50., 0.), axis=(0., 1., 0.), radius=10.0, seed=seed,
                       E_vector=(1., 0., 0.))
    return dict(name="config1_singlet", face_lists=[fl], rays=rays, wavelengths=np.array([0.78]),
                max_length=200.0, recursion_limit=200)


def config2_achromat(core, n=1000000, seed=0, reflection_threshold=0.1, transmission_threshold=0.1):
    """Config 2: EdmundOptic45805 achromatic doublet (raypier/achromats.py:244-298,
    423-435): 3 SphericalFace, 3 CoatedDispersiveMaterial (N-LAK22 / N-SF6 / air,
    MgF2-like n=1.37 coating 0.283 um on the outer faces), 5 wavelengths, elliptical
    polarisation."""
    F, M = core.cfaces, core.cmaterials
    owner = Pose(centre=(1.0, -2.0, 3.0), direction=(0.2, 1.0, 0.1), diameter=25.4)
    air = nondispersive(core, 1.0)
    d1, d2 = glass_curve(core, "N-LAK22"), glass_curve(core, "N-SF6")
    coat = nondispersive(core, 1.37)
    kw = dict(reflection_threshold=reflection_threshold,
              transmission_threshold=transmission_threshold)
    m1 = M.CoatedDispersiveMaterial(dispersion_inside=d1, dispersion_outside=air,
                                    dispersion_coating=coat, coating_thickness=0.283, **kw)
    m2 = M.CoatedDispersiveMaterial(dispersion_inside=d2, dispersion_outside=d1,
                                    dispersion_coating=coat, coating_thickness=0.0, **kw)
    m3 = M.CoatedDispersiveMaterial(dispersion_inside=air, dispersion_outside=d2,
                                    dispersion_coating=coat, coating_thickness=0.283, **kw)
    faces = [F.SphericalFace(owner=owner, diameter=25.4, material=m1, z_height=6.0, curvature=43.96),
             F.SphericalFace(owner=owner, diameter=25.4, material=m2, z_height=0.0, curvature=-42.90),
             F.SphericalFace(owner=owner, diameter=25.4, material=m3, z_height=-4.0,
                             curvature=-392.21)]
    fl = _facelist(core, owner, faces)
    axis = np.array(owner.direction)
    start = np.array(owner.centre) + 40.0 * axis
    wl = np.array([0.45, 0.55, 0.65, 0.8, 1.0])
    rays = disc_source(n, centre=start, axis=-axis, radius=10.0, seed=seed, n_wavelengths=len(wl),
                       E1=1.0, E2=0.5j)
    return dict(name="config2_achromat", face_lists=[fl], rays=rays, wavelengths=wl,
                max_length=200.0, recursion_limit=200)


def config3_aspheric_zernike(core, n=10000000, seed=3):
    """Config 3: AsphericFace (Newton) + DistortionFace(ShapedPlanarFace, Zernike)
    with an N-BK7 CoatedDispersiveMaterial (cf. examples/zernike_distortion_example.py,
    examples/aspheric_lens_example.py), jittered directions."""
    F, M, S, D = core.cfaces, core.cmaterials, core.cshapes, core.cdistortions
    owner = Pose(centre=(0., 0., 0.), direction=(0., 0., 1.))
    air, bk7, coat = nondispersive(core, 1.0), glass_curve(core, "N-BK7"), nondispersive(core, 1.37)
    shape = S.CircleShape(radius=10.0)
    m1 = M.CoatedDispersiveMaterial(dispersion_inside=bk7, dispersion_outside=air,
                                    dispersion_coating=coat, coating_thickness=0.1)
    m2 = M.CoatedDispersiveMaterial(dispersion_inside=air, dispersion_outside=bk7,
                                    dispersion_coating=coat, coating_thickness=0.1)
    f1 = F.AsphericFace(owner=owner, shape=shape, material=m1, z_height=8.0, curvature=30.0,
                        conic_const=-0.8, A4=1e-5, A6=-2e-8)
    base = F.ShapedPlanarFace(owner=owner, shape=shape, z_height=0.0)
    dist = D.ZernikeDistortion(unit_radius=10.0, j4=2e-3, j7=1e-3, j8=-1.5e-3, j12=5e-4)
    f2 = F.DistortionFace(owner=owner, base_face=base, distortion=dist, shape=shape, material=m2)
    fl = _facelist(core, owner, [f1, f2])
    rays = disc_source(n, centre=(0., 0., 40.), axis=(0., 0., -1.), radius=8.0, seed=seed,
                       jitter=0.01, E1=1.0, E2=0.0)
    return dict(name="config3_aspheric_zernike", face_lists=[fl], rays=rays,
                wavelengths=np.array([0.633]), max_length=200.0, recursion_limit=200)


def _extrusion_faces(core, owner, profile, z1, z2, material, trace_ends=True):
    """Extrusion.make_faces (raypier/prisms.py:81-100): circular pairwise over the
    profile, then the two PolygonFace end caps."""
    F = core.cfaces
    profile = np.asarray(profile, dtype=np.double)
    sides = []
    k = profile.shape[0]
    for i in range(k):
        (x2, y2), (x1, y1) = profile[i], profile[(i + 1) % k]
        sides.append(F.ExtrudedPlanarFace(owner=owner, z1=z1, z2=z2, x1=x1, y1=y1, x2=x2, y2=y2,
                                          material=material))
    if trace_ends:
        sides.append(F.PolygonFace(owner=owner, z_plane=z1, xy_points=profile, material=material))
        sides.append(F.PolygonFace(owner=owner, z_plane=z2, material=material, xy_points=profile,
                                   invert_normal=True))
    return sides


def michelson_face_lists(core):
    """UnpolarisingBeamsplitterCube(size=10) + two PEC mirrors (flat, R=2000 mm),
    examples/michelson_interferometer_example.py:10-40, raypier/beamsplitters.py:86-97."""
    F, M, S = core.cfaces, core.cmaterials, core.cshapes
    h = 5.0
    cube = Pose(centre=(0., 0., 0.), direction=(0., 0., 1.))
    glass = M.FullDielectricMaterial(n_inside=1.5, n_outside=1.0)
    faces = _extrusion_faces(core, cube, [(-h, -h), (-h, h), (h, h), (h, -h)], -h, h, glass)
    faces.append(F.ExtrudedPlanarFace(owner=cube, z1=-h, z2=h, x1=-h, y1=-h, x2=h, y2=h,
                                      material=M.PartiallyReflectiveMaterial(reflectivity=0.5)))
    fl_cube = _facelist(core, cube, faces)
    shape = S.CircleShape(radius=10.0)
    m1 = Pose(centre=(0., 20., 0.), direction=(0., -1., 0.))
    m2 = Pose(centre=(20., 0., 0.), direction=(-1., 0., 0.))
    fl_m1 = _facelist(core, m1, [F.ShapedPlanarFace(owner=m1, shape=shape, z_height=0.0,
                                                    material=M.PECMaterial())])
    fl_m2 = _facelist(core, m2, [F.ShapedSphericalFace(owner=m2, shape=shape, z_height=0.0,
                                                       curvature=2000.0, material=M.PECMaterial())])
    return [fl_cube, fl_m1, fl_m2]


def config5_michelson(core, n=20000, seed=5, gausslets=True, radius=3.0):
    """Config 5: Michelson interferometer with an unpolarising beamsplitter cube
    (branching ray tree), Gaussian-weighted collimated source of gausslets."""
    face_lists = michelson_face_lists(core)
    wl = np.array([1.0])
    rays = disc_source(n, centre=(-30., 0., 0.), axis=(1., 0., 0.), radius=radius, seed=seed,
                       E_vector=(0., 1., 0.), gaussian_sigma=5.0,
                       ray_type_id=GAUSSLET if gausslets else 0)
    out = dict(name="config5_michelson" + ("" if gausslets else "_rays"), face_lists=face_lists,
               wavelengths=wl, max_length=50.0, recursion_limit=200)
    if gausslets:
        rays['length'] = 50.0
        gc = core.ctracer.GaussletCollection.from_rays(rays)
        gc.config_parabasal_rays(wl, 0.5, 0.0)
        out['rays'] = gc.copy_as_array()
    else:
        out['rays'] = rays
    return out


def config_big_scene(core, n=3000, seed=55, gausslets=False, n_baffles=250):
    """The Michelson of config 5 inside a cage of `n_baffles` opaque ring-shaped stops (ShapedPlanarFace,
    Circle AND NOT Circle): > 232 faces, i.e. scene tables beyond the 40 KB the kernels stage in
    shared memory, so the trace runs the global-memory kernel instantiations (SS=false).  The stops sit
    along both arms with apertures wider than the beam; the stray reflections of the cube faces end on
    them."""
    F, M, S = core.cfaces, core.cmaterials, core.cshapes
    cfg = config5_michelson(core, n=n, seed=seed, gausslets=gausslets, radius=5.0)
    face_lists = cfg['face_lists']
    opaque = M.OpaqueMaterial()
    ring = S.CircleShape(radius=9.0) & ~S.CircleShape(radius=4.0)
    k = 0
    for arm, (axis, lo, hi) in enumerate((((0., 1., 0.), 6.0, 19.0), ((1., 0., 0.), 6.0, 19.0))):
        m = n_baffles // 2 if arm == 0 else n_baffles - n_baffles // 2
        for j in range(m):
            pos = lo + (hi - lo) * (j + 0.5) / m
            centre = tuple(pos * a for a in axis)
            owner = Pose(centre=centre, direction=axis)
            face = F.ShapedPlanarFace(owner=owner, shape=ring, z_height=0.0, material=opaque)
            face_lists.append(_facelist(core, owner, [face]))
            k += 1
    cfg['name'] = "config_big_scene" + ("" if gausslets else "_rays")
    return cfg


def icosphere(radius=5.0, subdivisions=2):
    """Triangulated sphere (outward-facing triangles): points N x 3, cells M x 3 int32."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    pts = [np.array(p, dtype=np.double) / np.sqrt(1 + t * t) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = pts[a] + pts[b]
                pts.append(m / np.sqrt((m ** 2).sum()))
                cache[key] = len(pts) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(pts) * radius, np.array(f, dtype=np.int32)


def bowl_mesh(half_width=10.0, n=24, focal=20.0):
    """Triangulated paraboloid z = r^2 / (4 f) on an n x n grid (normals towards +z)."""
    xs = np.linspace(-half_width, half_width, n)
    X, Y = np.meshgrid(xs, xs, indexing='ij')
    Z = (X ** 2 + Y ** 2) / (4.0 * focal)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    cells = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            cells += [(a, b, c), (a, c, d)]
    return pts, np.array(cells, dtype=np.int32)


def config_mesh(core, n=20000, seed=61, gausslets=False, mesh_n=40, ball_subdiv=2):
    """Triangle-mesh optics (raypier/meshes.py:16-67 builds OBBTreeFace from an STL file).
    Plain rays: a glass ball lens as an icosphere (FullDielectricMaterial: refraction, Fresnel reflections,
    TIR between facets) over a tilted faceted paraboloid mirror (PEC), both OBBTreeFace over an OBBTree.
    Gausslets: the faceted mirror under a flat PEC mirror (the beam bounces between them).  The dielectric
    ball is left out there on purpose: a parabasal ray that misses (or goes NaN by TIR) makes the REFERENCE
    read an uninitialised intersect_t.face_idx (OBBTreeFace.intersect_c, obbtree.pyx:913-932, never sets
    it), so its behaviour for such gausslets is undefined; this package drops them like every other face
    does (quirk Q16)."""
    import importlib
    OB = importlib.import_module(core.__name__ + ".obbtree")
    M, F = core.cmaterials, core.cfaces
    wl = np.array([0.633])

    def mesh_face(owner, pts, cells, material):
        tree = OB.OBBTree(np.ascontiguousarray(pts, dtype=np.double).copy(), np.ascontiguousarray(cells, dtype=np.int32).copy())
        tree.max_level = 100
        tree.number_of_cells_per_node = 6
        owner.mesh_points, owner.mesh_cells = pts, cells  # a reference OBBTree keeps its mesh private
        return OB.OBBTreeFace(tree=tree, material=material, owner=owner)

    mirror = Pose(centre=(0.0, 0.0, -25.0), direction=(0.1, -0.05, 1.0))
    pts, cells = bowl_mesh(half_width=60.0, n=mesh_n, focal=30.0)
    fl_mirror = _facelist(core, mirror, [mesh_face(mirror, pts, cells, M.PECMaterial())])
    if gausslets:
        flat = Pose(centre=(0.0, 0.0, 45.0), direction=(0.0, 0.0, 1.0), diameter=150.0, offset=0.0)
        fl_top = _facelist(core, flat, [F.CircularFace(owner=flat, diameter=150.0, material=M.PECMaterial())])
        face_lists = [fl_top, fl_mirror]
    else:
        ball = Pose(centre=(0.3, -0.2, 0.0), direction=(0.05, 0.02, 1.0))
        pts, cells = icosphere(radius=5.0, subdivisions=ball_subdiv)
        glass = M.FullDielectricMaterial(n_inside=1.5, n_outside=1.0)
        face_lists = [_facelist(core, ball, [mesh_face(ball, pts, cells, glass)]), fl_mirror]
    rays = disc_source(n, centre=(0., 0., 30.), axis=(0., 0., -1.), radius=4.0, seed=seed, E_vector=(1., 0., 0.),
                       E1=1.0, E2=0.4j, ray_type_id=GAUSSLET if gausslets else 0)
    out = dict(name="config_mesh" + ("" if gausslets else "_rays"), face_lists=face_lists, wavelengths=wl,
               max_length=120.0, recursion_limit=5)
    if gausslets:
        rays['length'] = 120.0
        gc = core.ctracer.GaussletCollection.from_rays(rays)
        gc.config_parabasal_rays(wl, 0.05, 0.0)
        out['rays'] = gc.copy_as_array()
    else:
        out['rays'] = rays
    return out


def config4_prisms(core, n=100000, seed=4):
    """Config 4 (TIR part): a rhomboid + a right-angle (Dove-like) prism built as
    extrusions with FullDielectricMaterial (examples/ctracer_demo_prisms.py), low
    thresholds so weak Fresnel branches are followed over many generations, plus an
    OpaqueMaterial beam stop (raypier/beamstop.py)."""
    F, M = core.cfaces, core.cmaterials
    p1 = Pose(centre=(0., 0., 0.), direction=(0., 0., 1.))
    glass = M.FullDielectricMaterial(n_inside=1.5, n_outside=1.0, reflection_threshold=0.02,
                                     transmission_threshold=0.02)
    rhomboid = [(-10., -5.), (-4., 5.), (10., 5.), (4., -5.)]
    fl1 = _facelist(core, p1, _extrusion_faces(core, p1, rhomboid, -6., 6., glass))
    p2 = Pose(centre=(30., 2., 0.), direction=(0., 0., 1.), rotation=10.0)
    glass2 = M.FullDielectricMaterial(n_inside=1.764, n_outside=1.0, reflection_threshold=0.02,
                                      transmission_threshold=0.02)
    tri = [(-8., -8.), (-8., 8.), (8., -8.)]
    fl2 = _facelist(core, p2, _extrusion_faces(core, p2, tri, -6., 6., glass2))
    stop = Pose(centre=(60., 0., 0.), direction=(-1., 0., 0.), diameter=40.0, offset=0.0)
    fl3 = _facelist(core, stop, [F.CircularFace(owner=stop, diameter=40.0,
                                                material=M.OpaqueMaterial())])
    rays = disc_source(n, centre=(-30., 0.3, 0.2), axis=(1., 0.02, 0.01), radius=3.0, seed=seed,
                       E1=1.0, E2=0.3j)
    return dict(name="config4_prisms", face_lists=[fl1, fl2, fl3], rays=rays,
                wavelengths=np.array([0.633]), max_length=300.0, recursion_limit=200)


def resample_relaunch(arr, max_length=80.0):
    """A deterministic stand-in for the decomposition callbacks of raypier/decompositions.py
    (PositionDecompositionPlane.evaluate_decomposed_rays): every SECOND captured gausslet is re-launched
    from the point where it met the plane, amplitudes halved, with a small deterministic tilt; pure numpy
    on a gausslet_dtype array, so the same function serves the reference (wrapped in GaussletCollection),
    the oracle and the CUDA path."""
    g = np.ascontiguousarray(arr[::2]).copy()
    b = g['base_ray']
    end = b['origin'] + b['direction'] * b['length'][:, None]
    shift = end - b['origin']
    tilt = np.array([0.0, 0.0, 0.01])
    d = b['direction'] + tilt[None, :]
    d /= np.sqrt((d * d).sum(axis=1))[:, None]
    g['base_ray']['origin'] = end
    g['base_ray']['direction'] = d
    g['base_ray']['accumulated_path'] = b['accumulated_path'] + b['length'] * b['refractive_index'].real
    g['base_ray']['E1_amp'] = b['E1_amp'] * 0.5
    g['base_ray']['E2_amp'] = b['E2_amp'] * 0.5
    g['base_ray']['parent_idx'] = np.arange(len(g), dtype=np.uint32)
    g['base_ray']['length'] = max_length
    g['para_rays']['origin'] = g['para_rays']['origin'] + shift[:, None, :]
    pd = g['para_rays']['direction'] + tilt[None, None, :]
    pd /= np.sqrt((pd * pd).sum(axis=2))[:, :, None]
    g['para_rays']['direction'] = pd
    g['para_rays']['length'] = max_length
    return g


def config_resample(core, n=2000, seed=71):
    """A decomposition plane in a beam path (raypier/decompositions.py: a CircularFace carrying a
    ResampleGaussletMaterial): gausslets pass a beamsplitter plate; the transmitted arm meets the
    decomposition plane (captured, handed to the callback, re-launched), then a curved mirror and a lens
    face; the reflected arm meets a plane mirror.  The callback's output joins generation 2 AFTER the
    regular children, and its rays are traced on like any others."""
    F, M, S = core.cfaces, core.cmaterials, core.cshapes
    plate = Pose(centre=(0., 0., 0.), direction=(1., 0., 1.), diameter=30.0, offset=0.0)
    fl_plate = _facelist(core, plate, [F.CircularFace(owner=plate, diameter=30.0,
                                                      material=M.PartiallyReflectiveMaterial(reflectivity=0.4))])
    dec = Pose(centre=(0., 0., 20.), direction=(0., 0., -1.), diameter=25.0, offset=0.0)
    mat = M.ResampleGaussletMaterial(eval_func=None)
    fl_dec = _facelist(core, dec, [F.CircularFace(owner=dec, diameter=25.0, material=mat)])
    shape = S.CircleShape(radius=12.0)
    mir = Pose(centre=(0., 0., 45.), direction=(0., 0., -1.))
    fl_mir = _facelist(core, mir, [F.ShapedSphericalFace(owner=mir, shape=shape, z_height=0.0, curvature=-400.0,
                                                         material=M.PECMaterial())])
    side = Pose(centre=(25., 0., 0.), direction=(-1., 0., 0.))
    fl_side = _facelist(core, side, [F.ShapedPlanarFace(owner=side, shape=shape, z_height=0.0,
                                                        material=M.PECMaterial())])
    wl = np.array([1.0])
    rays = disc_source(n, centre=(0., 0., -25.), axis=(0., 0., 1.), radius=3.0, seed=seed,
                       E_vector=(0., 1., 0.), gaussian_sigma=4.0, ray_type_id=GAUSSLET)
    rays['length'] = 80.0
    gc = core.ctracer.GaussletCollection.from_rays(rays)
    gc.config_parabasal_rays(wl, 0.4, 0.0)
    return dict(name="config_resample", face_lists=[fl_plate, fl_dec, fl_mir, fl_side], rays=gc.copy_as_array(),
                wavelengths=wl, max_length=80.0, recursion_limit=8, decomp_material=mat)


def config_uvpatch(core, n=4000, seed=81, gausslets=False, u_res=14, v_res=12):
    """UV patch faces (raypier/core/cbezier.pyx): a cubic-by-quadratic BezierPatch as a
    free-form mirror and a quadratic B-spline patch as a second one, a tilted collimated source (so the
    Newton iteration on the patch really iterates), and a plane absorber."""
    B, F, M = core.cbezier, core.cfaces, core.cmaterials
    import contextlib
    import io
    quiet = contextlib.redirect_stdout(io.StringIO())  # the reference's control_pts setter prints
    bez = B.BezierPatch(3, 2)
    xs, ys = np.linspace(-12., 12., 4), np.linspace(-10., 10., 3)
    X, Y = np.meshgrid(xs, ys, indexing='ij')
    Z = 0.012 * X * X - 0.02 * Y * Y + 0.4 * np.sin(X / 6.0) + 0.05 * X
    with quiet:
        bez.control_pts = np.ascontiguousarray(np.stack([X, Y, Z], axis=-1))
    own1 = Pose(centre=(0., 0., 30.), direction=(0., 0.2, 1.), u_res=u_res, v_res=v_res)
    f1 = B.UVPatchFace(owner=own1, patch=bez, u_res=u_res, v_res=v_res, material=M.PECMaterial())
    bsp = B.BSplinePatch(4, 3)
    xs, ys = np.linspace(-14., 14., 5), np.linspace(-12., 12., 4)
    X, Y = np.meshgrid(xs, ys, indexing='ij')
    Z = -0.01 * (X * X + Y * Y) + 0.3 * np.cos(Y / 5.0)
    with quiet:
        bsp.control_pts = np.ascontiguousarray(np.stack([X, Y, Z], axis=-1))
    bsp.u_degree, bsp.v_degree = 2, 2
    bsp.u_knots = np.array([0., 0., 0., 1. / 3, 2. / 3, 1., 1., 1.]) * 1.0000001  # open uniform; the last basis
    bsp.v_knots = np.array([0., 0., 0., 0.5, 1., 1., 1.]) * 1.0000001             # function is 0 AT the end knot
    own2 = Pose(centre=(0., -8., 2.), direction=(0., 0.35, 1.), u_res=u_res, v_res=v_res)
    f2 = B.UVPatchFace(owner=own2, patch=bsp, u_res=u_res, v_res=v_res, invert_normals=1, material=M.PECMaterial())
    stop = Pose(centre=(0., -20., 40.), direction=(0., 0.4, -1.), diameter=80.0, offset=0.0)
    f3 = F.CircularFace(owner=stop, diameter=80.0, material=M.OpaqueMaterial())
    fls = [_facelist(core, own1, [f1]), _facelist(core, own2, [f2]), _facelist(core, stop, [f3])]
    wl = np.array([0.8])
    rays = disc_source(n, centre=(0.5, -1., 0.), axis=(0.03, 0.05, 1.), radius=5.0, seed=seed, E_vector=(1., 0., 0.),
                       E1=1.0, E2=0.2j, ray_type_id=GAUSSLET if gausslets else 0)
    out = dict(name="config_uvpatch", face_lists=fls, wavelengths=wl, max_length=120.0, recursion_limit=6)
    if gausslets:
        rays['length'] = 120.0
        gc = core.ctracer.GaussletCollection.from_rays(rays)
        gc.config_parabasal_rays(wl, 0.3, 0.0)
        out['rays'] = gc.copy_as_array()
    else:
        out['rays'] = rays
    return out


def config4_grating(core, n=100000, seed=44, n_wavelengths=260):
    """Config 4 (grating part): the diffraction-grating dispersion compensator
    (examples/grating_dispersion_compensator_example.py:19-50): RectangularGrating
    1400 l/mm order -1, an achromat, a PEC mirror and a beam stop; broadband source
    with 260 wavelengths 0.76-0.80 um."""
    F, M = core.cfaces, core.cmaterials
    wl = np.linspace(0.76, 0.80, n_wavelengths)
    ang = math.radians(41.0)
    g = Pose(centre=(0., 0., 0.), direction=(-math.cos(ang), math.sin(ang), 0.0), rotation=180.0,
             length=25.0, width=25.0, offset=0.0)
    gmat = M.DiffractionGratingMaterial(lines_per_mm=1400.0, order=-1, efficiency=0.9,
                                        origin=(0., 0., 0.))
    fl_g = _facelist(core, g, [F.RectangularFace(owner=g, length=25.0, width=25.0, offset=0.0,
                                                 material=gmat)])
    mir = Pose(centre=(-57.9, 15.7, 0.), direction=(0.965, -0.261, 0.0), diameter=25.0, offset=0.0)
    fl_m = _facelist(core, mir, [F.CircularFace(owner=mir, diameter=25.0,
                                                material=M.PECMaterial())])
    stop = Pose(centre=(-120., 0., 0.), direction=(1., 0., 0.), diameter=200.0, offset=0.0)
    fl_s = _facelist(core, stop, [F.CircularFace(owner=stop, diameter=200.0,
                                                 material=M.OpaqueMaterial())])
    rays = disc_source(n, centre=(-80., 0., 0.), axis=(1., 0., 0.), radius=2.0, seed=seed,
                       n_wavelengths=n_wavelengths, E_vector=(0., 0., 1.))
    return dict(name="config4_grating", face_lists=[fl_g, fl_m, fl_s], rays=rays, wavelengths=wl,
                max_length=300.0, recursion_limit=200)


def _cpc_profile(n_seg=6, a_in=6.0, a_out=14.0, length=30.0):
    """A smooth trough wall from (a_in, 0) to (a_out, length) as a chain of cubic Bezier
    segments (the shape family of examples/CPC3.py / dielectrictroughs.py)."""
    ts = np.linspace(0.0, 1.0, n_seg + 1)

    def pt(t):
        return np.array([a_in + (a_out - a_in) * t ** 1.6, length * t])

    def dpt(t):
        return np.array([(a_out - a_in) * 1.6 * max(t, 1e-9) ** 0.6, length])

    segs = []
    for t0, t1 in zip(ts[:-1], ts[1:]):
        h = (t1 - t0) / 3.0
        segs.append([pt(t0), pt(t0) + h * dpt(t0), pt(t1) - h * dpt(t1), pt(t1)])
    return np.array(segs)


def config4_cpc(core, n=100000, seed=45):
    """Config 4 (CPC / trough part): two mirrored extruded-Bezier walls (PEC) forming a
    compound-parabolic-like trough (raypier.splines.Extruded_bezier, examples/CPC3.py), rays
    entering the wide aperture at a spread of angles, a detector (OpaqueMaterial) at the throat."""
    F, M = core.cfaces, core.cmaterials
    right = _cpc_profile()
    left = right.copy()
    left[:, :, 0] *= -1.0
    walls = Pose(centre=(0., 0., 0.), direction=(0., 0., 1.))
    walls.control_points, walls.z_height_1, walls.z_height_2 = right, -20.0, 20.0
    walls2 = Pose(centre=(0., 0., 0.), direction=(0., 0., 1.))
    walls2.control_points, walls2.z_height_1, walls2.z_height_2 = left, -20.0, 20.0
    pec = M.PECMaterial()
    f_r = F.ExtrudedBezierFace(owner=walls, beziercurves=right, z_height_1=-20.0, z_height_2=20.0, material=pec)
    f_l = F.ExtrudedBezierFace(owner=walls2, beziercurves=left, z_height_1=-20.0, z_height_2=20.0, material=pec)
    fl_r = _facelist(core, walls, [f_r])
    fl_l = _facelist(core, walls2, [f_l])
    det = Pose(centre=(0., -0.5, 0.), direction=(0., 1., 0.), length=14.0, width=40.0, offset=0.0)
    fl_d = _facelist(core, det, [F.RectangularFace(owner=det, length=14.0, width=40.0, offset=0.0,
                                                   material=M.OpaqueMaterial())])
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=ray_dtype)
    rays['origin'] = np.stack([rng.uniform(-13.0, 13.0, n), np.full(n, 35.0), rng.uniform(-15.0, 15.0, n)], axis=1)
    ang = rng.uniform(-0.35, 0.35, n)
    tilt = rng.normal(0.0, 0.05, n)
    d = np.stack([np.sin(ang), -np.cos(ang), tilt], axis=1)
    rays['direction'] = d / np.linalg.norm(d, axis=1)[:, None]
    rays['E_vector'] = (0.0, 0.0, 1.0)
    rays['E1_amp'] = 1.0
    rays['E2_amp'] = 0.4j
    rays['refractive_index'] = 1.0
    rays['normal'] = (0.0, 1.0, 0.0)
    rays['length'] = np.inf
    rays['ray_ident'] = np.arange(n, dtype=np.uint32)
    return dict(name="config4_cpc", face_lists=[fl_r, fl_l, fl_d], rays=rays,
                wavelengths=np.array([0.633]), max_length=120.0, recursion_limit=30)


def config_zoo(core, n=20000, seed=17, gausslets=False):
    """Parity "zoo": one station per face type / material class that the BASELINE configs do not
    reach -- a row of independent optics 30 mm apart along x, lit from z = -50 by a wide strip of
    jittered rays with three wavelengths, closed by two opaque stops.  Every face must be hit
    (tests assert Face.count > 0), so the faces ElipticalPlane, ImplicitBoundedPlanar,
    OrientedPolygon, OffAxisParabolic, Ellipsoidal, Saddle, Cylinderical, Axicon, ConicRevolution,
    ExtendedPolynomial, ShapedSpherical, ShapedPlanar (with AND / OR / XOR / Invert / Polygon
    shapes, Plane / Sphere / Cylinder / Difference / Intersection implicit surfaces) and the
    materials Transparent, LinearPolarising, Waveplate, Dielectric, absorbing FullDielectric and
    SingleLayerCoated (complex n: the general Fresnel / thin-film path), absorbing
    FullDielectricDispersive, Circular / Rectangular aperture, PartiallyReflective and PEC are all
    pinned against the reference."""
    F, M, S, I = core.cfaces, core.cmaterials, core.cshapes, core.cimplicit_surfs
    T = core.ctracer.Transform
    pitch = 30.0
    wl = np.array([0.5, 0.633, 0.85])
    lists = []

    def station(i, faces_fn, dx=0.0, **attrs):
        owner = Pose(centre=(pitch * i + dx, 0., 0.), direction=(0., 0., 1.), **attrs)
        lists.append(_facelist(core, owner, faces_fn(owner, (pitch * i, 0.0, 0.0))))

    circ8 = S.CircleShape(radius=8.0)
    station(0, lambda o, c: [F.ElipticalPlaneFace(owner=o, g_x=0.3, g_y=-0.2, diameter=16.0,
                                                   material=M.PECMaterial())], diameter=16.0)
    # inside the sphere, on one side of a plane, outside a tilted cylinder -- or in a small blob
    # (Difference is the reference's arithmetic out -= v, cimplicit_surfs.pyx:191-194, so it is
    # given a plane whose value stays small on the target plane)
    boundary = I.Union(I.Intersection(I.Sphere(centre=(0., 0., 0.), radius=9.0),
                                      I.Plane(origin=(6., 0., 0.), normal=(1., 0.2, 0.)),
                                      I.Invert(I.Cylinder(origin=(-2., 1., 0.), axis=(0., 0.1, 1.), radius=2.5))),
                       I.Difference(I.Sphere(centre=(7., 6., 1.), radius=2.0),
                                    I.Plane(origin=(7., 6., 1.), normal=(0., 0., 1.))))
    station(1, lambda o, c: [F.ImplicitBoundedPlanarFace(owner=o, target=I.Plane(origin=(0., 0., 1.), normal=(0.1, 0.05, 1.)),
                                                         boundary=boundary, material=M.TransparentMaterial())])
    hexagon = [(8 * math.cos(k * math.pi / 3), 8 * math.sin(k * math.pi / 3)) for k in range(6)]
    station(2, lambda o, c: [F.OrientedPolygonFace(owner=o, origin=(0., 0., 0.5), normal=(0.2, 0.1, 1.0),
                                                   x_axis=(1.0, 0.0, -0.2), xy_points=hexagon,
                                                   material=M.LinearPolarisingMaterial())])
    # off-axis paraboloid: the aperture is centred on local x = EFL, so the list sits EFL to the left
    def oap(o, c):
        f = F.OffAxisParabolicFace(owner=o, material=M.PECMaterial())
        f.EFL, f.diameter, f.height = 40.0, 16.0, 25.0  # plain attributes in the reference (parabolics.py:43-47)
        return [f]
    station(3, oap, dx=-40.0)
    ell = np.eye(4)
    ell[:3, 3] = (0.0, 0.0, -20.0)  # local (0,0,0) lies on the ellipsoid (ellipse frame z = -minor)

    def ellipsoid(o, c):
        f = F.EllipsoidalFace(owner=o, material=M.PECMaterial())
        return [f]
    station(4, ellipsoid, ellipse_trans=_VtkLikeTransform(ell), axes=(30.0, 20.0), X_bounds=(-8.0, 8.0),
            Y_bounds=(-8.0, 8.0), Z_bounds=(-5.0, 5.0))
    station(5, lambda o, c: [F.SaddleFace(owner=o, shape=circ8, z_height=0.0, curvature=0.02,
                                          material=M.DielectricMaterial(n_inside=1.5, n_outside=1.0))])
    station(6, lambda o, c: [F.CylindericalFace(owner=o, shape=S.RectangleShape(centre=(0., 0.), width=14.0, height=12.0),
                                                z_height=1.0, radius=40.0,
                                                material=M.FullDielectricMaterial(n_inside=1.6 + 0.02j, n_outside=1.0,
                                                                                  reflection_threshold=0.001,
                                                                                  transmission_threshold=0.001))])
    station(7, lambda o, c: [F.AxiconFace(owner=o, shape=circ8, z_height=0.0, gradient=0.08,
                                          material=M.SingleLayerCoatedMaterial(n_inside=1.5 + 0.01j, n_outside=1.0,
                                                                               n_coating=1.3 + 0.02j, thickness=0.12,
                                                                               reflection_threshold=0.001,
                                                                               transmission_threshold=0.001))])
    r2 = math.sqrt(0.5)
    station(8, lambda o, c: [F.ConicRevolutionFace(owner=o, shape=circ8, z_height=2.0, conic_const=-1.4, curvature=35.0,
                                                   material=M.WaveplateMaterial(retardance=0.25, fast_axis=(r2, r2, 0.0)))])
    coefs = np.array([[0.0, 0.0, 1e-2], [0.0, 2e-2, 0.0], [1e-2, 0.0, 0.0]])
    station(9, lambda o, c: [F.ExtendedPolynomialFace(owner=o, shape=circ8, z_height=1.0, conic_const=-0.5, curvature=40.0,
                                                      norm_radius=8.0, coefs=coefs,
                                                      material=M.CircularApertureMaterial(outer_radius=8.0, radius=5.0,
                                                                                          edge_width=1.0, origin=c))])
    xor = S.BooleanXOR(circ8, S.RectangleShape(centre=(2.0, 0.0), width=6.0, height=4.0))
    station(10, lambda o, c: [F.ShapedSphericalFace(owner=o, shape=xor, z_height=3.0, curvature=60.0,
                                                    material=M.RectangularApertureMaterial(outer_width=16.0, outer_height=16.0,
                                                                                           width=6.0, height=8.0, edge_width=1.5,
                                                                                           origin=c))])
    pentagon = [(8 * math.cos(0.3 + k * 2 * math.pi / 5), 8 * math.sin(0.3 + k * 2 * math.pi / 5)) for k in range(5)]
    poly = S.PolygonShape()
    poly.coordinates = np.ascontiguousarray(pentagon, dtype=np.double)  # the reference has no constructor keyword
    shp = S.BooleanAND(S.InvertShape(S.CircleShape(centre=(0.0, 0.0), radius=2.0)),
                       S.BooleanOR(poly, S.CircleShape(centre=(6.0, 0.0), radius=3.0)))
    station(11, lambda o, c: [F.ShapedPlanarFace(owner=o, shape=shp, z_height=0.0,
                                                 material=M.PartiallyReflectiveMaterial(reflectivity=0.3))])
    absorbing = glass_curve(core, "N-SF11", absorption=5.0)
    station(12, lambda o, c: [F.CircularFace(owner=o, z_plane=0.0,
                                             material=M.FullDielectricDispersiveMaterial(
                                                 dispersion_inside=absorbing, dispersion_outside=nondispersive(core, 1.0),
                                                 reflection_threshold=0.001, transmission_threshold=0.001))],
            diameter=16.0, offset=0.0)
    n_st = 13
    span = pitch * (n_st - 1)
    stop_owner_back = Pose(centre=(span / 2, 0., 60.), direction=(0., 0., 1.), length=span + 80.0, width=80.0, offset=0.0)
    stop_owner_front = Pose(centre=(span / 2, 0., -70.), direction=(0., 0., 1.), length=span + 80.0, width=80.0,
                            offset=0.0)
    for o in (stop_owner_back, stop_owner_front):
        lists.append(_facelist(core, o, [F.RectangularFace(owner=o, length=span + 80.0, width=80.0, offset=0.0,
                                                           z_plane=0.0, material=M.OpaqueMaterial())]))
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=ray_dtype)
    rays['origin'][:, 0] = rng.uniform(-10.0, span + 10.0, n)
    rays['origin'][:, 1] = rng.uniform(-10.0, 10.0, n)
    rays['origin'][:, 2] = -50.0
    d = np.array([0.0, 0.0, 1.0])[None, :] + rng.normal(0.0, 0.01, size=(n, 3))
    rays['direction'] = d / np.linalg.norm(d, axis=1)[:, None]
    rays['E_vector'] = (1.0, 0.0, 0.0)
    rays['E1_amp'] = 1.0
    rays['E2_amp'] = 0.3 + 0.2j
    rays['refractive_index'] = 1.0
    rays['normal'] = (0.0, 1.0, 0.0)
    rays['length'] = np.inf
    rays['wavelength_idx'] = rng.integers(0, len(wl), size=n, dtype=np.uint32)
    rays['ray_ident'] = np.arange(n, dtype=np.uint32)
    if gausslets:
        rays['ray_type_id'] = GAUSSLET
        rays['length'] = 200.0
        gc = core.ctracer.GaussletCollection.from_rays(rays)
        gc.config_parabasal_rays(wl, 0.05, 0.0)
        rays = gc.copy_as_array()
    return dict(name="config_zoo", face_lists=lists, rays=rays, wavelengths=wl, max_length=200.0,
                recursion_limit=8)



CONFIGS = {
    "zoo": config_zoo,
    "config1": config1_singlet,
    "config2": config2_achromat,
    "config3": config3_aspheric_zernike,
    "config4_prisms": config4_prisms,
    "config4_grating": config4_grating,
    "config4_cpc": config4_cpc,
    "config5": config5_michelson,
    "big_scene": config_big_scene,
    "mesh": config_mesh,
    "resample": config_resample,
    "uvpatch": config_uvpatch,
}


def build(core, name, **kw):
    return CONFIGS[name](core, **kw)
