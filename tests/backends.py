"""Two backends with one interface for the function-level known-answer tests: the plain-C
oracle (CPU, ``-m "not gpu"``) and the CUDA unit entry points of the C ABI (``-m gpu``)."""
import numpy as np

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import scene as SC


def ray_record(**kw):
    r = np.zeros(1, dtype=A.ray_dtype)
    r['length'] = np.inf
    r['refractive_index'] = 1.0
    for k, v in kw.items():
        r[k] = v
    return r


def ray_power(rec):
    n = rec['refractive_index'].real
    return float((abs(rec['E1_amp']) ** 2 + abs(rec['E2_amp']) ** 2) * n)


def unit_scene(core, faces, wavelengths=(1.0,), owner=None):
    fl = core.ctracer.FaceList(owner=owner)
    fl.faces = list(faces)
    if owner is not None:
        fl.sync_transforms()
    for f in faces:
        f.material.wavelengths = np.asarray(wavelengths, dtype=np.double)
    return SC.Scene([fl], np.asarray(wavelengths, dtype=np.double))


class OracleBackend(object):
    name = "oracle"

    def __init__(self):
        from oracle import oracle as O
        self.O = O

    def face_intersect(self, scene, face, p1, p2, is_base_ray=1):
        return self.O.face_intersect(scene, face, p1, p2, is_base_ray)

    def orientation(self, scene, face, point):
        return self.O.orientation(scene, face, point)

    def material_eval(self, scene, mat, ray, idx, point, normal, tangent=(1.0, 0.0, 0.0)):
        return self.O.material_eval(scene, mat, ray, idx, point, normal, tangent)

    def distortion(self, scene, dist, x, y):
        return self.O.distortion_z(scene, dist, x, y), self.O.distortion_zgrad(scene, dist, x, y)

    def trace(self, scene, rays, max_length, recursion_limit):
        gens, counts = self.O.trace_rays(scene, rays, recursion_limit, max_length)
        return gens, counts


class CudaBackend(object):
    name = "cuda"

    def __init__(self, engine):
        self.e = engine

    def face_intersect(self, scene, face, p1, p2, is_base_ray=1):
        self.e.set_scene(scene)
        return float(self.e.unit_face_intersect(face, p1, p2, is_base_ray)[0])

    def orientation(self, scene, face, point):
        self.e.set_scene(scene)
        n, t = self.e.unit_face_normal(face, point)
        return n[0], t[0]

    def material_eval(self, scene, mat, ray, idx, point, normal, tangent=(1.0, 0.0, 0.0)):
        self.e.set_scene(scene)
        out, cnt = self.e.unit_material_eval(mat, ray, point, normal, tangent)
        kids = out[0, :int(cnt[0])].copy()
        kids['parent_idx'] = idx  # the unit kernel numbers parents 0..n-1
        return kids

    def distortion(self, scene, dist, x, y):
        self.e.set_scene(scene)
        z, g = self.e.unit_distortion(dist, [x], [y])
        return float(z[0]), g[0]

    def trace(self, scene, rays, max_length, recursion_limit):
        self.e.set_scene(scene)
        res = self.e.trace(rays, max_length, recursion_limit)
        gens = res.generations()
        counts = res.face_counts.copy()
        res.free()
        return gens, counts
