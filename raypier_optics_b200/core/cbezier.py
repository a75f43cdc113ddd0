"""Host-side mirror of ``raypier.core.cbezier`` (raypier/core/cbezier.pyx): Bezier / B-spline patches and
the face traced through them.

``BezierPatch(N, M)`` / ``BSplinePatch(N, M)`` hold (N+1) x (M+1) control points; ``UVPatchFace(patch=...)``
tessellates the patch (``get_mesh``, cbezier.pyx:153-197), finds the nearest facet along a ray and polishes
the hit with a Newton iteration on the patch itself (:459-528).  The arithmetic of the evaluation loops is
kept scalar and in the reference's order, so a mirror patch tessellates to the same mesh as the reference's.
"""
import math

import numpy as np

from .ctracer import Face


def _n_basis(t, p, idx, knots):
    """cbezier.pyx:47-70"""
    if p == 0:
        return 1.0 if (knots[idx] <= t) and (t < knots[idx + 1]) else 0.0
    denom = knots[idx + p] - knots[idx]
    out = 0.0 if denom == 0.0 else ((t - knots[idx]) / denom) * _n_basis(t, p - 1, idx, knots)
    denom = knots[idx + p + 1] - knots[idx + 1]
    if denom != 0.0:
        out += ((knots[idx + p + 1] - t) / denom) * _n_basis(t, p - 1, idx + 1, knots)
    return out


def N_basis(t, p, idx, knots):
    return _n_basis(float(t), int(p), int(idx), [float(k) for k in knots])


class BaseUVPatch(object):
    def _eval_pt(self, u, v):
        raise NotImplementedError

    def eval_pt(self, u, v):
        return tuple(self._eval_pt(float(u), float(v)))

    def get_mesh(self, N, M, u_range=1., v_range=1.):
        """cbezier.pyx:153-197 -> (points (N*M, 3), cells ((N-1)*(M-1)*2, 3), uv (N*M, 2))"""
        du, dv = u_range / (N - 1), v_range / (M - 1)
        points = np.empty((N, M, 3), dtype=np.float64)
        uv = np.empty((N, M, 2), dtype=np.float64)
        for i in range(N):
            for j in range(M):
                u, v = i * du, j * dv
                uv[i, j] = (u, v)
                points[i, j] = self._eval_pt(u, v)
        pt_ids = np.arange(N * M).reshape(N, M)
        cells = np.empty(((N - 1) * (M - 1) * 2, 3), np.int64)
        ct = 0
        for i in range(N - 1):
            for j in range(M - 1):
                cells[ct] = (pt_ids[i, j], pt_ids[i, j + 1], pt_ids[i + 1, j + 1])
                ct += 1
                cells[ct] = (pt_ids[i, j], pt_ids[i + 1, j + 1], pt_ids[i + 1, j])
                ct += 1
        return points.reshape(-1, 3), cells, uv.reshape(-1, 2)


def _pow(x, n):
    # C pow(double, (double)int), what Cython emits for `u**i` with a C int exponent
    try:
        return math.pow(x, float(n))
    except (ValueError, ZeroDivisionError, OverflowError):
        if x == 0.0 and n < 0:
            return math.inf
        raise


class BezierPatch(BaseUVPatch):
    """cbezier.pyx:200-286"""

    def __init__(self, N, M):
        self.order_n, self.order_m = int(N), int(M)
        self._control_pts = np.zeros((N + 1, M + 1, 3))
        f = math.factorial
        self.binom_n = np.array([f(N) / (f(i) * f(N - i)) for i in range(N + 1)])
        self.binom_m = np.array([f(M) / (f(i) * f(M - i)) for i in range(M + 1)])

    @property
    def control_pts(self):
        return self._control_pts

    @control_pts.setter
    def control_pts(self, pts_in):
        pts = np.ascontiguousarray(pts_in, dtype=np.float64)
        if pts.shape != (self.order_n + 1, self.order_m + 1, 3):
            raise ValueError("Control points array must have shape (%d,%d,3). Got %r"
                             % (self.order_n + 1, self.order_m + 1, pts.shape))
        self._control_pts = pts

    def _eval_pt(self, u, v):
        N, M, c = self.order_n, self.order_m, self._control_pts
        x = y = z = 0.0
        for i in range(N + 1):
            for j in range(M + 1):
                coef = (float(self.binom_n[i]) * _pow(u, i) * _pow(1 - u, N - i) * float(self.binom_m[j]) * _pow(v, j)
                        * _pow(1 - v, M - j))
                x += coef * float(c[i, j, 0])
                y += coef * float(c[i, j, 1])
                z += coef * float(c[i, j, 2])
        return (x, y, z)


class BSplinePatch(BaseUVPatch):
    """cbezier.pyx:290-388: (n+1) control points and degree p need (p+n+2) knots per direction."""

    def __init__(self, N, M):
        self.order_n, self.order_m = int(N), int(M)
        self._control_pts = np.zeros((N + 1, M + 1, 3))
        self.u_knots = np.zeros(0)
        self.v_knots = np.zeros(0)
        self.u_degree = 0
        self.v_degree = 0

    control_pts = BezierPatch.control_pts

    def _eval_pt(self, u, v):
        c = self._control_pts
        uk, vk = [float(k) for k in self.u_knots], [float(k) for k in self.v_knots]
        x = y = z = 0.0
        for i in range(self.order_n + 1):
            for j in range(self.order_m + 1):
                coef1 = _n_basis(u, self.u_degree, i, uk)
                coef2 = _n_basis(v, self.v_degree, j, vk)
                x += coef1 * coef2 * float(c[i, j, 0])
                y += coef1 * coef2 * float(c[i, j, 1])
                z += coef1 * coef2 * float(c[i, j, 2])
        return (x, y, z)


class UVPatchFace(Face):
    """cbezier.pyx:391-550"""

    def __init__(self, **kwds):
        own = {k: kwds.pop(k) for k in ("u_res", "v_res", "atol", "invert_normals", "patch", "max_level", "cells_per_node")
               if k in kwds}
        Face.__init__(self, **kwds)
        self.name = "UV patch face"
        self.u_res = int(own.get("u_res", 20))
        self.v_res = int(own.get("v_res", 20))
        self.atol = float(own.get("atol", 1e-10))
        self.invert_normals = 1 if own.get("invert_normals", 0) else 0
        self.patch = own["patch"]
        from .obbtree import OBBTree
        pts, cells, uvs = self.patch.get_mesh(self.u_res, self.v_res)
        self.obbtree = OBBTree(pts.copy(), np.ascontiguousarray(cells, dtype=np.int32))
        self.obbtree.max_level = own.get("max_level", 100)
        self.obbtree.number_of_cells_per_node = own.get("cells_per_node", 2)
        self.obbtree.build_tree()
        self.uvs = uvs
