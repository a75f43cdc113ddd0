"""Host-side mirror of ``raypier.core.obbtree`` (raypier/core/obbtree.pyx): triangle-mesh optics.

``OBBTree(points, cells)`` holds the mesh (points N x 3 float64, cells M x 3 int32 -- always
triangles, obbtree.pyx:204-223) and ``OBBTreeFace(tree=..., material=...)`` is the Face traced through it
(obbtree.pyx:880-946).  In the reference the oriented-bounding-box tree is an acceleration structure for
``intersect_with_line_c`` (:367-400): its node test only prunes, the result is the nearest triangle with
``tolerance / |p2 - p1| <= alpha < 1``.  The device path carries its own BVH (axis-aligned boxes, binned
surface-area-heuristic splits, built in :func:`build_bvh` when the scene is flattened), so ``build_tree`` here has nothing to
compute; ``max_level`` / ``number_of_cells_per_node`` are kept for interface compatibility.
"""
import os

import numpy as np

from .ctracer import Face

# triangles per BVH leaf.  Measured on B200 (71k-facet scene, 1e6 random rays): 1 -> 5.1e8, 2 -> 4.3e8, 4 -> 3.9e8,
# 8 -> 3.0e8 seg/s -- the triangle test is fp64, the box test fp32.  RPX_BVH_LEAF: developer override (1..8, what the
# packed device nodes hold)
LEAF_CELLS = min(8, max(1, int(os.environ.get("RPX_BVH_LEAF", "1"))))
# split rule of build_bvh: "1" = binned surface-area heuristic on the widest centroid axis (16 bins, both sides >= 25 %
# of the node, the first 20 levels), "0" = median of the centroids.  The tree only prunes, so the hits are the same.
BVH_SAH = os.environ.get("RPX_BVH_SAH", "1") == "1"
SAH_BINS, SAH_LEVELS, SAH_MIN_FRACTION = 16, 20, 0.25


class OBBTree(object):
    def __init__(self, points, cells):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 3)
        if self.cells.size and (self.cells.min() < 0 or self.cells.max() >= len(self.points)):
            raise IndexError("cell refers to a missing point")
        self.level = 0
        self.max_level = 12               # obbtree.pyx:214-215
        self.number_of_cells_per_node = 1
        self.tolerance = 0.1              # :223

    def clear_tree(self):
        self.level = 0

    def build_tree(self):
        """obbtree.pyx:259-269.  The device BVH is built at flattening time; only ``level`` (> 0 = built)
        is kept, as the depth a balanced split of this mesh reaches."""
        n = max(len(self.cells), 1)
        self.level = max(1, int(np.ceil(np.log2(max(n / float(LEAF_CELLS), 1.0)))) + 1)


class OBBTreeFace(Face):
    """obbtree.pyx:880-946: ``intersect_c`` = nearest triangle along the ray, ``compute_normal_c`` = the
    flat normal of the triangle that was hit (``intersect_t.piece_idx``)."""

    def __init__(self, **kwds):
        tree = kwds['tree']
        Face.__init__(self, **{k: v for k, v in kwds.items() if k != 'tree'})
        self.name = "OBBTree face"
        self.obbtree = tree
        if tree.level <= 0:
            tree.build_tree()


def triangle_records(points, cells):
    """Per cell: p1, v1 = p2 - p1, v2 = p3 - p1, n = v1 x v2 (the quantities line_intersects_cell_c,
    obbtree.pyx:310-343, derives on every call), with cross_ written out as ctracer's (a.y*b.z - a.z*b.y, ...)."""
    p1 = points[cells[:, 0]]
    v1 = points[cells[:, 1]] - p1
    v2 = points[cells[:, 2]] - p1
    n = np.stack([v1[:, 1] * v2[:, 2] - v1[:, 2] * v2[:, 1],
                  v1[:, 2] * v2[:, 0] - v1[:, 0] * v2[:, 2],
                  v1[:, 0] * v2[:, 1] - v1[:, 1] * v2[:, 0]], axis=1)
    return p1, v1, v2, n


def build_bvh(points, cells, leaf_cells=LEAF_CELLS, sah=None):
    """Binary BVH over the triangles: axis-aligned boxes, the centroids sorted along the widest axis and split at
    the median -- or (``sah``, default ``BVH_SAH``) where a binned surface-area heuristic puts the cut: cost =
    area(left box) * n_left + area(right box) * n_right over the 15 boundaries of 16 equal bins, both sides
    keeping >= 25 % of the triangles so that the depth stays within what the packed device format holds, for the
    first 20 levels (median below).  On the bench meshes the device walk visits 9.6 % fewer inner nodes with it
    (CPU emulation of the ordered walk: 22.8 -> 20.6 per ray and face).  Leaves of <= ``leaf_cells`` triangles.  Returns (order, nodes): ``order`` = cell ids in leaf
    order, ``nodes`` (K x 8 float64) = box min, box max, then (left, right) for an inner node (children
    always have larger ids than their parent) or (-(first) - 1, count) for a leaf, indexing ``order``.
    Boxes are padded by 1e-9 of the mesh size so the slab test on the device stays conservative.

    Built level by level with whole-array numpy operations (one segmented sort per level; nodes numbered
    breadth first), so that single-triangle leaves -- what the device walk is fastest with, its triangle test
    being fp64 and its box test fp32 -- stay affordable: 580k triangles in ~3 s."""
    tri = points[cells]                      # M x 3 x 3
    lo, hi = tri.min(axis=1), tri.max(axis=1)
    cen = tri.mean(axis=1)
    pad = 1e-9 * max(float((points.max(axis=0) - points.min(axis=0)).max()), 1e-300)
    M = len(cells)
    order = np.arange(M, dtype=np.int64)
    first = np.array([0], dtype=np.int64)
    count = np.array([M], dtype=np.int64)
    levels = []
    next_id = 1
    depth = 0
    if sah is None:
        sah = BVH_SAH

    def area(l, h):
        d = np.maximum(h - l, 0.0)
        return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]

    def seg_reduce(ufunc, arr, first, count):
        """ufunc-reduce arr[first_i : first_i + count_i] for every (disjoint, ascending) segment"""
        idx = np.column_stack([first, first + count]).ravel()
        if idx[-1] >= len(arr):
            idx = idx[:-1]
        return ufunc.reduceat(arr, idx, axis=0)[::2]

    while len(first):
        lo_o, hi_o = lo[order], hi[order]
        rec = np.zeros((len(first), 8))
        rec[:, 0:3] = seg_reduce(np.minimum, lo_o, first, count) - pad
        rec[:, 3:6] = seg_reduce(np.maximum, hi_o, first, count) + pad
        split = count > leaf_cells
        rec[~split, 6] = -first[~split] - 1
        rec[~split, 7] = count[~split]
        sf, sc = first[split], count[split]
        ns = len(sf)
        rec[split, 6] = next_id + 2 * np.arange(ns)
        rec[split, 7] = next_id + 2 * np.arange(ns) + 1
        levels.append(rec)
        if ns == 0:
            break
        cen_o = cen[order]
        cmin = seg_reduce(np.minimum, cen_o, sf, sc)
        ext = seg_reduce(np.maximum, cen_o, sf, sc) - cmin
        axis = np.argmax(ext, axis=1)
        seg = np.repeat(np.arange(ns), sc)                                   # segment of every position
        pos = np.repeat(sf - np.concatenate([[0], np.cumsum(sc)[:-1]]), sc) + np.arange(int(sc.sum()))
        key = cen_o[pos, axis[seg]]
        perm = np.lexsort((key, seg))                                        # by segment, then by the split coordinate
        order[pos] = order[pos][perm]
        n_left = sc // 2
        if sah and depth < SAH_LEVELS:
            # the sorted positions of a segment fall into SAH_BINS equal bins of its centroid range: per (segment, bin)
            # count and box by one reduceat, prefix / suffix boxes over the bins, cheapest admissible boundary
            rows = np.arange(ns)
            ext_a, cmin_a = ext[rows, axis], cmin[rows, axis]
            scale = np.where(ext_a > 0, SAH_BINS / np.where(ext_a > 0, ext_a, 1.0), 0.0)
            b = np.minimum(((key[perm] - cmin_a[seg]) * scale[seg]).astype(np.int64), SAH_BINS - 1)
            gid = seg * SAH_BINS + b                                         # ascending along the sorted positions
            cnt = np.bincount(gid, minlength=ns * SAH_BINS).reshape(ns, SAH_BINS)
            lo_s, hi_s = lo[order[pos]], hi[order[pos]]
            starts = np.flatnonzero(np.concatenate([[True], gid[1:] != gid[:-1]]))
            blo = np.full((ns * SAH_BINS, 3), np.inf)
            bhi = np.full((ns * SAH_BINS, 3), -np.inf)
            blo[gid[starts]] = np.minimum.reduceat(lo_s, starts, axis=0)
            bhi[gid[starts]] = np.maximum.reduceat(hi_s, starts, axis=0)
            blo, bhi = blo.reshape(ns, SAH_BINS, 3), bhi.reshape(ns, SAH_BINS, 3)
            plo, phi = np.minimum.accumulate(blo, axis=1), np.maximum.accumulate(bhi, axis=1)
            slo = np.minimum.accumulate(blo[:, ::-1], axis=1)[:, ::-1]
            shi = np.maximum.accumulate(bhi[:, ::-1], axis=1)[:, ::-1]
            nl = np.cumsum(cnt, axis=1)[:, :-1]
            nr = sc[:, None] - nl
            with np.errstate(invalid='ignore'):
                cost = area(plo[:, :-1], phi[:, :-1]) * nl + area(slo[:, 1:], shi[:, 1:]) * nr
            lim = np.maximum((sc * SAH_MIN_FRACTION).astype(np.int64), 1)[:, None]
            cost = np.where((nl >= lim) & (nr >= lim), cost, np.inf)
            kbest = np.argmin(cost, axis=1)
            n_left = np.where(np.isfinite(cost[rows, kbest]), nl[rows, kbest], n_left)
        first = np.column_stack([sf, sf + n_left]).ravel()
        count = np.column_stack([n_left, sc - n_left]).ravel()
        next_id += 2 * ns
        depth += 1
    return order, np.concatenate(levels)
