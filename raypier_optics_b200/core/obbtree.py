"""Host-side mirror of ``raypier.core.obbtree`` (raypier/core/obbtree.pyx): triangle-mesh optics.

``OBBTree(points, cells)`` holds the mesh (points N x 3 float64, cells M x 3 int32 -- always
triangles, obbtree.pyx:204-223) and ``OBBTreeFace(tree=..., material=...)`` is the Face traced through it
(obbtree.pyx:880-946).  In the reference the oriented-bounding-box tree is an acceleration structure for
``intersect_with_line_c`` (:367-400): its node test only prunes, the result is the nearest triangle with
``tolerance / |p2 - p1| <= alpha < 1``.  The device path carries its own BVH (axis-aligned boxes, median
split, built in :func:`build_bvh` when the scene is flattened), so ``build_tree`` here has nothing to
compute; ``max_level`` / ``number_of_cells_per_node`` are kept for interface compatibility.
"""
import os

import numpy as np

from .ctracer import Face

# triangles per BVH leaf.  Measured on B200 (71k-facet scene, 1e6 random rays): 1 -> 5.1e8, 2 -> 4.3e8, 4 -> 3.9e8,
# 8 -> 3.0e8 seg/s -- the triangle test is fp64, the box test fp32.  RPX_BVH_LEAF: developer override (1..8, what the
# packed device nodes hold)
LEAF_CELLS = min(8, max(1, int(os.environ.get("RPX_BVH_LEAF", "1"))))


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
        is kept, as the depth a median split of this mesh reaches."""
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


def build_bvh(points, cells, leaf_cells=LEAF_CELLS):
    """Binary BVH over the triangles: axis-aligned boxes, median split of the centroids along the widest
    axis, leaves of <= ``leaf_cells`` triangles.  Returns (order, nodes): ``order`` = cell ids in leaf
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
        ext = seg_reduce(np.maximum, cen_o, sf, sc) - seg_reduce(np.minimum, cen_o, sf, sc)
        axis = np.argmax(ext, axis=1)
        seg = np.repeat(np.arange(ns), sc)                                   # segment of every position
        pos = np.repeat(sf - np.concatenate([[0], np.cumsum(sc)[:-1]]), sc) + np.arange(int(sc.sum()))
        key = cen_o[pos, axis[seg]]
        perm = np.lexsort((key, seg))                                        # by segment, then by the split coordinate
        order[pos] = order[pos][perm]
        half = sc // 2
        first = np.column_stack([sf, sf + half]).ravel()
        count = np.column_stack([half, sc - half]).ravel()
        next_id += 2 * ns
    return order, np.concatenate(levels)
