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

# triangles per BVH leaf (RPX_BVH_LEAF: developer override for measurements; 1..8, the packed device nodes hold <= 8)
LEAF_CELLS = min(8, max(1, int(os.environ.get("RPX_BVH_LEAF", "4"))))


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
    Boxes are padded by 1e-9 of the mesh size so the slab test on the device stays conservative."""
    tri = points[cells]                      # M x 3 x 3
    lo, hi = tri.min(axis=1), tri.max(axis=1)
    cen = tri.mean(axis=1)
    pad = 1e-9 * max(float((points.max(axis=0) - points.min(axis=0)).max()), 1e-300)
    order = np.arange(len(cells), dtype=np.int64)
    nodes = []
    stack = [(0, len(cells), -1, 0)]         # first, count, parent node, which child slot
    while stack:
        first, count, parent, slot = stack.pop()
        ids = order[first:first + count]
        k = len(nodes)
        rec = np.zeros(8)
        rec[0:3] = lo[ids].min(axis=0) - pad
        rec[3:6] = hi[ids].max(axis=0) + pad
        nodes.append(rec)
        if parent >= 0:
            nodes[parent][6 + slot] = k
        if count <= leaf_cells:
            rec[6], rec[7] = -(first) - 1, count
            continue
        c = cen[ids]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        half = count // 2
        part = np.argpartition(c[:, axis], half)
        order[first:first + count] = ids[part]
        # right child pushed first so that the left child is built (numbered) next
        stack.append((first + half, count - half, k, 1))
        stack.append((first, half, k, 0))
    return order, np.array(nodes)
