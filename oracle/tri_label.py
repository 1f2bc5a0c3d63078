"""CPU oracle for the mesh labeller (TEST INFRASTRUCTURE ONLY).

* ``parse_contours`` / ``prepare_polygons`` restate how ``create_mesh`` and
  ``divide_triangles_into_groups`` turn the polygon strings into the sorted polygon list
  (femm_generator.py:445-459, 49-60, 88-115).
* ``label_triangles`` calls the C restatement of ``process_triangle`` (oracle/tri_label.c);
  ``label_triangles_py`` is an independent pure-Python statement for small cases.

PARITY UNPINNED against Shapely/GEOS (not installed; the reference has no expected labels).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_build", "liboracle_tri.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "tri_label.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_tri_label.restype = C.c_int
        _lib.oracle_tri_label.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_tri_margins.restype = C.c_int
        _lib.oracle_tri_margins.argtypes = _lib.oracle_tri_label.argtypes
    return _lib


def parse_contours(polygon_strings, outer_index=None):
    """create_mesh, femm_generator.py:454-459: every polygon string except the outer one becomes
    a list of floats [cls, x, y, x, y, ...]."""
    out = []
    for k, s in enumerate(polygon_strings):
        if k == outer_index or not isinstance(s, str):
            continue
        out.append(list(map(float, s.strip().split(" "))))
    return out


def prepare_polygons(contours):
    """divide_triangles_into_groups :49-60 + build_polygons_with_area :88-115.

    Drops contours with fewer than 9 numbers, closes rings, computes |shoelace| area
    (== shapely Polygon.area for a ring without holes) and stable-sorts ascending by area.
    Returns (poly_xy [V,2] f64, poly_off [P+1] i32, poly_cls [P] i32, areas [P] f64).
    """
    polys = []
    for c in contours:
        if len(c) < 9:
            continue
        cls = int(c[0])
        pts = [(float(c[i]), float(c[i + 1])) for i in range(1, len(c) - 1, 2)]
        if pts[0] != pts[-1]:
            pts.append(pts[0])
        a = np.asarray(pts, np.float64)
        area = abs(0.5 * float(np.sum(a[:-1, 0] * a[1:, 1] - a[1:, 0] * a[:-1, 1])))
        polys.append((a, cls, area))
    polys.sort(key=lambda t: t[2])
    if not polys:
        return np.zeros((0, 2)), np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0)
    xy = np.ascontiguousarray(np.concatenate([p[0] for p in polys]))
    off = np.zeros(len(polys) + 1, np.int32)
    off[1:] = np.cumsum([len(p[0]) for p in polys])
    return xy, off, np.asarray([p[1] for p in polys], np.int32), np.asarray([p[2] for p in polys])


def label_triangles(nodes_xy, tri, poly_xy, poly_off, poly_cls, outer_cls=4) -> np.ndarray:
    nodes_xy = np.ascontiguousarray(nodes_xy, np.float64)
    tri = np.ascontiguousarray(tri, np.int64)
    poly_xy = np.ascontiguousarray(poly_xy, np.float64)
    poly_off = np.ascontiguousarray(poly_off, np.int32)
    poly_cls = np.ascontiguousarray(poly_cls, np.int32)
    out = np.empty(len(tri), np.int32)
    rc = _load().oracle_tri_label(nodes_xy.ctypes.data, tri.ctypes.data, len(tri), poly_xy.ctypes.data,
                                  poly_off.ctypes.data, poly_cls.ctypes.data, len(poly_cls), outer_cls,
                                  out.ctypes.data)
    assert rc == 0
    return out


# ------------------------------------------------------------------ pure-Python cross-check
def _contains(ring, q):
    inside = False
    for (ux, uy), (vx, vy) in zip(ring[:-1], ring[1:]):
        if (uy > q[1]) != (vy > q[1]):
            if q[0] < (vx - ux) * (q[1] - uy) / (vy - uy) + ux:
                inside = not inside
    return inside


def _clip(poly, a, b):
    out = []
    dx, dy = b[0] - a[0], b[1] - a[1]
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = dx * (p[1] - a[1]) - dy * (p[0] - a[0])
        sq = dx * (q[1] - a[1]) - dy * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def _area(poly):
    return 0.5 * sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[i][1] * poly[(i + 1) % len(poly)][0]
                     for i in range(len(poly)))


def label_triangles_py(nodes_xy, tri, poly_xy, poly_off, poly_cls, outer_cls=4) -> np.ndarray:
    out = np.empty(len(tri), np.int32)
    rings = [[tuple(v) for v in poly_xy[poly_off[p]:poly_off[p + 1]]] for p in range(len(poly_cls))]
    for t, (i, j, k) in enumerate(tri):
        a, b, c = tuple(nodes_xy[i]), tuple(nodes_xy[j]), tuple(nodes_xy[k])
        a2 = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
        if a2 < 0:
            b, c, a2 = c, b, -a2
        area = 0.5 * a2
        ctr = ((a[0] + b[0] + c[0]) / 3.0, (a[1] + b[1] + c[1]) / 3.0)
        best, mx = outer_cls, 0.0
        for ring, cls in zip(rings, poly_cls):
            if cls == outer_cls:
                continue
            if _contains(ring, ctr):
                best = cls
                break
            if not area > 0:
                continue
            sub = ring[:-1]
            sign = 1.0 if _area(sub) >= 0 else -1.0
            for e0, e1 in ((a, b), (b, c), (c, a)):
                sub = _clip(sub, e0, e1)
                if not sub:
                    break
            inter = sign * _area(sub) if sub else 0.0
            if not inter > 1e-9 * area:          # noise floor, see tri_label.c
                inter = 0.0
            if inter / area > 0.5:
                best = cls
                break
            if inter > mx:
                mx, best = inter, cls
        out[t] = best
    return out


def decision_margins(nodes_xy, tri, poly_xy, poly_off, poly_cls, outer_cls=4) -> np.ndarray:
    """Per triangle, how far the closest comparison of process_triangle was from flipping (dimensionless, see
    oracle_tri_margins in tri_label.c).  Margins >= 1e-6 are six orders of magnitude above fp64 rounding: such a
    triangle gets the same label from any fp64 implementation of the predicates, GEOS included."""
    nodes_xy = np.ascontiguousarray(nodes_xy, np.float64)
    tri = np.ascontiguousarray(tri, np.int64)
    poly_xy = np.ascontiguousarray(poly_xy, np.float64)
    poly_off = np.ascontiguousarray(poly_off, np.int32)
    poly_cls = np.ascontiguousarray(poly_cls, np.int32)
    out = np.empty(len(tri), np.float64)
    rc = _load().oracle_tri_margins(nodes_xy.ctypes.data, tri.ctypes.data, len(tri), poly_xy.ctypes.data,
                                    poly_off.ctypes.data, poly_cls.ctypes.data, len(poly_cls), outer_cls, out.ctypes.data)
    assert rc == 0
    return out


# ------------------------------------------------------------------ exact rational statement (small cases)
def label_triangles_exact(nodes_xy, tri, poly_xy, poly_off, poly_cls, outer_cls=4) -> np.ndarray:
    """process_triangle with every predicate evaluated in exact rational arithmetic (``fractions.Fraction`` over the
    fp64 inputs): point-in-polygon, the Sutherland-Hodgman clip and both area comparisons have no rounding at all, and
    no noise-floor rule is needed -- a lower-dimensional overlap has an intersection area of exactly 0.  The centroid
    is the fp64 mean GEOS / the kernel compute (it is an input of the predicate, not part of it).  Pure Python with
    bounding-box rejection: for meshes of a few hundred triangles."""
    from fractions import Fraction as Fr
    out = np.empty(len(tri), np.int32)
    rings, boxes = [], []
    for p in range(len(poly_cls)):
        r = [(Fr(float(x)), Fr(float(y))) for x, y in poly_xy[poly_off[p]:poly_off[p + 1]]]
        rings.append(r)
        xs, ys = [v[0] for v in r], [v[1] for v in r]
        boxes.append((min(xs), min(ys), max(xs), max(ys)))

    def contains(ring, q):
        inside = False
        for (ux, uy), (vx, vy) in zip(ring[:-1], ring[1:]):
            if (uy > q[1]) != (vy > q[1]) and q[0] < (vx - ux) * (q[1] - uy) / (vy - uy) + ux:
                inside = not inside
        if not inside:
            return False
        for (ux, uy), (vx, vy) in zip(ring[:-1], ring[1:]):        # strictly inside: not on the boundary
            if (vx - ux) * (q[1] - uy) == (vy - uy) * (q[0] - ux) and min(ux, vx) <= q[0] <= max(ux, vx) and \
                    min(uy, vy) <= q[1] <= max(uy, vy):
                return False
        return True

    def clip(poly, a, b):
        res = []
        dx, dy = b[0] - a[0], b[1] - a[1]
        n = len(poly)
        for i in range(n):
            p, q = poly[i], poly[(i + 1) % n]
            sp = dx * (p[1] - a[1]) - dy * (p[0] - a[0])
            sq = dx * (q[1] - a[1]) - dy * (q[0] - a[0])
            if sp >= 0:
                res.append(p)
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                res.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        return res

    def area2(poly):
        return sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[i][1] * poly[(i + 1) % len(poly)][0] for i in range(len(poly)))

    ring_sign = [1 if area2(r[:-1]) >= 0 else -1 for r in rings]
    for t, (i, j, k) in enumerate(tri):
        fa, fb, fc = nodes_xy[i], nodes_xy[j], nodes_xy[k]
        ctr = (Fr(float((fa[0] + fb[0] + fc[0]) / 3.0)), Fr(float((fa[1] + fb[1] + fc[1]) / 3.0)))
        a, b, c = (Fr(float(fa[0])), Fr(float(fa[1]))), (Fr(float(fb[0])), Fr(float(fb[1]))), (Fr(float(fc[0])), Fr(float(fc[1])))
        a2 = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
        if a2 < 0:
            b, c, a2 = c, b, -a2
        tx0, ty0 = min(a[0], b[0], c[0]), min(a[1], b[1], c[1])
        tx1, ty1 = max(a[0], b[0], c[0]), max(a[1], b[1], c[1])
        best, mx = outer_cls, Fr(0)
        for ring, cls, box, sg in zip(rings, poly_cls, boxes, ring_sign):
            if cls == outer_cls:
                continue
            if box[0] > tx1 or box[2] < tx0 or box[1] > ty1 or box[3] < ty0:
                continue                                           # disjoint boxes: not contained, empty intersection
            if contains(ring, ctr):
                best = cls
                break
            if not a2 > 0:
                continue
            sub = ring[:-1]
            for e0, e1 in ((a, b), (b, c), (c, a)):
                sub = clip(sub, e0, e1)
                if not sub:
                    break
            inter2 = sg * area2(sub) if sub else Fr(0)             # twice the intersection area
            if inter2 * 2 > a2:                                    # inter / tri_area > 0.5
                best = cls
                break
            if inter2 > mx:
                mx, best = inter2, cls
        out[t] = best
    return out
