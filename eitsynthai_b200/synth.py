"""Deterministic synthetic inputs for tests and benchmarks (host side, numpy only).

The reference's datasets and weights are unavailable offline, so every test and
benchmark of this repo runs on the seeded generators below:

* ``phantom_slice``   – 512x512 int16 chest-CT-like slice (stored value = HU + 1024)
* ``phantom_series``  – N slices with rib markers on the coronal mid-row
* ``phantom_structures`` / ``teacher_heads`` – YOLO-seg head tensors + prototypes that
  decode to the phantom's tissues (random-init weights give no detections)
* ``random_heads``    – head tensors with realistic statistics and a chosen candidate count
* ``delaunay_mesh``   – a triangular mesh standing in for Gmsh output

Everything is a pure function of its seed; no file or network access.
"""
from __future__ import annotations

import numpy as np

SLICE = 512
#: (cls, cy, cx, ry, rx)  class ids follow kt_service/ai_tools/utils.py:498-505
#: 0 bone, 1 muscles, 2 lung, 3 adipose
_ELLIPSES = {
    "fat": (3, 256, 256, 170, 220, -90),
    "muscle": (1, 256, 256, 150, 200, 40),
    "lung_l": (2, 240, 170, 90, 70, -800),
    "lung_r": (2, 240, 342, 90, 70, -800),
    "spine": (0, 360, 256, 25, 25, 400),
}


def _ellipse_level(shape, cy, cx, ry, rx):
    """Level-set (>0 inside) of an axis-aligned ellipse."""
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]].astype(np.float64)
    return 1.0 - ((yy - cy) / ry) ** 2 - ((xx - cx) / rx) ** 2


def phantom_hu(seed: int = 0, lung_scale: float = 1.0, size: int = SLICE) -> np.ndarray:
    """HU image of the phantom (int32), before storage offset."""
    s = size / 512.0
    hu = np.full((size, size), -1000, np.int32)
    for name in ("fat", "muscle", "lung_l", "lung_r", "spine"):
        _, cy, cx, ry, rx, val = _ELLIPSES[name]
        if name.startswith("lung"):
            ry, rx = ry * lung_scale, rx * lung_scale
        hu[_ellipse_level(hu.shape, cy * s, cx * s, ry * s, rx * s) > 0] = val
    # CT table: rows 470-479, cols 60-449
    hu[int(470 * s):int(480 * s), int(60 * s):int(450 * s)] = 200
    rng = np.random.default_rng(seed)
    hu += rng.integers(-15, 16, hu.shape, dtype=np.int32)
    return hu


def phantom_slice(seed: int = 0, intercept: int = -1024, lung_scale: float = 1.0,
                  size: int = SLICE) -> np.ndarray:
    """Stored int16 pixel values: ``hu - intercept`` (slope 1)."""
    return (phantom_hu(seed, lung_scale, size) - intercept).astype(np.int16)


def rib_marker_rows(n_slices: int = 320):
    """z positions of the 12 synthetic rib pairs (clipped to the series)."""
    step = max(1, (n_slices - 80) // 12)
    return [40 + step * r for r in range(12) if 40 + step * r < n_slices - 8]


def phantom_series(n_slices: int = 320, seed: int = 0, intercept: int = -1024,
                   shuffle_seed: int | None = 1, size: int = SLICE, z_range: tuple | None = None):
    """Synthetic series.

    Returns ``(pixels[n,H,W] int16 in *file order*, instance_numbers[n] int32)``; file
    order is a seeded shuffle of z so the InstanceNumber sort (utils.py:96) is exercised.
    Slice z carries phantom ``P(seed*1000 + z)`` with a z-dependent lung scale and, near
    each rib position, 400-HU blobs on row H/2 at x = W/2 +- (120 + 4 r).  ``z_range=(z0, z1)``
    generates only that shard of the series (slice content does not depend on the shard).
    """
    zs = np.arange(n_slices) if z_range is None else np.arange(z_range[0], z_range[1])
    vol = np.empty((len(zs), size, size), np.int16)
    ribs = rib_marker_rows(n_slices)
    yy, xx = np.mgrid[0:size, 0:size]
    for z in zs:
        scale = 0.6 + 0.4 * np.sin(np.pi * (z + 0.5) / n_slices)
        hu = phantom_hu(seed * 1000 + int(z), lung_scale=float(scale), size=size)
        for r, zr in enumerate(ribs):
            dz = abs(int(z) - zr)
            if dz <= 6:
                rad2 = 36 - dz * dz
                for sgn in (-1, 1):
                    cx = size // 2 + sgn * (120 + 4 * r) * size // 512
                    hu[(yy - size // 2) ** 2 + (xx - cx) ** 2 <= rad2] = 400
        vol[z - zs[0]] = (hu - intercept).astype(np.int16)
    inst = (zs + 1).astype(np.int32)
    if shuffle_seed is not None:
        perm = np.random.default_rng(shuffle_seed).permutation(len(zs))
        vol, inst = vol[perm], inst[perm]
    return vol, inst


# ----------------------------------------------------------------------------------------
# YOLO-seg head tensors


def anchor_count(h: int, w: int) -> int:
    return sum((h // s) * (w // s) for s in (8, 16, 32))


def phantom_structures(size: int = SLICE):
    """[(cls, level_set[size,size] float64, xyxy box)] for the phantom's tissue instances."""
    s = size / 512.0
    out = []
    lv = {k: _ellipse_level((size, size), v[1] * s, v[2] * s, v[3] * s, v[4] * s)
          for k, v in _ELLIPSES.items()}
    lungs = np.maximum(lv["lung_l"], lv["lung_r"])
    fat = np.minimum(lv["fat"], -lv["muscle"])                  # ring
    muscle = np.minimum(np.minimum(lv["muscle"], -lungs), -lv["spine"])
    for name, cls, level in (("spine", 0, lv["spine"]), ("muscle", 1, muscle),
                             ("lung_l", 2, lv["lung_l"]), ("lung_r", 2, lv["lung_r"]),
                             ("fat", 3, fat)):
        ys, xs = np.nonzero(level > 0)
        box = (float(xs.min()), float(ys.min()), float(xs.max() + 1), float(ys.max() + 1))
        out.append((cls, level, box))
    # a few rib-like bone specks around the lungs
    for k in range(10):
        ang = 2 * np.pi * k / 10
        cy, cx = (250 + 118 * np.sin(ang)) * s, (256 + 168 * np.cos(ang)) * s
        level = _ellipse_level((size, size), cy, cx, 7 * s, 10 * s)
        ys, xs = np.nonzero(level > 0)
        out.append((0, level, (float(xs.min()), float(ys.min()), float(xs.max() + 1),
                               float(ys.max() + 1))))
    return out


def teacher_heads(seed: int = 0, size: int = SLICE, per_structure=(8, 32), nc: int = 4,
                  nm: int = 32, jitter: float = 4.0, dtype=np.float32):
    """Head tensor ``(4+nc+nm, A)`` + protos ``(nm, size/4, size/4)`` that decode to the phantom.

    Protos: channel ``cls*8 + v`` is the 1/4-resolution level set of the union of class
    ``cls`` structures, perturbed by variant ``v``; coefficients are one-hot on the
    instance's class (variant drawn at random) plus N(0, 0.02) noise.  Candidates are
    planted at seeded anchor positions, all other anchors score ~0.01.
    """
    rng = np.random.default_rng(seed)
    structs = phantom_structures(size)
    A = anchor_count(size, size)
    head = np.zeros((4 + nc + nm, A), np.float32)
    head[4:4 + nc] = rng.uniform(0.0, 0.02, (nc, A)).astype(np.float32)
    head[0:2] = rng.uniform(0, size, (2, A)).astype(np.float32)
    head[2:4] = rng.uniform(4, 64, (2, A)).astype(np.float32)
    head[4 + nc:] = rng.normal(0, 0.3, (nm, A)).astype(np.float32)
    mh = size // 4
    protos = np.zeros((nm, mh, mh), np.float32)
    per_cls = nm // nc
    for cls in range(nc):
        lvl = None
        for c, level, _ in structs:
            if c == cls:
                lvl = level if lvl is None else np.maximum(lvl, level)
        small = np.clip(lvl, -0.25, 0.25).reshape(mh, 4, mh, 4).mean(axis=(1, 3))
        for v in range(per_cls):
            protos[cls * per_cls + v] = (small * (4.0 + v) + rng.normal(0, 0.02, small.shape)
                                         - 0.01 * v).astype(np.float32)
    used = rng.permutation(A)
    k = 0
    for cls, _, (x1, y1, x2, y2) in structs:
        cnt = int(rng.integers(per_structure[0], per_structure[1] + 1))
        for _ in range(cnt):
            a = used[k]
            k += 1
            j = rng.normal(0, jitter, 4)
            bx1, by1, bx2, by2 = x1 + j[0], y1 + j[1], x2 + j[2], y2 + j[3]
            head[0, a], head[1, a] = (bx1 + bx2) / 2, (by1 + by2) / 2
            head[2, a], head[3, a] = max(bx2 - bx1, 2.0), max(by2 - by1, 2.0)
            head[4:4 + nc, a] = rng.uniform(0.0, 0.05, nc)
            head[4 + cls, a] = rng.uniform(0.3, 0.95)
            coef = rng.normal(0, 0.02, nm)
            coef[cls * per_cls + int(rng.integers(0, per_cls))] += 1.0
            head[4 + nc:, a] = coef
    return head.astype(dtype), protos.astype(dtype)


def random_heads(batch: int, n_cand: int, seed: int = 0, size: int = SLICE, nc: int = 4,
                 nm: int = 32, dtype=np.float32):
    """Random head tensors ``(B, 4+nc+nm, A)`` + protos with ~``n_cand`` candidates above 0.3."""
    rng = np.random.default_rng(seed)
    A = anchor_count(size, size)
    head = np.empty((batch, 4 + nc + nm, A), np.float32)
    head[:, 0:2] = rng.uniform(0, size, (batch, 2, A))
    head[:, 2:4] = rng.gamma(2.0, size / 12.0, (batch, 2, A)) + 2.0
    head[:, 4:4 + nc] = rng.uniform(0.0, 0.25, (batch, nc, A))
    head[:, 4 + nc:] = rng.normal(0, 0.5, (batch, nm, A))
    for b in range(batch):
        idx = rng.choice(A, size=min(n_cand, A), replace=False)
        cls = rng.integers(0, nc, idx.size)
        head[b, 4 + cls, idx] = rng.uniform(0.3001, 0.99, idx.size)
    protos = rng.normal(0, 1.0, (batch, nm, size // 4, size // 4)).astype(np.float32)
    # smooth so masks have structure rather than salt-and-pepper
    protos = (protos + np.roll(protos, 1, -1) + np.roll(protos, 1, -2)
              + np.roll(protos, (1, 1), (-1, -2))) * 0.5
    return head.astype(dtype), protos.astype(dtype)


# ----------------------------------------------------------------------------------------
# meshes


def delaunay_mesh(bbox, pitch: float, seed: int = 0, inside=None):
    """Jittered-grid Delaunay mesh over ``bbox=(x0,y0,x1,y1)``.

    Returns ``(nodes[Nn,2] float64, tris[T,3] int64)``; when ``inside`` (callable on
    (x, y) arrays -> bool) is given, only triangles whose centroid is inside are kept and
    the node table is compacted (gmsh returns only used nodes too).
    """
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    x0, y0, x1, y1 = bbox
    gx = np.arange(x0, x1 + pitch, pitch)
    gy = np.arange(y0, y1 + pitch, pitch)
    X, Y = np.meshgrid(gx, gy)
    pts = np.stack([X.ravel(), Y.ravel()], 1)
    pts += rng.uniform(-0.3, 0.3, pts.shape) * pitch
    tri = Delaunay(pts).simplices.astype(np.int64)
    if inside is not None:
        c = pts[tri].mean(axis=1)
        tri = tri[inside(c[:, 0], c[:, 1])]
    used = np.unique(tri)
    remap = np.full(len(pts), -1, np.int64)
    remap[used] = np.arange(used.size)
    return np.ascontiguousarray(pts[used]), np.ascontiguousarray(remap[tri])
