"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OpenCV steps behind the reference's polygon extraction.

Reference: ``create_list_crd_from_color_output`` (kt_service/ai_tools/utils.py:1191-1279) and
``get_only_body_mask_contours`` (utils.py:1157-1188) call, per tissue colour,

    cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)   (utils.py:1247-1251)
    cv2.arcLength(cnt, True) * 0.001 -> cv2.approxPolyDP(cnt, eps, True)   (utils.py:1256-1257)

OpenCV (opencv-python-headless 4.13.0 in this image; the reference pins none) is a third-party dependency, so
this file restates its published algorithms -- Suzuki-Abe border following as OpenCV's tracer implements it, the
float32-segment / float64-sum arc length, and the Douglas-Peucker variant of ``approxPolyDP`` for closed integer
curves (start-point search by three farthest-point sweeps, explicit slice stack, final collinear clean-up).
The restatement is pinned: tests/test_oracle_contours.py runs it against cv2 itself on the reference's golden
label images, on random blobs and on random closed curves.  K13 (csrc/k13_polygons.cu) follows these functions
line by line; only tests/ import this module.
"""
from __future__ import annotations

import numpy as np

# chain code -> (dx, dy): 0 = east, counter-clockwise on screen (y grows downwards)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)

#: utils.py:1228-1233 dict order: class string and the code-image value (B<<2|G<<1|R of the BGR colour)
CLASS_ORDER = (("3", 3), ("0", 7), ("1", 1), ("2", 6))


def frame_connected_background(mask: np.ndarray) -> np.ndarray:
    """Zero pixels 4-connected to the image frame through zero pixels (the outside of every external contour)."""
    H, W = mask.shape
    bg = mask == 0
    reach = np.zeros_like(bg)
    stack = [(y, x) for y in range(H) for x in (0, W - 1) if bg[y, x]] + [(y, x) for x in range(W) for y in (0, H - 1) if bg[y, x]]
    for y, x in stack:
        reach[y, x] = True
    while stack:
        y, x = stack.pop()
        for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
            if 0 <= yy < H and 0 <= xx < W and bg[yy, xx] and not reach[yy, xx]:
                reach[yy, xx] = True
                stack.append((yy, xx))
    return reach


def trace_border(mask: np.ndarray, y0: int, x0: int, simple: bool):
    """Outer border of the 8-connected component whose raster-first pixel is (y0, x0); points (x, y) in OpenCV's
    order.  ``simple``: CHAIN_APPROX_SIMPLE (a point where the chain code changes), else every border pixel.
    Returns None when the border holds a pixel that precedes (y0, x0) in raster order (not a first pixel)."""
    H, W = mask.shape

    def on(y, x):
        return 0 <= y < H and 0 <= x < W and mask[y, x] != 0

    s = 4
    while True:                                   # clockwise from the west neighbour
        s = (s - 1) & 7
        if on(y0 + DY[s], x0 + DX[s]) or s == 4:
            break
    if s == 4 and not on(y0 + DY[4], x0 + DX[4]):
        return [(x0, y0)]                         # isolated pixel
    y1, x1 = y0 + DY[s], x0 + DX[s]
    y3, x3, prev_s = y0, x0, s ^ 4
    pts = []
    while True:
        while True:                               # counter-clockwise from the pixel we came from
            s = (s + 1) & 7
            y4, x4 = y3 + DY[s], x3 + DX[s]
            if on(y4, x4):
                break
        if not simple or s != prev_s:
            pts.append((x3, y3))
        prev_s = s
        if y4 < y0 or (y4 == y0 and x4 < x0):
            return None
        if y4 == y0 and x4 == x0 and y3 == y1 and x3 == x1:
            break
        y3, x3, s = y4, x4, (s + 4) & 7
    return pts


def find_external_contours(mask: np.ndarray, simple: bool = True):
    """cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE | NONE): list of (n, 2) int32 arrays (x, y), in
    OpenCV's order (the contour found last in the raster scan comes first)."""
    m = np.asarray(mask) != 0
    H, W = m.shape
    reach = frame_connected_background(m)
    up = np.zeros_like(m)
    up[1:] = m[:-1]
    left = np.zeros_like(m); left[:, 1:] = m[:, :-1]
    upl = np.zeros_like(m); upl[1:, 1:] = m[:-1, :-1]
    upr = np.zeros_like(m); upr[1:, :-1] = m[:-1, 1:]
    outside_left = np.ones_like(m); outside_left[:, 1:] = reach[:, :-1]
    tips = m & ~left & ~up & ~upl & ~upr & outside_left
    out = []
    ys, xs = np.nonzero(tips)
    for y, x in zip(ys[::-1], xs[::-1]):
        pts = trace_border(m, int(y), int(x), simple)
        if pts is not None:
            out.append(np.asarray(pts, np.int32).reshape(-1, 2))
    return out


def arc_length_closed(pts: np.ndarray) -> float:
    """cv2.arcLength(cnt, True): float32 segment lengths summed in float64."""
    p = np.asarray(pts).reshape(-1, 2).astype(np.float32)
    if len(p) <= 1:
        return 0.0
    d = p - np.roll(p, 1, axis=0)
    seg = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
    total = 0.0
    for v in seg:                                  # sequential float64 accumulation, like the C loop
        total += float(v)
    return total


def approx_poly_dp_closed(pts: np.ndarray, eps: float) -> np.ndarray:
    """cv2.approxPolyDP(cnt, eps, True) for an integer contour."""
    src = [(int(x), int(y)) for x, y in np.asarray(pts).reshape(-1, 2)]
    count = len(src)
    if count == 0:
        return np.zeros((0, 2), np.int32)
    eps = float(eps) * float(eps)
    dst = []
    stack = []
    # 1. approximately the two farthest points
    pos = 0
    right_start = 0
    le_eps = False
    start_pt = (-1000000, -1000000)
    for _ in range(3):
        max_dist = 0.0
        pos = (pos + right_start) % count
        start_pt = src[pos]
        pos = pos + 1 if pos + 1 < count else 0
        for j in range(1, count):
            pt = src[pos]
            pos = pos + 1 if pos + 1 < count else 0
            dx, dy = float(pt[0] - start_pt[0]), float(pt[1] - start_pt[1])
            dist = dx * dx + dy * dy
            if dist > max_dist:
                max_dist = dist
                right_start = j
        le_eps = max_dist <= eps
    # 2. the two initial slices
    if not le_eps:
        slice_start = pos % count
        right_end = slice_start
        slice_end = right_start = (right_start + slice_start) % count
        stack.append((right_start, right_end))
        stack.append((slice_start, slice_end))
    else:
        dst.append(start_pt)
    # 3. split until every slice is within eps of its chord
    while stack:
        s_start, s_end = stack.pop()
        end_pt = src[s_end]
        pos = s_start
        start_pt = src[pos]
        pos = pos + 1 if pos + 1 < count else 0
        if pos != s_end:
            # distance of every inner point to the chord SEGMENT (OpenCV >= 4.9; older releases measured to the
            # infinite line), first maximum wins.  Integer keys = squared distance x chord length^2: exact order.
            dx, dy = end_pt[0] - start_pt[0], end_pt[1] - start_pt[1]
            L = dx * dx + dy * dy
            max_key = 0
            r_start = 0
            while pos != s_end:
                pt = src[pos]
                pos = pos + 1 if pos + 1 < count else 0
                px, py = pt[0] - start_pt[0], pt[1] - start_pt[1]
                t = px * dx + py * dy
                if t <= 0:
                    key = (px * px + py * py) * L
                elif t >= L:
                    key = ((pt[0] - end_pt[0]) ** 2 + (pt[1] - end_pt[1]) ** 2) * L
                else:
                    c = py * dx - px * dy
                    key = c * c
                if key > max_key:
                    max_key = key
                    r_start = (pos + count - 1) % count
            le = float(max_key) <= eps * float(L)
        else:
            le = True
            r_start = 0
        if le:
            dst.append(start_pt)
        else:
            stack.append((r_start, s_end))
            stack.append((s_start, r_start))
    # 4. drop points on [almost] straight lines
    count = new_count = len(dst)
    pos = count - 1
    start_pt = dst[pos]; pos = pos + 1 if pos + 1 < count else 0
    wpos = pos
    pt = dst[pos]; pos = pos + 1 if pos + 1 < count else 0
    i = 0
    while i < count and new_count > 2:
        end_pt = dst[pos]; pos = pos + 1 if pos + 1 < count else 0
        dx, dy = float(end_pt[0] - start_pt[0]), float(end_pt[1] - start_pt[1])
        dist = abs((pt[0] - start_pt[0]) * dy - (pt[1] - start_pt[1]) * dx)
        inner = (pt[0] - start_pt[0]) * (end_pt[0] - pt[0]) + (pt[1] - start_pt[1]) * (end_pt[1] - pt[1])
        if dist * dist <= 0.5 * eps * (dx * dx + dy * dy) and dx != 0 and dy != 0 and inner >= 0:
            new_count -= 1
            dst[wpos] = start_pt = end_pt
            wpos = wpos + 1 if wpos + 1 < count else 0
            pt = dst[pos]; pos = pos + 1 if pos + 1 < count else 0
            i += 2
            continue
        dst[wpos] = start_pt = pt
        wpos = wpos + 1 if wpos + 1 < count else 0
        pt = end_pt
        i += 1
    return np.asarray(dst[:new_count], np.int32).reshape(-1, 2)


def label_polygons(code: np.ndarray, body: np.ndarray | None):
    """The polygon list of create_list_crd_from_color_output on a code image: [(class string, (n, 2) int32)], the
    tissue polygons closed as utils.py:1260-1266 closes them, then the body outline (class '4', every border pixel,
    the last contour with >= 5 points in OpenCV's order, utils.py:1173-1184)."""
    out = []
    for cls, val in CLASS_ORDER:
        for cnt in find_external_contours(code == val, simple=True):
            ap = approx_poly_dp_closed(cnt, 0.001 * arc_length_closed(cnt))
            if len(ap) > 2 and not np.array_equal(ap[0], ap[-1]):
                ap = np.concatenate([ap, ap[:1]], 0)
            out.append((cls, ap))
    if body is not None:
        chosen = None
        if np.asarray(body).any():
            for cnt in find_external_contours(np.asarray(body) != 0, simple=False):
                if len(cnt) < 5:
                    continue
                chosen = cnt[:-1] if np.array_equal(cnt[0], cnt[-1]) else cnt
        out.append(("4", chosen))
    return out
