"""Host-side logic of the hot path: the small, non-data-parallel decisions the reference
makes in Python (orientation rules, letterbox geometry, polygon list preparation).  The
data-parallel work itself lives in libeitb200 (``ops``)."""
from __future__ import annotations

import numpy as np


def front_geometry(height: int, patient_position="HFS", iop=(1, 0, 0, 0, 1, 0), patient_orientation=None):
    """Which row of every axial slice forms the coronal image, and the flips to apply.

    Follows axial_to_sagittal + the mid-plane pick (kt_service/ai_tools/utils.py:114-163,
    ai_tools.py:98-99): transpose (2,1,0); FFS -> reverse z (:130-132); IOP[0]==-1 -> reverse x
    (:148-149); IOP[4]==-1 -> reverse y (:150-151); if not HFS and PatientOrientation is given,
    [0]=='L' -> fliplr (x) and [1]=='P' -> flipud (z) (:155-160); plane ``H // 2`` of the y axis.
    Returns ``(row, flip_x, flip_z)``.
    """
    flip_z = patient_position == "FFS"
    flip_x = iop is not None and iop[0] == -1
    flip_y = iop is not None and iop[4] == -1
    if patient_position != "HFS" and patient_orientation:
        if patient_orientation[0] == "L":
            flip_x = not flip_x
        if patient_orientation[1] == "P":
            flip_z = not flip_z
    row = height // 2
    if flip_y:
        row = height - 1 - row
    return row, bool(flip_x), bool(flip_z)


def instance_order(instance_numbers) -> np.ndarray:
    """convert_to_3d sorts the datasets by int(InstanceNumber) with Python's stable sort
    (utils.py:96); returns the file-order indices in that order (int32)."""
    return np.argsort(np.asarray(instance_numbers, np.int64), kind="stable").astype(np.int32)


def letterbox_geometry(h: int, w: int, imgsz: int, stride: int = 32):
    """ultralytics LetterBox(auto=True) geometry -> (nh, nw, top, bottom, left, right)."""
    r = min(imgsz / h, imgsz / w)
    nw, nh = int(round(w * r)), int(round(h * r))
    dw, dh = (imgsz - nw) % stride, (imgsz - nh) % stride
    dw, dh = dw / 2, dh / 2
    return (nh, nw, int(round(dh - 0.1)), int(round(dh + 0.1)), int(round(dw - 0.1)), int(round(dw + 0.1)))


def scale_boxes_params(net_shape, orig_shape):
    """ultralytics scale_boxes: (gain, pad_x, pad_y) mapping network-input px to original px."""
    gain = min(net_shape[0] / orig_shape[0], net_shape[1] / orig_shape[1])
    pad_x = round((net_shape[1] - orig_shape[1] * gain) / 2 - 0.1)
    pad_y = round((net_shape[0] - orig_shape[0] * gain) / 2 - 0.1)
    return gain, pad_x, pad_y


def find_outer_index(polygon_strings):
    """find_outer_contour's fast path (femm_generator.py:588-590): the first class-'4' line."""
    for i, line in enumerate(polygon_strings):
        if isinstance(line, str) and line[:1] == "4":
            return i
    return None


def parse_contours(polygon_strings, outer_index=None):
    """create_mesh (femm_generator.py:454-459): every polygon line except the outer one, as
    [cls, x, y, x, y, ...] floats."""
    out = []
    for k, s in enumerate(polygon_strings):
        if k == outer_index or not isinstance(s, str) or not s.strip():
            continue
        out.append(list(map(float, s.strip().split(" "))))
    return out


def prepare_polygons(contours):
    """divide_triangles_into_groups (femm_generator.py:49-60) + build_polygons_with_area (:88-115):
    drop contours with < 9 numbers, close the rings, sort (stable) by ascending area.

    Returns ``(poly_xy [V,2] f64, poly_off [P+1] i32, poly_cls [P] i32)`` ready for ``ops.tri_label``.
    """
    polys = []
    for c in contours:
        if len(c) < 9:
            continue
        pts = np.asarray(c[1:1 + 2 * ((len(c) - 1) // 2)], np.float64).reshape(-1, 2)
        if not np.array_equal(pts[0], pts[-1]):
            pts = np.vstack([pts, pts[:1]])
        area = abs(0.5 * float(np.sum(pts[:-1, 0] * pts[1:, 1] - pts[1:, 0] * pts[:-1, 1])))
        polys.append((pts, int(c[0]), area))
    polys.sort(key=lambda t: t[2])
    if not polys:
        return np.zeros((0, 2), np.float64), np.zeros(1, np.int32), np.zeros(0, np.int32)
    xy = np.ascontiguousarray(np.concatenate([p[0] for p in polys]))
    off = np.zeros(len(polys) + 1, np.int32)
    off[1:] = np.cumsum([len(p[0]) for p in polys])
    return xy, off, np.asarray([p[1] for p in polys], np.int32)
