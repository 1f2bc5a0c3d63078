"""CPU oracle for the reference-owned imaging arithmetic (TEST INFRASTRUCTURE ONLY).

A restatement -- not a copy -- of the numpy/OpenCV/scipy arithmetic in
``/root/reference/kt_service/ai_tools/utils.py`` and ``ai_tools.py`` for the hot path
(SURVEY.md §8 rows a3, a5-a10, a15-a20).  Each function cites the reference lines it
follows.  It is pinned against outputs of the *actual* reference functions (stub-import,
``oracle/ref_import.py``) frozen under ``tests/golden/`` by ``oracle/gen_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product package never does.

Label images are handled both as the reference's BGR colour image and as a compact
*code* image: ``code = B<<2 | G<<1 | R`` per pixel with channel values in {0,255}
(black 0, muscle/red 1, adipose 3, lung 6, bone 7).
"""
from __future__ import annotations

from collections import Counter

import cv2
import numpy as np
from scipy import ndimage as ndi

# BGR colours, kt_service/ai_tools/utils.py:468-473
COLORS = {"bone": (255, 255, 255), "muscles": (0, 0, 255), "lung": (255, 255, 0),
          "adipose": (0, 255, 255)}
CLASS_NAMES = ("bone", "muscles", "lung", "adipose")          # class id -> name, utils.py:498-505
CODE_OF_CLASS = (7, 1, 6, 3)
CODE_BLACK, CODE_MUSCLE, CODE_ADIPOSE, CODE_LUNG, CODE_BONE = 0, 1, 3, 6, 7


def code_to_bgr(code: np.ndarray) -> np.ndarray:
    out = np.empty(code.shape + (3,), np.uint8)
    out[..., 0] = ((code >> 2) & 1) * 255
    out[..., 1] = ((code >> 1) & 1) * 255
    out[..., 2] = (code & 1) * 255
    return out


def bgr_to_code(img: np.ndarray) -> np.ndarray:
    assert set(np.unique(img)).issubset({0, 255}), "colour image is not a pure 0/255 image"
    return (((img[..., 0] > 0) << 2) | ((img[..., 1] > 0) << 1) | (img[..., 2] > 0)).astype(np.uint8)


# ---------------------------------------------------------------------------- a6
def classic_norm(volume: np.ndarray, window_level: int = 40, window_width: int = 400) -> np.ndarray:
    """utils.py:272-313 -- clip to the window, scale to 0..255 (truncation), rotate 180.

    The reference evaluates ``((clip - lo) / (hi - lo) * 255).astype(uint8)`` in float64;
    for the integer inputs involved that equals ``((clip - lo) * 255) // (hi - lo)``
    (checked for every value in tests/test_oracle_imaging.py).
    """
    lo = window_level - window_width // 2
    hi = window_level + window_width // 2
    c = np.clip(volume.astype(np.int64), lo, hi)
    u8 = (((c - lo) * 255) // (hi - lo)).astype(np.uint8)
    return np.ascontiguousarray(u8[..., ::-1, ::-1])


# ---------------------------------------------------------------------------- a7/a8
def hu_threshold(pixel_array: np.ndarray, intercept: int, slope: int) -> np.ndarray:
    """utils.py:551-566 -- flipud, HU = slope*px + intercept wrapped to int16, (-500, 1000)."""
    px = np.flipud(pixel_array).astype(np.int64)
    hu = (slope * px + intercept).astype(np.int16)            # .astype(int16) wraps, utils.py:559
    return ((hu > -500) & (hu < 1000)).astype(np.uint8)


def open5(mask01: np.ndarray) -> np.ndarray:
    """utils.py:562,569 -- MORPH_OPEN with a 5x5 box (erode then dilate, cv2 border rules)."""
    return cv2.morphologyEx(mask01, cv2.MORPH_OPEN, np.ones((5, 5), np.uint8))


def largest_contour_fill(mask01: np.ndarray) -> np.ndarray:
    """utils.py:572-582 -- external contours, max by contourArea, filled with 255."""
    contours, _ = cv2.findContours(mask01, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    best = max(contours, key=cv2.contourArea, default=None)
    if best is None:
        return mask01                                          # untouched 0/1 mask, utils.py:579
    out = np.zeros_like(mask01)
    cv2.drawContours(out, [best], 0, 255, -1)
    return out


def largest_contour_fill_np(mask01: np.ndarray) -> np.ndarray:
    """Library-free statement of ``largest_contour_fill`` (what the CUDA kernel implements).

    * the pixels outside every external contour are the 4-connected non-foreground
      pixels reachable from the (padded) image frame;
    * each 8-connected component of the remainder is one external contour filled;
    * ``cv2.contourArea`` of such a contour equals ``N4 + N3/2`` where N4/N3 count the
      2x2 pixel blocks with 4 / exactly 3 pixels in the filled region;
    * cv2 lists contours in reverse raster order of their first pixel and Python's
      ``max`` keeps the first maximum, so ties go to the region starting last.
    """
    fg = mask01 > 0
    pad = np.pad(~fg, 1, constant_values=True)
    lab4, _ = ndi.label(pad)
    outside = (lab4 == lab4[0, 0])[1:-1, 1:-1]
    filled = ~outside
    lab8, n = ndi.label(filled, structure=np.ones((3, 3), int))
    if n == 0:
        return mask01
    fp = np.pad(filled.astype(np.int32), 1)
    lp = np.pad(lab8, 1)
    s = fp[:-1, :-1] + fp[1:, :-1] + fp[:-1, 1:] + fp[1:, 1:]
    owner = np.maximum(np.maximum(lp[:-1, :-1], lp[1:, :-1]), np.maximum(lp[:-1, 1:], lp[1:, 1:]))
    area2 = np.bincount(owner[s == 4], minlength=n + 1) * 2 + np.bincount(owner[s == 3], minlength=n + 1)
    best = max(range(n, 0, -1), key=lambda i: area2[i])        # label order = raster order of first pixel
    return np.where(lab8 == best, 255, 0).astype(np.uint8)


def body_mask(pixel_array: np.ndarray, intercept: int, slope: int) -> np.ndarray:
    """get_axial_slice_body_mask, utils.py:526-585."""
    return largest_contour_fill(open5(hu_threshold(pixel_array, intercept, slope)))


def body_mask_nii(hu_img: np.ndarray) -> np.ndarray:
    """get_axial_slice_body_mask_nii, utils.py:588-618 (no flip, no rescale)."""
    m = ((hu_img > -500) & (hu_img < 1000)).astype(np.uint8)
    return largest_contour_fill(open5(m))


def apply_mask(norm_u8: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """ai_tools.py:212 -- cv2.bitwise_and(x, x, mask=m): x where m != 0 else 0."""
    return np.where(mask != 0, norm_u8, 0).astype(np.uint8)


# ---------------------------------------------------------------------------- a2/a3
def front_rows(slices_sorted: np.ndarray, patient_position: str = "HFS",
               iop=(1, 0, 0, 0, 1, 0), patient_orientation=None) -> np.ndarray:
    """Un-normalised coronal image (N, W) -- utils.py:114-163 + ai_tools.py:98-99.

    ``slices_sorted`` is (N, H, W) in InstanceNumber order.  The reference stacks to
    (H, W, N), transposes to (N, W, H), applies the orientation flips and takes plane
    ``H // 2`` of the last axis; that is one row per slice.
    """
    n, h, w = slices_sorted.shape
    flip_z = patient_position == "FFS"                         # utils.py:130-132
    flip_x = iop[0] == -1                                      # utils.py:148-149
    flip_y = iop[4] == -1                                      # utils.py:150-151
    if patient_position != "HFS" and patient_orientation:      # utils.py:155-160
        if patient_orientation[0] == "L":
            flip_x = not flip_x
        if patient_orientation[1] == "P":
            flip_z = not flip_z
    row = h // 2
    if flip_y:
        row = h - 1 - row
    rows = slices_sorted[:, row, :]
    if flip_x:
        rows = rows[:, ::-1]
    if flip_z:
        rows = rows[::-1]
    return np.ascontiguousarray(rows)


def minmax_u8(img: np.ndarray) -> np.ndarray:
    """cv2.normalize(x, None, 0, 255, NORM_MINMAX, CV_8U) -- ai_tools.py:101.

    Restated arithmetic (OpenCV 4.13, checked in tests against cv2 itself): scale and
    shift in float64, cast to float32, one fused multiply-add per pixel in float32,
    round-half-even, saturate.
    """
    mn, mx = float(img.min()), float(img.max())
    d = mx - mn
    scale = 255.0 * (1.0 / d if d > np.finfo(np.float64).eps else 0.0)
    shift = 0.0 - mn * scale
    a, b = np.float32(scale), np.float32(shift)
    v = (img.astype(np.float64) * float(a) + float(b)).astype(np.float32)   # == fmaf for |v|<2^24
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def front_slice_norm(slices_sorted, patient_position="HFS", iop=(1, 0, 0, 0, 1, 0),
                     patient_orientation=None) -> np.ndarray:
    return minmax_u8(front_rows(slices_sorted, patient_position, iop, patient_orientation))


# ---------------------------------------------------------------------------- a5
def search_number_axial_slice(xyxy: np.ndarray, custom_number_slise: int = 0,
                              image_width: int = 512) -> list:
    """utils.py:166-269 -- right-of-midline boxes, stable sort by y1, ribs 6 and 7 (0-based 5, 6)."""
    xyxy = np.asarray(xyxy, np.float32).reshape(-1, 4)
    mid = image_width / 2
    right = [b for b in xyxy if b[0] > mid]
    right.sort(key=lambda b: b[1])                             # stable
    if len(right) < 7:
        return []                                              # reference raises -> returns []
    y6, y7 = right[5][1], right[6][1]
    between = int(np.float32(abs(np.float32(y6 + y7))) / 2)
    return [int(y6), int(y7), between + custom_number_slise]


# ---------------------------------------------------------------------------- a10
def get_axial_slice_size(img) -> object:
    """utils.py:1282-1307 -- height if it is 256 or 512, else the (clobbered) default ``[]``."""
    if img is None or not hasattr(img, "shape"):
        return []
    h = img.shape[0]
    return h if h in (256, 512) else []


# ---------------------------------------------------------------------------- a15/a16
def class_union_masks(masks: np.ndarray, cls: np.ndarray, size: int) -> np.ndarray:
    """Per-class union of instance masks, (4, S, S) bool -- the information content of
    create_segmentations_masks (utils.py:437-523): saturating adds of one colour per class."""
    out = np.zeros((4, size, size), bool)
    for m, c in zip(masks, cls):
        c = int(c)
        if 0 <= c <= 3:
            out[c] |= np.asarray(m) > 0
    return out


def create_segmentations_masks(masks: np.ndarray, cls: np.ndarray, size: int) -> dict:
    u = class_union_masks(masks, cls, size)
    out = {}
    for c, name in enumerate(CLASS_NAMES):
        img = np.zeros((size, size, 3), np.uint8)
        img[u[c]] = COLORS[name]
        out[name] = img
    return out


def overlay_codes(union: np.ndarray) -> np.ndarray:
    """overlay_segmentation_masks, utils.py:395-434: saturating add of the class colours ==
    bitwise OR of the 3-bit colour codes."""
    code = np.zeros(union.shape[1:], np.uint8)
    for c in range(4):
        code[union[c]] |= CODE_OF_CLASS[c]
    return code


# ---------------------------------------------------------------------------- a17
_NB8 = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))   # utils.py:734-736


def clear_codes(body_mask: np.ndarray, code: np.ndarray, min_polygon_size: int = 5) -> np.ndarray:
    """clear_color_output, utils.py:691-755, on code images.

    1. black & body==255 -> muscle.  2. 4-connected components of the pixels that are
    neither black nor muscle.  3. components with < 5 px, in scipy label order, take the
    most common non-background colour among the 8-neighbours of their pixels (direction-
    major, then pixel order; Counter keeps first-seen on ties; the component's own pixels
    count; the image is updated in place so later components see earlier repaints);
    no such neighbour -> muscle.
    """
    out = code.copy()
    h, w = out.shape
    out[(out == CODE_BLACK) & (body_mask == 255)] = CODE_MUSCLE
    nonbg = (out != CODE_BLACK) & (out != CODE_MUSCLE)
    lab, n = ndi.label(nonbg)
    if n == 0:
        return out
    sizes = np.bincount(lab.ravel(), minlength=n + 1)
    small = np.nonzero(sizes[1:] < min_polygon_size)[0] + 1
    if small.size == 0:
        return out
    objs = ndi.find_objects(lab)
    for idx in small:
        sl = objs[idx - 1]
        ys, xs = np.nonzero(lab[sl] == idx)
        ys = ys + sl[0].start
        xs = xs + sl[1].start
        votes = []
        for dy, dx in _NB8:
            ny, nx = ys + dy, xs + dx
            ok = (ny >= 0) & (ny < h) & (nx >= 0) & (nx < w)
            for v in out[ny[ok], nx[ok]]:
                if v != CODE_BLACK and v != CODE_MUSCLE:
                    votes.append(int(v))
        out[ys, xs] = Counter(votes).most_common(1)[0][0] if votes else CODE_MUSCLE
    return out


# ---------------------------------------------------------------------------- a18
def highlight_small_codes(code: np.ndarray, point_threshold: int = 5) -> np.ndarray:
    """highlight_small_masks, utils.py:758-843, on code images.

    For bone, muscle, adipose (the "air" colour (0,150,255) never occurs): external
    contours (CHAIN_APPROX_SIMPLE) of the exact-colour mask of the *input*; contours with
    <= 5 points are filled with the most common colour of the 1-px ring around the filled
    contour, read from the progressively updated output in raster order, ignoring the
    target colour and black; no candidate -> unchanged.
    """
    out = code.copy()
    for target in (CODE_BONE, CODE_MUSCLE, CODE_ADIPOSE):      # dict order utils.py:782-787
        mask = np.where(code == target, 255, 0).astype(np.uint8)
        contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for cnt in contours:
            if len(cnt) > point_threshold:
                continue
            cm = np.zeros(code.shape, np.uint8)
            cv2.drawContours(cm, [cnt], -1, 255, cv2.FILLED)
            ring = cv2.dilate(cm, np.ones((3, 3), np.uint8)) - cm
            votes = [int(v) for v in out[ring == 255] if v != target and v != CODE_BLACK]
            fill = Counter(votes).most_common(1)[0][0] if votes else target
            out[cm == 255] = fill
    return out


# ---------------------------------------------------------------------------- a19
def create_color_codes(union: np.ndarray, body_mask=None) -> np.ndarray:
    """create_color_output, utils.py:989-1010 on codes: overlay -> (clear if mask) -> highlight."""
    code = overlay_codes(union)
    if body_mask is not None and np.any(body_mask):
        code = clear_codes(body_mask, code)
    return highlight_small_codes(code)


# ---------------------------------------------------------------------------- a20
def body_contour_string(body_mask) -> object:
    """get_only_body_mask_contours, utils.py:1157-1188: the *last* external contour with >= 5
    points as '4 x y x y ...'; ``[]`` when there is none."""
    res = []
    if body_mask is None or not body_mask.any():
        return res                                             # findContours([]) raises -> []
    binary = body_mask if body_mask.dtype == np.uint8 else (body_mask > 0).astype(np.uint8) * 255
    contours, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    for cnt in contours:
        if len(cnt) < 5:
            continue
        pts = cnt.reshape(-1, 2)
        res = "4 " + " ".join(f"{int(x)} {int(y)}" for x, y in pts)
    return res


def polygons_from_codes(code: np.ndarray, pixel_spacing, body_mask=None) -> list:
    """create_list_crd_from_color_output, utils.py:1191-1279.

    Net effect of the RGB<->BGR double swap: channel triples are matched as stored, in the
    order adipose "3", bone "0", muscle "1", lung "2"; external contours (SIMPLE) are
    simplified with approxPolyDP(eps = 0.001 * arcLength) and closed.
    """
    out = []
    for target, name in ((CODE_ADIPOSE, "3"), (CODE_BONE, "0"), (CODE_MUSCLE, "1"), (CODE_LUNG, "2")):
        mask = np.where(code == target, 255, 0).astype(np.uint8)
        contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for cnt in contours:
            approx = cv2.approxPolyDP(cnt, 0.001 * cv2.arcLength(cnt, True), True)
            pts = approx.reshape(-1, 2)
            if len(pts) > 2 and not np.array_equal(pts[0], pts[-1]):
                pts = np.vstack([pts, pts[:1]])
            out.append(name + " " + " ".join(f"{x} {y}" for x, y in pts))
    if body_mask is not None:
        out.append(body_contour_string(body_mask))
    return [str(pixel_spacing[0]), str(pixel_spacing[1])] + out
