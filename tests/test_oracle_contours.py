"""oracle/contours.py (the OpenCV restatement K13 follows) against cv2 itself -- the dependency the reference calls
at utils.py:1247-1257 -- and against the polygon lists the real reference produced (tests/golden)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from eitsynthai_b200 import synth
from oracle import contours as C
from oracle import imaging as O
from oracle.gen_golden import segmentation_case


def _blobs(rng, H, W, p, smooth):
    a = rng.random((H, W)).astype(np.float32)
    if smooth:
        a = cv2.GaussianBlur(a, (0, 0), smooth)
    return (a > np.quantile(a, 1 - p)).astype(np.uint8) * 255


def test_contours_arclength_approx_match_cv2_on_random_masks():
    rng = np.random.default_rng(0)
    n_contours = 0
    for it in range(40):
        H, W = (int(v) for v in rng.integers(6, 72, 2))
        m = _blobs(rng, H, W, rng.uniform(0.15, 0.75), float(rng.choice([0, 1, 2, 4])))
        for simple in (True, False):
            ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE if simple else cv2.CHAIN_APPROX_NONE)
            mine = C.find_external_contours(m, simple)
            assert len(ref) == len(mine)
            for r, q in zip(ref, mine):
                assert np.array_equal(r.reshape(-1, 2), q)
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for cnt in ref:
            n_contours += 1
            L = cv2.arcLength(cnt, True)
            assert L == C.arc_length_closed(cnt)
            for f in (0.001, 0.01, 0.05, 0.2):
                assert np.array_equal(cv2.approxPolyDP(cnt, f * L, True).reshape(-1, 2), C.approx_poly_dp_closed(cnt, f * L)), (it, f)
    assert n_contours > 400


def test_approx_uses_the_distance_to_the_chord_segment():
    """OpenCV >= 4.9 measures to the chord segment, not to the infinite line: a point that projects beyond the chord's
    end is kept although it is close to the line."""
    tri = np.array([[0, 0], [40, 0], [43, 1], [40, 2], [0, 2]], np.int32)
    for eps in (0.5, 1.5, 2.5, 3.5):
        assert np.array_equal(cv2.approxPolyDP(tri.reshape(-1, 1, 2), eps, True).reshape(-1, 2), C.approx_poly_dp_closed(tri, eps))


@pytest.mark.parametrize("tag,seed,size,noise,use_body", [
    ("seg0", 0, 512, 0, True), ("seg2", 2, 256, 25, False), ("seg3", 3, 512, 200, True)])
def test_label_polygons_equal_the_reference_lists(golden_polygons, tag, seed, size, noise, use_body):
    masks, cls = segmentation_case(seed, size, noise)
    union = O.class_union_masks(masks, cls, size)
    body = None
    if use_body:
        ic = -1024 if seed % 2 == 0 else 0
        body = O.body_mask(synth.phantom_slice(seed, ic, size=size), ic, 1)
    final = O.create_color_codes(union, body)
    got = []
    for name, pts in C.label_polygons(final, body):
        got.append([] if pts is None else name + " " + " ".join(f"{x} {y}" for x, y in pts))
    assert got == golden_polygons[tag][2:]
