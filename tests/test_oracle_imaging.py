"""Oracle (oracle/imaging.py) vs golden vectors produced by the ACTUAL reference code
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from eitsynthai_b200 import synth
from oracle import imaging as O
from oracle.gen_golden import DOCSTRING_BOXES, segmentation_case


@pytest.mark.parametrize("tag,seed,intercept", [("p0", 0, -1024), ("p3hu", 3, 0)])
def test_norm_body_apply(golden, tag, seed, intercept):
    px = synth.phantom_slice(seed, intercept)
    norm = O.classic_norm(px)
    assert np.array_equal(norm, golden[f"{tag}_norm"])
    body = O.body_mask(px, intercept, 1)
    assert np.array_equal(body, golden[f"{tag}_body"])
    assert np.array_equal(O.largest_contour_fill_np(O.open5(O.hu_threshold(px, intercept, 1))), body)
    assert np.array_equal(O.apply_mask(norm, body), golden[f"{tag}_normbody"])


def test_norm_every_int16(golden):
    allv = np.arange(-32768, 32768, dtype=np.int16).reshape(256, 256)
    assert np.array_equal(O.classic_norm(allv), golden["norm_all_int16"])


def test_body_mask_nii(golden):
    assert np.array_equal(O.body_mask_nii(synth.phantom_hu(3).astype(np.int16)), golden["p3hu_body_nii"])


@pytest.mark.parametrize("tag,pp,iop,po", [
    ("hfs", "HFS", [1, 0, 0, 0, 1, 0], None), ("ffs", "FFS", [1, 0, 0, 0, 1, 0], None),
    ("ffs_neg", "FFS", [-1, 0, 0, 0, -1, 0], ["L", "P"]), ("hfp", "HFP", [1, 0, 0, 0, -1, 0], ["L", "A"])])
def test_front(golden, tag, pp, iop, po):
    vol, inst = synth.phantom_series(40, seed=5, size=512)
    srt = vol[np.argsort(inst, kind="stable")]
    rows = O.front_rows(srt, pp, iop, po)
    assert np.array_equal(rows, golden[f"front_{tag}_raw"])
    assert np.array_equal(O.minmax_u8(rows), golden[f"front_{tag}_u8"])


def test_minmax_matches_cv2_adversarial():
    import cv2
    rng = np.random.default_rng(11)
    for _ in range(60):
        lo, hi = int(rng.integers(-3000, 0)), int(rng.integers(1, 3000))
        a = rng.integers(lo, hi + 1, (37, 512)).astype(np.int16)
        a[0, 0], a[0, 1] = lo, hi
        assert np.array_equal(O.minmax_u8(a), cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U))
    c = np.full((4, 512), 7, np.int16)
    assert np.array_equal(O.minmax_u8(c), cv2.normalize(c, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U))


def test_rib_known_answer(golden):
    # the reference's own docstring boxes -> [162, 201, 182]
    assert list(golden["rib_kat_custom0"]) == [162, 201, 182]
    assert O.search_number_axial_slice(DOCSTRING_BOXES, 0) == [162, 201, 182]
    assert O.search_number_axial_slice(DOCSTRING_BOXES, 2) == list(golden["rib_kat_custom2"]) == [162, 201, 184]
    assert O.search_number_axial_slice(DOCSTRING_BOXES[:8], 0) == list(golden["rib_kat_few"]) == []


@pytest.mark.parametrize("tag,seed,size,noise,use_body", [
    ("seg0", 0, 512, 0, True), ("seg1", 1, 512, 60, True), ("seg2", 2, 256, 25, False), ("seg3", 3, 512, 200, True)])
def test_label_image(golden, golden_polygons, tag, seed, size, noise, use_body):
    masks, cls = segmentation_case(seed, size, noise)
    union = O.class_union_masks(masks, cls, size)
    d = O.create_segmentations_masks(masks, cls, size)
    for c, name in enumerate(O.CLASS_NAMES):
        assert np.array_equal(union[c] * 255, golden[f"{tag}_cls_{name}"])
        assert np.array_equal(d[name][..., 0] | d[name][..., 1] | d[name][..., 2], golden[f"{tag}_cls_{name}"])
    code = O.overlay_codes(union)
    assert np.array_equal(O.code_to_bgr(code), golden[f"{tag}_overlay"])
    body = None
    if use_body:
        ic = -1024 if seed % 2 == 0 else 0
        body = O.body_mask(synth.phantom_slice(seed, ic, size=size), ic, 1)
        assert np.array_equal(O.code_to_bgr(O.clear_codes(body, code)), golden[f"{tag}_clear"])
    final = O.create_color_codes(union, body)
    assert np.array_equal(O.code_to_bgr(final), golden[f"{tag}_color"])
    assert O.polygons_from_codes(final, [0.753906, 0.753906], body) == golden_polygons[tag]


def test_axial_slice_size():
    assert O.get_axial_slice_size(np.zeros((512, 512))) == 512
    assert O.get_axial_slice_size(np.zeros((256, 256))) == 256
    assert O.get_axial_slice_size(np.zeros((300, 300))) == []
    assert O.get_axial_slice_size(None) == []


def test_crop_mask_int_is_python_slicing():
    """The integer crop of late-2025 ultralytics (n < 50 on the CPU): rounded bounds, half to even, and Python's
    meaning of negative slice bounds; equals the float crop for boxes with integer corners inside the map."""
    import torch
    from oracle import yolo_post as Y
    m = torch.ones((4, 12, 16))
    boxes = torch.tensor([[2.0, 3.0, 9.0, 8.0], [2.5, 3.5, 9.5, 8.5], [-1.2, 0.0, 7.0, 20.0], [3.0, -2.6, 30.0, -0.4]])
    out = Y.crop_mask_int(m, boxes)
    assert torch.equal(out[0], Y.crop_mask(m[:1], boxes[:1])[0])
    keep = torch.zeros((12, 16)); keep[4:8, 2:10] = 1            # 2.5 -> 2, 3.5 -> 4, 9.5 -> 10, 8.5 -> 8
    assert torch.equal(out[1], keep)
    keep = torch.zeros((12, 16)); keep[:, 15:7] = 1               # x1 = -1 -> columns [0, 15) cleared, [7, 16) cleared: nothing left
    assert torch.equal(out[2], keep)
    keep = torch.zeros((12, 16)); keep[9:12, 3:16] = 1            # y1 = -3 -> rows [0, 9) cleared; y2 = round(-0.4) = 0 -> rows [0, 12) cleared
    assert out[3].sum() == 0
    r = Y.process_mask(torch.randn(32, 16, 16), torch.randn(3, 32), torch.tensor([[0., 0, 64, 64]] * 3), (64, 64), crop="cpu")
    assert r.shape == (3, 64, 64)
