"""The oracle against the ACTUAL reference code on fresh random inputs (beyond the frozen golden vectors).

Runs only where the reference tree is mounted (the authoring container); skipped on the GPU box.
``oracle/ref_import.py`` imports kt_service/ai_tools/utils.py unmodified with its four missing
third-party modules stubbed."""
import numpy as np
import pytest

from eitsynthai_b200 import synth
from oracle import imaging as O
from oracle.ref_import import DuckDataset, load_reference_utils, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return load_reference_utils()


class _Det:
    def __init__(self, xyxy):
        self.xyxy = xyxy


def test_classic_norm_random_windows(ref):
    rng = np.random.default_rng(0)
    for _ in range(12):
        level, width = int(rng.integers(-800, 800)), int(rng.integers(2, 1500)) * 2
        px = rng.integers(-3000, 3000, (64, 96)).astype(np.int16)
        assert np.array_equal(O.classic_norm(px, level, width), ref.classic_norm(px, level, width)), (level, width)


def test_body_mask_random_blobs(ref):
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:256, 0:256]
    for k in range(10):
        hu = np.full((256, 256), -1000, np.int32)
        for _ in range(int(rng.integers(1, 6))):
            cy, cx, ry, rx = rng.integers(20, 236), rng.integers(20, 236), rng.integers(5, 70), rng.integers(5, 70)
            hu[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = int(rng.integers(-400, 900))
        hu += rng.integers(-40, 41, hu.shape)
        intercept, slope = (-1024, 1) if k % 2 else (0, 1)
        px = (hu - intercept).astype(np.int16)
        want = ref.get_axial_slice_body_mask(DuckDataset(px, intercept=intercept, slope=slope))
        got = O.body_mask(px, intercept, slope)
        assert np.array_equal(got, want), k
        assert np.array_equal(O.largest_contour_fill_np(O.open5(O.hu_threshold(px, intercept, slope))), want), k


def test_rib_selection_random_boxes(ref):
    rng = np.random.default_rng(2)
    for _ in range(200):
        k = int(rng.integers(0, 30))
        b = rng.uniform(0, 512, (k, 4)).astype(np.float32)
        if k > 4:
            b[2, 1] = b[3, 1]
        custom = int(rng.integers(-5, 6))
        want = ref.search_number_axial_slice(_Det(b), custom)
        assert O.search_number_axial_slice(b, custom) == list(want)


def test_label_image_cleanup_random(ref):
    rng = np.random.default_rng(3)
    S = 96
    yy, xx = np.mgrid[0:S, 0:S]
    for it in range(12):
        base = rng.choice([0, 1, 3, 6, 7], p=[.25, .3, .15, .15, .15], size=(S // 8, S // 8)).astype(np.uint8)
        code = np.kron(base, np.ones((8, 8), np.uint8))
        nz = rng.random((S, S)) < rng.choice([0.01, 0.05, 0.15])
        code[nz] = rng.choice([0, 1, 3, 6, 7], size=int(nz.sum()))
        body = np.where(((yy - 48) / 40) ** 2 + ((xx - 48) / 44) ** 2 < 1, 255, 0).astype(np.uint8)
        bgr = O.code_to_bgr(code)
        assert np.array_equal(O.code_to_bgr(O.clear_codes(body, code)), ref.clear_color_output(body, bgr)), it
        assert np.array_equal(O.code_to_bgr(O.highlight_small_codes(code)), ref.highlight_small_masks(bgr)), it
        final = O.highlight_small_codes(O.clear_codes(body, code))
        want_list = ref.create_list_crd_from_color_output(O.code_to_bgr(final), [0.7, 0.7], body)
        assert O.polygons_from_codes(final, [0.7, 0.7], body) == want_list, it


def test_front_slice_random_orientations(ref):
    import cv2
    rng = np.random.default_rng(4)
    vol = rng.integers(-1000, 1500, (11, 32, 32)).astype(np.int16)          # already in InstanceNumber order
    img3d = np.stack(list(vol), axis=-1)
    for pp in ("HFS", "FFS", "HFP", "FFP"):
        for iop in ([1, 0, 0, 0, 1, 0], [-1, 0, 0, 0, 1, 0], [1, 0, 0, 0, -1, 0], [-1, 0, 0, 0, -1, 0]):
            for po in (None, ["L", "P"], ["R", "A"], ["L", "A"]):
                sag = ref.axial_to_sagittal(img3d, pp, iop, po)
                front = np.ascontiguousarray(sag[:, :, sag.shape[-1] // 2])
                assert np.array_equal(O.front_rows(vol, pp, iop, po), front), (pp, iop, po)
                assert np.array_equal(O.minmax_u8(front), cv2.normalize(front, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U))
