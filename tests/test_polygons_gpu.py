"""K13 (eitb_label_polygons / eitb_polygons_for_mesh) on the B200 against the reference's frozen polygon lists, against
OpenCV itself (the dependency the reference calls, utils.py:1247-1257) and against the pinned restatement."""
import numpy as np
import pytest
import torch

from eitsynthai_b200 import host, synth
from oracle import contours as C
from oracle import imaging as O
from oracle.gen_golden import segmentation_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eitsynthai_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def strings(polys, status, body_given):
    out = [f"{c} " + " ".join(f"{x} {y}" for x, y in p) for c, p in polys]
    if body_given and status & 8:
        out.append([])
    return out


def cv2_strings(code, body):
    """create_list_crd_from_color_output on a code image, with OpenCV (what the reference executes)."""
    cv2 = pytest.importorskip("cv2")
    out = []
    for name, val in C.CLASS_ORDER:
        contours, _ = cv2.findContours(np.where(code == val, 255, 0).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for cnt in contours:
            ap = cv2.approxPolyDP(cnt, 0.001 * cv2.arcLength(cnt, True), True).reshape(-1, 2)
            if len(ap) > 2 and not np.array_equal(ap[0], ap[-1]):
                ap = np.vstack([ap, ap[:1]])
            out.append(name + " " + " ".join(f"{x} {y}" for x, y in ap))
    if body is not None:
        res = []
        if body.any():
            contours, _ = cv2.findContours(body, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
            for cnt in contours:
                if len(cnt) >= 5:
                    res = "4 " + " ".join(f"{int(x)} {int(y)}" for x, y in cnt.reshape(-1, 2))
        out.append(res)
    return out


def _final(seed, size, noise, use_body):
    masks, cls = segmentation_case(seed, size, noise)
    union = O.class_union_masks(masks, cls, size)
    body = None
    if use_body:
        ic = -1024 if seed % 2 == 0 else 0
        body = O.body_mask(synth.phantom_slice(seed, ic, size=size), ic, 1)
    return O.create_color_codes(union, body), body


@pytest.mark.parametrize("tag,seed,size,noise,use_body", [
    ("seg0", 0, 512, 0, True), ("seg1", 1, 512, 60, True), ("seg2", 2, 256, 25, False), ("seg3", 3, 512, 200, True)])
def test_polygons_equal_the_reference_lists(ops, golden_polygons, tag, seed, size, noise, use_body):
    code, body = _final(seed, size, noise, use_body)
    lp = ops.label_polygons(dev(code)[None], None if body is None else dev(body)[None])
    (st, polys), = lp.to_host()
    assert st & 7 == 0
    assert strings(polys, st, use_body) == golden_polygons[tag][2:]


def _noisy(rng, H, W, smooth, p_black):
    import cv2
    f = [cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), smooth) if smooth else rng.random((H, W)) for _ in range(5)]
    code = np.array([0, 1, 3, 6, 7], np.uint8)[np.argmax(np.stack(f) + np.array([p_black, 0, 0, 0, 0])[:, None, None], 0)]
    b = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), max(smooth, 1.0))
    body = (b > np.quantile(b, 0.4)).astype(np.uint8) * 255
    return code, body


@pytest.mark.parametrize("H,W,smooth", [(32, 32, 0), (64, 96, 1.0), (96, 64, 0), (256, 256, 2.0), (512, 512, 3.0), (128, 1024, 1.5)])
def test_polygons_match_opencv_on_noisy_label_images(ops, H, W, smooth):
    rng = np.random.default_rng(H * 7 + W)
    codes, bodies = zip(*[_noisy(rng, H, W, smooth, p) for p in (0.0, 0.02, 0.1)])
    for with_body in (True, False):
        lp = ops.label_polygons(dev(np.stack(codes)), dev(np.stack(bodies)) if with_body else None, max_polys=H * W // 2,
                                max_points=3 * H * W)
        for b, (st, polys) in enumerate(lp.to_host()):
            assert st & 7 == 0, st
            want = cv2_strings(codes[b], bodies[b] if with_body else None)
            got = strings(polys, st, with_body)
            assert len(got) == len(want)
            assert got == want, next(i for i, (g, w) in enumerate(zip(got, want)) if g != w)
            if H * W <= 64 * 96:                                  # the pinned restatement agrees as well
                mine = [[] if p is None else n + " " + " ".join(f"{x} {y}" for x, y in p)
                        for n, p in C.label_polygons(codes[b], bodies[b] if with_body else None)]
                assert mine == want


def test_polygons_edge_cases(ops):
    z = np.zeros((2, 64, 64), np.uint8)
    z[1, 0, 0] = 7; z[1, 63, 63] = 1; z[1, 10:12, 0:64] = 3; z[1, 30, 30] = 6; z[1, 31, 31] = 6
    body = np.zeros((2, 64, 64), np.uint8)
    body[1, 5, 5:8] = 255                                          # 3 border pixels: skipped, the reference appends []
    lp = ops.label_polygons(dev(z), dev(body))
    res = lp.to_host()
    assert res[0] == (8, [])
    assert strings(res[1][1], res[1][0], True) == cv2_strings(z[1], body[1])
    # capacity overflow is reported, not silent
    code, _ = _noisy(np.random.default_rng(1), 64, 64, 0, 0.0)
    st = ops.label_polygons(dev(code)[None], None, max_polys=8).to_host()[0][0]
    assert st & 1
    st = ops.label_polygons(dev(code)[None], None, max_polys=4096, max_points=16).to_host()[0][0]
    assert st & 2


@pytest.mark.parametrize("tag,seed,size,noise,use_body", [("seg0", 0, 512, 0, True), ("seg3", 3, 512, 200, True), ("seg2", 2, 256, 25, False)])
def test_polygons_feed_the_triangle_labeller_on_the_device(ops, golden_polygons, tag, seed, size, noise, use_body):
    code, body = _final(seed, size, noise, use_body)
    lp = ops.label_polygons(dev(code)[None], None if body is None else dev(body)[None])
    xy, off, cls, n = ops.polygons_for_mesh(lp)
    strs = golden_polygons[tag][2:]
    outer = host.find_outer_index(strs)
    wxy, woff, wcls = host.prepare_polygons(host.parse_contours(strs, outer))
    P = int(n[0])
    assert P == len(wcls)
    assert np.array_equal(off[0, :P + 1].cpu().numpy(), woff) and np.array_equal(cls[0, :P].cpu().numpy(), wcls)
    assert np.array_equal(xy[0, :woff[-1]].cpu().numpy(), wxy)
    nodes, tri = synth.delaunay_mesh((10, 20, size - 20, size - 30), 4.0, seed=2)
    a = ops.tri_label(dev(nodes), dev(tri), xy[0, :woff[-1]].contiguous(), off[0, :P + 1].contiguous(), cls[0, :P].contiguous())
    b = ops.tri_label(dev(nodes), dev(tri), dev(wxy), dev(woff), dev(wcls))
    assert torch.equal(a, b)
