"""CUDA path (through the C-ABI) vs the CPU oracle and the golden vectors.  B200 only."""
import json
import os

import numpy as np
import pytest
import torch

from eitsynthai_b200 import synth
from oracle import imaging as O
from oracle import yolo_post as Y

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from eitsynthai_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ----------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_hu_window_matches_reference_golden(ops, golden, dtype):
    px = np.stack([synth.phantom_slice(0, -1024), synth.phantom_slice(3, 0)])
    body = np.stack([golden["p0_body"], golden["p3hu_body"]])
    u8, nchw = ops.hu_window(dev(px), body_mask=None, nchw_dtype=dtype)
    assert np.array_equal(u8[0].cpu().numpy(), golden["p0_norm"])
    assert np.array_equal(u8[1].cpu().numpy(), golden["p3hu_norm"])
    u8m, nchwm = ops.hu_window(dev(px), body_mask=dev(body), nchw_dtype=dtype)
    assert np.array_equal(u8m[0].cpu().numpy(), golden["p0_normbody"])
    assert np.array_equal(u8m[1].cpu().numpy(), golden["p3hu_normbody"])
    # ultralytics preprocess: u8 -> dtype -> /255, three equal channels
    want = u8m.cpu().to(dtype) / 255
    for c in range(3):
        assert torch.equal(nchwm[:, c].cpu(), want)
    if dtype != torch.float32:                       # channels-last storage: same logical tensor
        _, cl = ops.hu_window(dev(px), body_mask=dev(body), want_u8=False, nchw_dtype=dtype, channels_last=True)
        assert cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(cl, nchwm)
        _, cl2 = ops.hu_window(dev(px), rot180=False, want_u8=False, nchw_dtype=dtype, channels_last=True)
        assert torch.equal(cl2, ops.hu_window(dev(px), rot180=False, nchw_dtype=dtype)[1])


def test_hu_window_every_int16(ops, golden):
    allv = np.arange(-32768, 32768, dtype=np.int16).reshape(1, 256, 256)
    u8, _ = ops.hu_window(dev(allv), nchw_dtype=None)
    assert np.array_equal(u8[0].cpu().numpy(), golden["norm_all_int16"])


def test_hu_window_other_windows_and_no_rotation(ops):
    rng = np.random.default_rng(3)
    px = rng.integers(-2000, 3000, (3, 64, 128)).astype(np.int16)
    for level, width, rot in ((40, 400, False), (-600, 1500, True), (300, 3000, True)):
        u8, _ = ops.hu_window(dev(px), lo=level - width // 2, hi=level + width // 2, rot180=rot, nchw_dtype=None)
        want = O.classic_norm(px, level, width)
        if not rot:
            want = want[..., ::-1, ::-1]
        assert np.array_equal(u8.cpu().numpy(), want)


def test_hu_window_channels_last_ragged_shapes(ops):
    rng = np.random.default_rng(4)
    for shape in ((3, 40, 72), (2, 8, 8), (5, 256, 256), (1, 24, 104)):
        px = rng.integers(-400, 500, shape).astype(np.int16)
        m = (rng.random(shape) > 0.4).astype(np.uint8) * 255
        for dt in (torch.float16, torch.bfloat16):
            _, a = ops.hu_window(dev(px), body_mask=dev(m), want_u8=False, nchw_dtype=dt, channels_last=True)
            u8, b = ops.hu_window(dev(px), body_mask=dev(m), want_u8=True, nchw_dtype=dt, channels_last=False)
            assert torch.equal(a, b)
            assert np.array_equal(u8.cpu().numpy(), O.apply_mask(O.classic_norm(px), m))


def test_u8_to_nchw(ops):
    g = np.arange(256, dtype=np.uint8).repeat(64).reshape(1, 128, 128)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        out = ops.u8_to_nchw(dev(g), dt).cpu()
        want = torch.from_numpy(g).to(dt) / 255
        for c in range(3):
            assert torch.equal(out[:, c], want)


# ----------------------------------------------------------------------------------- K3 / K4
@pytest.mark.parametrize("tag,pp,iop,po", [
    ("hfs", "HFS", [1, 0, 0, 0, 1, 0], None), ("ffs", "FFS", [1, 0, 0, 0, 1, 0], None),
    ("ffs_neg", "FFS", [-1, 0, 0, 0, -1, 0], ["L", "P"]), ("hfp", "HFP", [1, 0, 0, 0, -1, 0], ["L", "A"])])
def test_front_rows_golden(ops, golden, tag, pp, iop, po):
    from eitsynthai_b200.host import front_geometry
    vol, inst = synth.phantom_series(40, seed=5, size=512)
    order = np.argsort(inst, kind="stable").astype(np.int32)
    row, fx, fz = front_geometry(512, pp, iop, po)
    rows, mm = ops.front_rows(dev(vol), dev(order), 40, row, fx, fz)
    assert np.array_equal(rows.cpu().numpy(), golden[f"front_{tag}_raw"])
    assert mm.cpu().tolist() == [int(golden[f"front_{tag}_raw"].min()), int(golden[f"front_{tag}_raw"].max())]
    u8 = ops.minmax_u8(rows, mm)
    assert np.array_equal(u8.cpu().numpy(), golden[f"front_{tag}_u8"])


def test_minmax_u8_adversarial(ops):
    rng = np.random.default_rng(11)
    for _ in range(20):
        lo, hi = int(rng.integers(-3000, 0)), int(rng.integers(1, 3000))
        a = rng.integers(lo, hi + 1, (37, 512)).astype(np.int16)
        a[0, 0], a[0, 1] = lo, hi
        mm = torch.tensor([lo, hi], dtype=torch.int32, device=DEV)
        assert np.array_equal(ops.minmax_u8(dev(a), mm).cpu().numpy(), O.minmax_u8(a))
    c = np.full((4, 512), 7, np.int16)
    mm = torch.tensor([7, 7], dtype=torch.int32, device=DEV)
    assert np.array_equal(ops.minmax_u8(dev(c), mm).cpu().numpy(), O.minmax_u8(c))


def test_letterbox_matches_cv2(ops):
    rng = np.random.default_rng(5)
    for (h, w, imgsz) in ((320, 512, 640), (40, 512, 640), (512, 512, 512), (301, 512, 640), (700, 512, 640)):
        g = rng.integers(0, 256, (h, w)).astype(np.uint8)
        nh, nw, top, bottom, left, right = Y.letterbox_geometry(h, w, imgsz)
        want = Y.preprocess(g, imgsz, torch.float32)
        out = ops.letterbox_nchw(dev(g[None]), nh, nw, top, left, nh + top + bottom, nw + left + right, torch.float32)
        assert out.shape == want.shape
        assert torch.equal(out.cpu(), want), (h, w, imgsz, (out.cpu() != want).sum())


def test_rib_select_known_answer_and_random(ops, golden):
    from oracle.gen_golden import DOCSTRING_BOXES
    rng = np.random.default_rng(0)
    cases = [DOCSTRING_BOXES, DOCSTRING_BOXES[:8], np.zeros((0, 4), np.float32)]
    for _ in range(40):
        k = int(rng.integers(0, 40))
        b = rng.uniform(0, 512, (k, 4)).astype(np.float32)
        if k > 3:
            b[1, 1] = b[2, 1]                       # tie on y1 -> stable order matters
        cases.append(b)
    max_k = 64
    xy = np.zeros((len(cases), max_k, 4), np.float32)
    kk = np.zeros(len(cases), np.int32)
    custom = rng.integers(-3, 4, len(cases)).astype(np.int32)
    for i, b in enumerate(cases):
        xy[i, :len(b)] = b
        kk[i] = len(b)
    out = ops.rib_select(dev(xy), dev(kk), 512.0, dev(custom)).cpu().numpy()
    assert list(ops.rib_select(dev(xy[:1]), dev(kk[:1])).cpu().numpy()[0]) == [162, 201, 182, 1]
    for i, b in enumerate(cases):
        want = O.search_number_axial_slice(b, int(custom[i]))
        if want == []:
            assert out[i, 3] == 0
        else:
            assert list(out[i]) == want + [1]


# ----------------------------------------------------------------------------------- K5
def _nms_case(ops, head, nc, **kw):
    B = head.shape[0]
    dets, idx, n = ops.nms(dev(head), nc, **kw)
    dets, idx, n = dets.cpu(), idx.cpu(), n.cpu()
    for b in range(B):
        want, widx = Y.nms(torch.from_numpy(head[b]).float(), nc, kw.get("conf", 0.3), kw.get("iou", 0.7),
                           kw.get("max_det", 300))
        assert int(n[b]) == want.shape[0], (b, int(n[b]), want.shape[0])
        assert torch.equal(idx[b, :n[b]].long(), widx)
        assert torch.equal(dets[b, :n[b]], want)


@pytest.mark.parametrize("n_cand", [0, 1, 50, 300, 3000])
def test_nms_random_heads_bit_exact(ops, n_cand):
    head, _ = synth.random_heads(3, n_cand, seed=n_cand + 1)
    _nms_case(ops, head, 4)


def test_nms_teacher_and_rib_shapes(ops):
    head, _ = synth.teacher_heads(seed=0)
    _nms_case(ops, head[None], 4)
    head256, _ = synth.teacher_heads(seed=2, size=256)
    _nms_case(ops, head256[None], 4)
    # rib model: nc=1, A=5460, dense overlapping boxes, score ties
    rng = np.random.default_rng(9)
    A = 5460
    h = np.zeros((2, 37, A), np.float32)
    h[:, 0] = rng.uniform(0, 640, (2, A)); h[:, 1] = rng.uniform(0, 416, (2, A))
    h[:, 2:4] = rng.uniform(8, 40, (2, 2, A))
    h[:, 4] = np.round(rng.uniform(0, 0.6, (2, A)), 2)        # many exactly equal scores
    h[:, 5:] = rng.normal(0, 1, (2, 32, A))
    _nms_case(ops, h, 1)
    _nms_case(ops, h, 1, conf=0.05, iou=0.5, max_det=100)


def test_nms_half_precision_heads(ops):
    head, _ = synth.random_heads(2, 200, seed=4)
    for dt in (torch.float16, torch.bfloat16):
        hq = torch.from_numpy(head).to(dt)
        dets, idx, n = ops.nms(hq.to(DEV), 4)
        for b in range(2):
            want, widx = Y.nms(hq[b].float(), 4)
            assert int(n[b]) == want.shape[0]
            assert torch.equal(idx[b, :n[b]].cpu().long(), widx)
            assert torch.equal(dets[b, :n[b]].cpu(), want)


# ----------------------------------------------------------------------------------- K6
def _unpack_bits(bits, W):
    b = np.unpackbits(bits, axis=-1, bitorder="little")
    return b[..., :W]


def _decode_case(ops, head, protos, variant, size, tol=1e-4):
    """Per-instance masks, overlay codes and areas vs the CPU restatement."""
    vname = "logit" if not (variant & 1) else "sigmoid"
    dets, idx, n = ops.nms(dev(head[None]), 4)
    code, area, bits = ops.mask_decode(dets, n, dev(protos[None]), variant, want_area=True, want_bits=True)
    r = Y.postprocess(torch.from_numpy(head), torch.from_numpy(protos), 4, (size, size), (size, size),
                      variant=vname, drop_empty=False, crop="cpu" if variant & 4 else "float")
    nn = int(n[0])
    assert nn == r["masks"].shape[0]
    got = _unpack_bits(bits[0, :nn].cpu().numpy(), size)
    want = r["masks"].numpy()
    bad = int((got != want).sum())
    assert bad <= tol * want.size, (bad, want.size)
    assert np.array_equal(area[0, :nn].cpu().numpy(), got.reshape(nn, size * size).sum(1))
    # overlay of the kernel's own masks == code image (exact); and close to the oracle's overlay
    cls = r["cls"].numpy().astype(int)
    assert np.array_equal(O.overlay_codes(O.class_union_masks(got, cls, size)), code[0].cpu().numpy())
    wcode = O.overlay_codes(O.class_union_masks(want, cls, size))
    assert (wcode != code[0].cpu().numpy()).mean() <= tol
    for c in range(4):
        a, b_ = wcode == O.CODE_OF_CLASS[c], code[0].cpu().numpy() == O.CODE_OF_CLASS[c]
        if a.any():
            assert (a & b_).sum() / (a | b_).sum() >= 0.999
    return bad


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("seed,size", [(0, 512), (1, 512), (2, 256)])
def test_mask_decode_teacher(ops, variant, seed, size):
    head, protos = synth.teacher_heads(seed=seed, size=size)
    _decode_case(ops, head, protos, variant, size)


@pytest.mark.parametrize("n_cand", [0, 1, 50, 300])
def test_mask_decode_random(ops, n_cand):
    head, protos = synth.random_heads(1, n_cand, seed=7 + n_cand)
    _decode_case(ops, head[0], protos[0], 0, 512)


def _few_box_heads(seed, n_boxes, size=512):
    """Teacher-like head with < 50 detections whose boxes hang over every image edge (negative and > size
    coordinates, x.5 bounds): the regime of the integer crop, including Python's negative slice bounds."""
    rng = np.random.default_rng(seed)
    head, protos = synth.random_heads(1, 0, seed=seed, size=size)
    head, protos = head[0], protos[0]
    head[4:8] = 0.01
    A = head.shape[1]
    slots = rng.choice(A, n_boxes, replace=False)
    for k, a in enumerate(slots):
        cx, cy = rng.uniform(-20, size + 20, 2)
        w, h = rng.uniform(30, 300, 2)
        if k % 3 == 0:                                           # bounds landing on .5 in prototype pixels (ties -> even)
            cx, w = 4 * round(cx / 4) + 2.0, 8 * round(w / 8) + 4.0
        head[0, a], head[1, a], head[2, a], head[3, a] = cx, cy, w, h
        head[4 + k % 4, a] = rng.uniform(0.4, 0.9) - 1e-3 * k
    return head, protos


@pytest.mark.parametrize("seed,n_boxes", [(0, 7), (1, 30), (2, 49), (3, 80)])
@pytest.mark.parametrize("path", [4 | 0x20, 4 | 0x10])
def test_mask_decode_int_crop_variant(ops, seed, n_boxes, path):
    """variant bit 2: the rounded-integer crop ultralytics applies on the CPU for < 50 masks (and the float crop
    from 50 masks on), on both the tensor-core and the CUDA-core path."""
    head, protos = _few_box_heads(seed, n_boxes)
    protos16 = protos.astype(np.float16).astype(np.float32)
    _decode_case(ops, head, protos16 if not (path & 0x10) else protos, path, 512)
    if not (path & 0x10):
        dets, _, n = ops.nms(dev(head[None]), 4)
        c_tc, _, _ = ops.mask_decode(dets, n, torch.from_numpy(protos16[None]).half().to(DEV), 4 | 0x20)
        c_cc, _, _ = ops.mask_decode(dets, n, torch.from_numpy(protos16[None]).half().to(DEV), 4 | 0x10)
        assert (c_tc != c_cc).float().mean() <= 1e-4


def test_mask_decode_batch_and_half_protos(ops):
    head, protos = synth.random_heads(4, 40, seed=21)
    dets, idx, n = ops.nms(dev(head), 4)
    code32, _, _ = ops.mask_decode(dets, n, dev(protos))
    for b in range(4):
        c1, _, _ = ops.mask_decode(dets[b:b + 1].contiguous(), n[b:b + 1].contiguous(), dev(protos[b:b + 1]))
        assert torch.equal(c1[0], code32[b])
    ph = torch.from_numpy(protos).half()
    code16, _, _ = ops.mask_decode(dets, n, ph.to(DEV))
    code16ref, _, _ = ops.mask_decode(dets, n, ph.float().to(DEV))
    assert torch.equal(code16, code16ref)          # fp16 protos are widened exactly
    cl = ph.to(DEV).contiguous(memory_format=torch.channels_last)     # NHWC storage, same result
    assert torch.equal(ops.mask_decode(dets, n, cl, 0x10)[0], code16)
    # NHWC fp16 is what the network emits: by default its logits come from warp-level MMAs (other summation order)
    assert (ops.mask_decode(dets, n, cl)[0] != code16).float().mean().item() <= 1e-4


@pytest.mark.parametrize("variant", [0, 1, 4])
@pytest.mark.parametrize("case", ["teacher", "random50", "random300", "empty", "fewbox", "size256"])
def test_mask_decode_warp_mma_path(ops, case, variant):
    """Default path for fp16 NHWC prototypes with 32 channels (the network's output): logits by mma.sync with fp16 hi + lo
    coefficients.  Against the scalar kernel (variant bit 4) and the CPU restatement, masks / areas / codes, every variant,
    partial tiles (256-pixel size) and fp32 heads whose coefficients are not fp16 numbers."""
    size = 512
    if case == "teacher":
        head, protos = synth.teacher_heads(seed=5)
        head, protos = head[None], protos[None]
    elif case == "fewbox":
        head, protos = _few_box_heads(2, 30)
        head, protos = head[None], protos[None]
    elif case == "size256":
        size = 256
        head, protos = synth.random_heads(3, 60, seed=13, size=256)
    else:
        head, protos = synth.random_heads(2, {"random50": 50, "random300": 300, "empty": 0}[case], seed=41)
    ph = torch.from_numpy(protos).half()
    cl = ph.to(DEV).contiguous(memory_format=torch.channels_last)
    dets, idx, n = ops.nms(dev(head), 4)
    mm, area_mm, bits_mm = ops.mask_decode(dets, n, cl, variant, want_area=True, want_bits=True)
    cc, area_cc, bits_cc = ops.mask_decode(dets, n, cl, variant | 0x10, want_area=True, want_bits=True)
    total = bits_cc.numel() * 8
    diff = int((np.unpackbits(bits_mm.cpu().numpy()) != np.unpackbits(bits_cc.cpu().numpy())).sum())
    assert diff <= 1e-5 * max(total, 1), (diff, total)
    assert (mm != cc).float().mean().item() <= 1e-4
    got_bits = bits_mm.cpu().numpy()
    for b in range(head.shape[0]):
        nn = int(n[b])
        got = _unpack_bits(got_bits[b, :nn], size)
        assert np.array_equal(area_mm[b, :nn].cpu().numpy(), got.reshape(nn, size * size).sum(1))
        r = Y.postprocess(torch.from_numpy(head[b]), ph[b].float(), 4, (size, size), (size, size),
                          variant="logit" if not (variant & 1) else "sigmoid", drop_empty=False,
                          crop="cpu" if variant & 4 else "float")
        want = r["masks"].numpy()
        assert int((got != want).sum()) <= 1e-4 * max(want.size, 1)
        wcode = O.overlay_codes(O.class_union_masks(want, r["cls"].numpy().astype(int), size))
        assert (wcode != mm[b].cpu().numpy()).mean() <= 1e-4


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("case", ["teacher", "random50", "random300", "empty"])
def test_mask_decode_tensor_core_path(ops, case, variant):
    """fp16 prototypes go through tcgen05.mma; (variant bit 5); the result must match the CUDA-core kernel (variant bit 4)
    and the CPU restatement within the mask tolerance, for NCHW and channels-last prototypes."""
    if case == "teacher":
        head, protos = synth.teacher_heads(seed=4)
        head, protos = head[None], protos[None]
    else:
        head, protos = synth.random_heads(2, {"random50": 50, "random300": 300, "empty": 0}[case], seed=31)
    ph = torch.from_numpy(protos).half()
    dets, idx, n = ops.nms(dev(head), 4)
    tc, area_tc, bits_tc = ops.mask_decode(dets, n, ph.to(DEV), variant | 0x20, want_area=True, want_bits=True)
    cc, area_cc, bits_cc = ops.mask_decode(dets, n, ph.to(DEV), variant | 0x10, want_area=True, want_bits=True)
    total = bits_cc.numel() * 8
    diff = int((np.unpackbits(bits_tc.cpu().numpy()) != np.unpackbits(bits_cc.cpu().numpy())).sum())
    assert diff <= 1e-5 * max(total, 1), (diff, total)
    assert (tc != cc).float().mean().item() <= 1e-4
    nhwc, _, _ = ops.mask_decode(dets, n, ph.to(DEV).contiguous(memory_format=torch.channels_last), variant | 0x20)
    assert torch.equal(nhwc, tc)
    for b in range(head.shape[0]):
        r = Y.postprocess(torch.from_numpy(head[b]), ph[b].float(), 4, (512, 512), (512, 512),
                          variant="logit" if variant == 0 else "sigmoid", drop_empty=False)
        want = O.overlay_codes(O.class_union_masks(r["masks"].numpy(), r["cls"].numpy().astype(int), 512))
        assert (want != tc[b].cpu().numpy()).mean() <= 1e-4


def test_mask_decode_tensor_core_small_size_ragged_batch(ops):
    """tcgen05 path at the 256-pixel size (64x64 prototypes: partial tiles), images with 0 / few / >128
    instances in one batch, and max_det smaller than the candidate count."""
    heads, protos = [], []
    for n_c, seed in ((0, 1), (7, 2), (300, 3)):
        h, p = synth.random_heads(1, n_c, seed=seed, size=256)
        heads.append(h[0]); protos.append(p[0])
    head, ph = np.stack(heads), torch.from_numpy(np.stack(protos)).half()
    for max_det in (300, 40):
        dets, idx, n = ops.nms(dev(head), 4, max_det=max_det)
        assert n.cpu().tolist()[0] == 0 and int(n[2]) <= max_det
        tc, a1, b1 = ops.mask_decode(dets, n, ph.to(DEV), 0x20, want_area=True, want_bits=True)
        cc, a2, b2 = ops.mask_decode(dets, n, ph.to(DEV), 0x10, want_area=True, want_bits=True)
        assert (tc != cc).float().mean().item() <= 1e-4
        assert int((np.unpackbits(b1.cpu().numpy()) != np.unpackbits(b2.cpu().numpy())).sum()) <= 1e-5 * b1.numel() * 8
        assert float((tc[0] != 0).sum()) == 0
        for b in range(3):
            r = Y.postprocess(torch.from_numpy(head[b]), ph[b].float(), 4, (256, 256), (256, 256), drop_empty=False)
            m, c = r["masks"].numpy()[:max_det], r["cls"].numpy().astype(int)[:max_det]
            want = O.overlay_codes(O.class_union_masks(m, c, 256))
            assert (want != tc[b].cpu().numpy()).mean() <= 1e-4


def test_codes_to_bgr(ops):
    code = np.random.default_rng(0).choice([0, 1, 3, 6, 7], (3, 64, 64)).astype(np.uint8)
    assert np.array_equal(ops.codes_to_bgr(dev(code)).cpu().numpy(), O.code_to_bgr(code))


# ----------------------------------------------------------------------------------- K8
def _polys(tag):
    from oracle import tri_label as TL
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "reference_polygons.json")) as f:
        strs = json.load(f)[tag][2:]
    outer = next((i for i, s in enumerate(strs) if isinstance(s, str) and s[:1] == "4"), None)
    return TL.prepare_polygons(TL.parse_contours(strs, outer))


@pytest.mark.parametrize("tag,pitch", [("seg0", 3.0), ("seg1", 2.5), ("seg2", 1.5), ("seg3", 1.08)])
def test_tri_label_bit_exact(ops, tag, pitch):
    from oracle import tri_label as TL
    xy, off, cls, _ = _polys(tag)
    bbox = (20, 40, 490, 470) if tag != "seg2" else (10, 20, 245, 235)
    nodes, tri = synth.delaunay_mesh(bbox, pitch, seed=3)
    want = TL.label_triangles(nodes, tri, xy, off, cls)
    got = ops.tri_label(dev(nodes), dev(tri), dev(xy), dev(off), dev(cls)).cpu().numpy()
    assert np.array_equal(got, want), (int((got != want).sum()), len(tri))
    assert len(np.unique(want)) >= 4


def test_tri_label_edge_cases(ops):
    from oracle import tri_label as TL
    xy, off, cls, _ = _polys("seg0")
    nodes, tri = synth.delaunay_mesh((100, 100, 200, 200), 5.0, seed=1)
    # no polygons at all -> outer class; empty triangle list; clockwise + degenerate triangles
    got = ops.tri_label(dev(nodes), dev(tri), dev(np.zeros((0, 2))), dev(np.zeros(1, np.int32)),
                        dev(np.zeros(0, np.int32))).cpu().numpy()
    assert (got == 4).all()
    assert ops.tri_label(dev(nodes), dev(tri[:0]), dev(xy), dev(off), dev(cls)).numel() == 0
    tri2 = tri[:, ::-1].copy()
    tri2[::7, 2] = tri2[::7, 1]                      # zero-area triangles
    want = TL.label_triangles(nodes, tri2, xy, off, cls)
    got = ops.tri_label(dev(nodes), dev(tri2), dev(xy), dev(off), dev(cls)).cpu().numpy()
    assert np.array_equal(got, want)


def test_tri_label_raster(ops):
    rng = np.random.default_rng(2)
    code = rng.choice([0, 1, 3, 6, 7], (512, 512)).astype(np.uint8)
    nodes, tri = synth.delaunay_mesh((-5, -5, 520, 520), 9.0, seed=4)
    got = ops.tri_label_raster(dev(nodes), dev(tri), dev(code)).cpu().numpy()
    c = nodes[tri].mean(1)
    px, py = np.floor(c[:, 0] + 0.5).astype(int), np.floor(c[:, 1] + 0.5).astype(int)
    ok = (px >= 0) & (px < 512) & (py >= 0) & (py < 512)
    lut = np.full(8, 4); lut[7], lut[1], lut[6], lut[3] = 0, 1, 2, 3
    want = np.where(ok, lut[code[py.clip(0, 511), px.clip(0, 511)]], 4)
    assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------- CC / K2
def _canon(lab):
    """scipy labels (1..n in raster order of first pixel) -> smallest pixel index per component."""
    from scipy import ndimage as ndi
    out = np.full(lab.shape, -2, np.int64)
    if lab.max() > 0:
        idx = np.arange(lab.size).reshape(lab.shape)
        mins = ndi.minimum(idx, lab, index=np.arange(1, lab.max() + 1))
        out[lab > 0] = np.asarray(mins)[lab[lab > 0] - 1]
    return out


@pytest.mark.parametrize("conn", [4, 8])
def test_cc_label_matches_scipy(ops, conn):
    from scipy import ndimage as ndi
    rng = np.random.default_rng(conn)
    st = None if conn == 4 else np.ones((3, 3), int)
    masks = []
    for dens, shape in ((0.5, (512, 512)), (0.62, (512, 512)), (0.3, (512, 512)), (0.9, (512, 512))):
        masks.append((rng.random(shape) < dens).astype(np.uint8))
    spiral = np.zeros((512, 512), np.uint8)              # one long snake through every strip
    spiral[::4, :] = 1
    spiral[2::8, -1] = 1; spiral[1::8, -1] = 1; spiral[3::8, -1] = 1
    spiral[6::8, 0] = 1; spiral[5::8, 0] = 1; spiral[7::8, 0] = 1
    masks += [spiral, np.zeros((512, 512), np.uint8), np.ones((512, 512), np.uint8)]
    m = np.stack(masks)
    got = ops.cc_label(dev(m), conn).cpu().numpy()
    for k in range(len(masks)):
        lab, _ = ndi.label(masks[k], structure=st)
        assert np.array_equal(got[k], _canon(lab)), k
    # other shapes: 256x256 and a ragged one
    for shape in ((256, 256), (100, 72), (3, 8)):
        mk = (rng.random((2,) + shape) < 0.55).astype(np.uint8)
        got = ops.cc_label(dev(mk), conn).cpu().numpy()
        for k in range(2):
            lab, _ = ndi.label(mk[k], structure=st)
            assert np.array_equal(got[k], _canon(lab)), (shape, k)


def test_cc_label_outside(ops):
    from scipy import ndimage as ndi
    rng = np.random.default_rng(5)
    m = (rng.random((3, 512, 512)) < 0.58).astype(np.uint8)
    got = ops.cc_label(dev(m), 4, link_outside=True).cpu().numpy()
    for k in range(3):
        lab, _ = ndi.label(np.pad(m[k], 1, constant_values=1))
        outside = (lab == lab[0, 0])[1:-1, 1:-1]
        assert np.array_equal(got[k] == -1, outside)
        inner = _canon(ndi.label(m[k])[0])
        sel = (m[k] > 0) & ~outside
        assert np.array_equal(got[k][sel], inner[sel])
        assert (got[k][m[k] == 0] == -2).all()


def test_body_mask_golden(ops, golden):
    px = np.stack([synth.phantom_slice(0, -1024), synth.phantom_slice(3, 0)])
    b0 = ops.body_mask(dev(px[:1]), 1, -1024, True).cpu().numpy()[0]
    b1 = ops.body_mask(dev(px[1:]), 1, 0, True).cpu().numpy()[0]
    assert np.array_equal(b0, golden["p0_body"])
    assert np.array_equal(b1, golden["p3hu_body"])
    hu = synth.phantom_hu(3).astype(np.int16)
    assert np.array_equal(ops.body_mask(dev(hu[None]), 1, 0, False).cpu().numpy()[0], golden["p3hu_body_nii"])


def test_body_mask_adversarial_vs_oracle(ops):
    """Several blobs with near-tied areas, holes, frame contact, nested islands, empty input."""
    rng = np.random.default_rng(8)
    cases = []
    for k in range(12):
        hu = np.full((512, 512), -1000, np.int32)
        yy, xx = np.mgrid[0:512, 0:512]
        for _ in range(int(rng.integers(1, 7))):
            cy, cx = rng.integers(30, 480, 2)
            ry, rx = rng.integers(8, 120, 2)
            hu[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = 50
        for _ in range(int(rng.integers(0, 5))):                 # holes and islands
            cy, cx = rng.integers(60, 450, 2)
            r = int(rng.integers(5, 40))
            hu[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = -900
            if rng.random() < 0.5:
                hu[(yy - cy) ** 2 + (xx - cx) ** 2 < (r // 2) ** 2] = 30
        hu += rng.integers(-30, 31, hu.shape)
        if k % 4 == 0:
            hu[rng.random(hu.shape) < 0.02] = 2000               # salt noise, removed by the opening
        cases.append((hu + 1024).astype(np.int16))
    two = np.full((512, 512), 24, np.int16)                      # two equal squares: tie -> the later one
    two[100:150, 100:150] = 1100; two[300:350, 300:350] = 1100
    cases += [two, np.full((512, 512), 0, np.int16), np.full((512, 512), 1100, np.int16)]
    px = np.stack(cases)
    got = ops.body_mask(dev(px), 1, -1024, True).cpu().numpy()
    for k in range(len(cases)):
        want = O.body_mask(px[k], -1024, 1)
        assert np.array_equal(got[k], want), (k, int((got[k] != want).sum()))
    for shape in ((3, 100, 72), (2, 40, 96), (1, 7, 32)):           # u8 fallback path and tiny bit-image cases
        odd = rng.integers(900, 1300, shape).astype(np.int16)
        odd[:, shape[1] // 4: 3 * shape[1] // 4, shape[2] // 5: 4 * shape[2] // 5] = 1100
        got = ops.body_mask(dev(odd), 1, -1024, True).cpu().numpy()
        for k in range(shape[0]):
            assert np.array_equal(got[k], O.body_mask(odd[k], -1024, 1)), (shape, k)
    small = rng.integers(0, 2000, (5, 256, 256)).astype(np.int16)
    small[:, 60:200, 50:210] = 1050
    got = ops.body_mask(dev(small), 1, -1024, True).cpu().numpy()
    for k in range(5):
        assert np.array_equal(got[k], O.body_mask(small[k], -1024, 1))


# ----------------------------------------------------------------------------------- K7
@pytest.mark.parametrize("tag,seed,size,use_body", [("seg0", 0, 512, True), ("seg1", 1, 512, True),
                                                    ("seg2", 2, 256, False), ("seg3", 3, 512, True)])
def test_label_cleanup_reference_golden(ops, golden, tag, seed, size, use_body):
    """Overlay image produced by the real reference -> our clean-up == the real reference's."""
    code = O.bgr_to_code(golden[f"{tag}_overlay"])
    body = None
    if use_body:
        ic = -1024 if seed % 2 == 0 else 0
        body = O.body_mask(synth.phantom_slice(seed, ic, size=size), ic, 1)
    got = ops.label_cleanup(dev(code[None]), None if body is None else dev(body[None]))[0].cpu().numpy()
    assert np.array_equal(O.code_to_bgr(got), golden[f"{tag}_color"])


def _noisy_codes(rng, S, it):
    base = rng.choice([0, 1, 3, 6, 7], p=[.2, .35, .15, .15, .15], size=(S // 8, S // 8)).astype(np.uint8)
    code = np.kron(base, np.ones((8, 8), np.uint8))
    nz = rng.random((S, S)) < rng.choice([0.005, 0.03, 0.12])
    code[nz] = rng.choice([0, 1, 3, 6, 7], size=int(nz.sum()))
    if it % 3 == 0:                                    # rings with nested islands, big 4-point rectangles
        code[5:20, 5:20] = 7; code[8:17, 8:17] = 6; code[11:13, 11:13] = 7
        code[25:35, 22:36] = 1; code[27:33, 24:34] = 6; code[29, 28] = 1
        code[40:44, 40:90] = 3; code[41, 60] = 0
    return code


def test_label_cleanup_adversarial_vs_oracle(ops):
    rng = np.random.default_rng(12)
    S = 128
    codes = np.stack([_noisy_codes(rng, S, it) for it in range(48)])
    yy, xx = np.mgrid[0:S, 0:S]
    bodies = np.stack([np.where(((yy - 64) / rng.integers(30, 64)) ** 2 + ((xx - 64) / rng.integers(30, 64)) ** 2 < 1, 255, 0)
                       for _ in range(48)]).astype(np.uint8)
    bodies[5] = 0                                        # empty body mask: clear_color_output is skipped
    got_nb = ops.label_cleanup(dev(codes.copy()), None).cpu().numpy()
    got_b = ops.label_cleanup(dev(codes.copy()), dev(bodies)).cpu().numpy()
    for k in range(len(codes)):
        assert np.array_equal(got_nb[k], O.highlight_small_codes(codes[k])), ("nobody", k)
        want = O.highlight_small_codes(O.clear_codes(bodies[k], codes[k]) if bodies[k].any() else codes[k])
        assert np.array_equal(got_b[k], want), ("body", k, int((got_b[k] != want).sum()))


def test_label_cleanup_full_size_pipeline(ops):
    """512x512: decoded teacher masks + specks -> overlay -> clean-up, vs the oracle end to end."""
    from oracle.gen_golden import segmentation_case
    for seed, noise in ((5, 150), (6, 400)):
        masks, cls = segmentation_case(seed, 512, noise)
        code = O.overlay_codes(O.class_union_masks(masks, cls, 512))
        body = O.body_mask(synth.phantom_slice(seed, -1024), -1024, 1)
        got = ops.label_cleanup(dev(code[None].copy()), dev(body[None]))[0].cpu().numpy()
        assert np.array_equal(got, O.create_color_codes(O.class_union_masks(masks, cls, 512), body))


def _real_set(k):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "reference_polygon_sets.npz"))
    xy, off, cls = z[f"set{k}_xy"], z[f"set{k}_off"], z[f"set{k}_cls"]
    return [[float(cls[p])] + xy[off[p]:off[p + 1]].reshape(-1).tolist() for p in range(len(cls))]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_tri_label_reference_real_polygon_sets(ops, k):
    """The six real-data polygon lists of the reference's mesh_service_trials.py (inputs only): rings
    that touch / retrace themselves, mm units and a class-4 body contour in set 6."""
    from eitsynthai_b200 import host
    from oracle import tri_label as TL
    contours = _real_set(k)
    outer = next((i for i, c in enumerate(contours) if int(c[0]) == 4), None)
    inner = [c for i, c in enumerate(contours) if i != outer]
    xy, off, cls, _ = TL.prepare_polygons([list(c) for c in inner])
    hx, hoff, hcls = host.prepare_polygons([list(c) for c in inner])
    assert np.array_equal(hx, xy) and np.array_equal(hoff, off) and np.array_equal(hcls, cls)
    lo, hi = xy.min(0), xy.max(0)
    pitch = float(max(hi - lo)) / 220.0
    nodes, tri = synth.delaunay_mesh((lo[0] - 2 * pitch, lo[1] - 2 * pitch, hi[0] + 2 * pitch, hi[1] + 2 * pitch), pitch, seed=k)
    want = TL.label_triangles(nodes, tri, xy, off, cls)
    got = ops.tri_label(dev(nodes), dev(tri), dev(xy), dev(off), dev(cls)).cpu().numpy()
    assert np.array_equal(got, want), (k, int((got != want).sum()), len(tri))
    assert len(np.unique(want)) >= 3
