"""The kt_service mirror (reference names and signatures) on the GPU vs the golden vectors of the
real reference, and the whole post-CNN chain of the pipeline vs the oracle."""
import numpy as np
import pytest
import torch

from eitsynthai_b200 import synth
from oracle import imaging as O
from oracle import yolo_post as Y
from oracle.gen_golden import DOCSTRING_BOXES, _Results, segmentation_case
from oracle.ref_import import DuckDataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    from eitsynthai_b200.kt_service.ai_tools import utils
    utils.set_device("cuda:0")
    return utils


def test_utils_mirror_matches_reference_golden(U, golden, golden_polygons):
    px = synth.phantom_slice(0, -1024)
    assert np.array_equal(U.classic_norm(px), golden["p0_norm"])
    ds = DuckDataset(px, intercept=-1024, slope=1)
    body = U.get_axial_slice_body_mask(ds)
    assert np.array_equal(body, golden["p0_body"])
    assert np.array_equal(U.apply_body_mask(U.classic_norm(px), body), golden["p0_normbody"])
    assert np.array_equal(U.get_axial_slice_body_mask_nii(synth.phantom_hu(3).astype(np.int16)), golden["p3hu_body_nii"])
    assert U.get_pixel_spacing(ds) == [0.753906, 0.753906]
    assert U.get_axial_slice_size(px) == 512 and U.get_axial_slice_size(np.zeros((300, 300))) == []

    class D:
        xyxy = DOCSTRING_BOXES
    assert U.search_number_axial_slice(D()) == [162, 201, 182]
    assert U.search_number_axial_slice(D(), 2) == [162, 201, 184]
    D.xyxy = DOCSTRING_BOXES[:8]
    assert U.search_number_axial_slice(D()) == []                 # the reference's sentinel

    vol, inst = synth.phantom_series(40, seed=5, size=512)
    assert np.array_equal(U.front_slice_from_series(vol, inst), golden["front_hfs_u8"])
    assert np.array_equal(U.front_slice_from_series(vol, inst, "FFS", [-1, 0, 0, 0, -1, 0], ["L", "P"]), golden["front_ffs_neg_u8"])


@pytest.mark.parametrize("tag,seed,size,noise,use_body", [("seg0", 0, 512, 0, True), ("seg2", 2, 256, 25, False), ("seg3", 3, 512, 200, True)])
def test_label_image_and_polygons_mirror(U, golden, golden_polygons, tag, seed, size, noise, use_body):
    masks, cls = segmentation_case(seed, size, noise)
    d = U.create_segmentations_masks(_Results(masks, cls, size))
    for name in ("bone", "muscles", "lung", "adipose"):
        assert np.array_equal(d[name][..., 0] | d[name][..., 1] | d[name][..., 2], golden[f"{tag}_cls_{name}"])
    body = None
    if use_body:
        ic = -1024 if seed % 2 == 0 else 0
        body = U.get_axial_slice_body_mask(DuckDataset(synth.phantom_slice(seed, ic, size=size), intercept=ic))
    color = U.create_color_output(d, body)
    assert np.array_equal(color, golden[f"{tag}_color"])
    assert U.create_list_crd_from_color_output(color, [0.753906, 0.753906], body) == golden_polygons[tag]


def test_errors_return_the_reference_sentinels(U):
    assert U.classic_norm(None) == []
    assert U.get_axial_slice_body_mask(object()) == []
    assert U.create_color_output(None) is None
    assert U.search_number_axial_slice(object()) == []


@pytest.fixture(scope="module")
def pipe():
    from eitsynthai_b200.pipeline import ImagingPipeline
    return ImagingPipeline("cuda:0")


def test_post_cnn_chain_matches_oracle(pipe):
    """CNN on the GPU; everything after it (NMS, mask decode, overlay, clean-up) vs the CPU oracle
    fed with the very same head / prototype tensors."""
    from eitsynthai_b200 import ops
    px = np.stack([synth.phantom_slice(s) for s in (11, 12, 13)])
    pxd = torch.from_numpy(px).cuda()
    body = ops.body_mask(pxd, 1, -1024, True)
    _, x = ops.hu_window(pxd, body_mask=body, want_u8=False, nchw_dtype=pipe.dtype)
    with torch.no_grad():
        head, protos = pipe._net(pipe.axial_model_512, x.contiguous(memory_format=torch.channels_last))
    code, body2, n = pipe.segment(pxd)
    assert torch.equal(body, body2)
    for b in range(3):
        wbody = O.body_mask(px[b], -1024, 1)
        assert np.array_equal(body[b].cpu().numpy(), wbody)
        r = Y.postprocess(head[b].float().cpu(), protos[b].float().cpu(), 4, (512, 512), (512, 512))
        assert int(n[b]) == r["dets"].shape[0] or int(n[b]) >= r["dets"].shape[0]     # the oracle drops empty masks
        union = O.class_union_masks(r["masks"].numpy(), r["cls"].numpy().astype(int), 512)
        want = O.create_color_codes(union, wbody)
        got = code[b].cpu().numpy()
        assert (got != want).mean() <= 1e-4, (b, int((got != want).sum()))
        for c in O.CODE_OF_CLASS:
            a, g = want == c, got == c
            if a.any():
                assert (a & g).sum() / (a | g).sum() >= 0.999


def _duck_series(n=48, seed=3):
    vol, inst = synth.phantom_series(n, seed=seed)
    return [DuckDataset(vol[i], instance_number=int(inst[i])) for i in range(n)], vol, inst


def test_entry_points(pipe):
    from eitsynthai_b200.kt_service.ai_tools import ai_tools as A
    slices, vol, inst = _duck_series()
    nodes, tri = synth.delaunay_mesh((60, 80, 450, 430), 6.0, seed=1)
    frame = A.DICOMToMask().get_coordinate_slice_from_dicom_frame(list(slices), mesh=(nodes, tri))
    assert frame["status"] == "success" and frame["label_codes"].shape == (512, 512)
    assert len(frame["mesh_data"]["CLASS"]) == len(tri) and set(frame["mesh_data"]["CLASS"]) <= {0, 1, 2, 3, 4}
    code, body, n = pipe.segment(torch.from_numpy(slices[-1].pixel_array[None]).cuda())
    assert np.array_equal(frame["label_codes"], code[0].cpu().numpy())
    # series route: random-init rib model rarely yields 7 right-side boxes -> the reference's [] sentinel or a full answer
    seq = A.DICOMSequencesToMask().get_coordinate_slice_from_dicom(list(slices), mesh=(nodes, tri))
    assert seq == [] or seq["status"] == "success"
    img = O.apply_mask(O.classic_norm(vol[0]), O.body_mask(vol[0], -1024, 1))
    ans = A.ImageToMask().get_coordinate_slice_from_image(img)
    assert ans["status"] == "success" and ans["mesh_data"] == []
    c2, _, _ = pipe.segment_u8(torch.from_numpy(img[None]).cuda())
    assert np.array_equal(ans["label_codes"], c2[0].cpu().numpy())
    # three equal channels are the same request; a coloured upload goes BGR -> RGB through the three-channel stem
    # (ai_tools.py:134): on a gray image both stems must agree up to the summation order of the 27 taps
    ans3 = A.ImageToMask().get_coordinate_slice_from_image(np.repeat(img[..., None], 3, axis=2))
    assert np.array_equal(ans3["label_codes"], ans["label_codes"])
    c3, _, _ = pipe.segment_bgr_u8(torch.from_numpy(np.repeat(img[None, ..., None], 3, axis=3)).cuda())
    assert (c3[0].cpu().numpy() != ans["label_codes"]).mean() <= 2e-3
    tinted = np.stack([img, img // 2, 255 - img], axis=2)
    ansc = A.ImageToMask().get_coordinate_slice_from_image(tinted)
    cc, _, _ = pipe.segment_bgr_u8(torch.from_numpy(tinted[None]).cuda())
    assert ansc["status"] == "success" and np.array_equal(ansc["label_codes"], cc[0].cpu().numpy())
    nii = A.NIIToMask().get_coordinate_slice_from_nii({"hu": synth.phantom_hu(4).astype(np.int16), "pixel_spacing": [0.7, 0.7]}, mesh=(nodes, tri))
    assert nii["status"] == "success" and nii["polygons"][0] == "0.7"
    assert A.DICOMToMask().get_coordinate_slice_from_dicom_frame(b"not a zip") == []


def test_run_series_matches_chunked_segment(pipe):
    vol, inst = synth.phantom_series(24, seed=9)
    from eitsynthai_b200.pipeline import SeriesMeta
    host_px = torch.from_numpy(vol).pin_memory()
    out_host = torch.empty((24, 512, 512), dtype=torch.uint8).pin_memory()
    res = pipe.run_series(host_px, SeriesMeta(inst), out_host, chunk=10)
    torch.cuda.synchronize()
    code, _, n = pipe.segment(torch.from_numpy(vol).cuda())
    assert torch.equal(res.labels, code) and np.array_equal(out_host.numpy(), code.cpu().numpy())
    srt = vol[np.argsort(inst, kind="stable")]
    assert np.array_equal(res.front_u8.cpu().numpy(), O.front_slice_norm(srt))


def test_cnn_fused_epilogues_match_plain_torch():
    """BN folding + K9 epilogues (bias/SiLU/residual/concat-slice) + fused upsample-concat vs the plain
    PyTorch module with the same weights (fp16 on both sides)."""
    from eitsynthai_b200.yolo_seg import YOLO11sSeg, build_model
    fused = build_model(4, "cuda:0", torch.float16, seed=5, fuse=True)
    plain = build_model(4, "cuda:0", torch.float16, seed=5, fuse=False)
    for mod in plain.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            assert float(mod.running_var.float().mean()) == 1.0
    x = torch.rand(3, 3, 256, 256, device="cuda").half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        h1, p1 = fused(x)
        h0, p0 = plain(x)
    for a, b in ((h1, h0), (p1, p0)):
        err = (a.float() - b.float()).abs().max().item()
        assert err <= 2e-2 * max(1.0, b.float().abs().max().item()), err


def test_conv_epilogue_and_upsample_concat_kernels():
    from eitsynthai_b200 import ops
    torch.manual_seed(0)
    cl = torch.channels_last
    y = torch.randn(2, 32, 20, 24, device="cuda").half().contiguous(memory_format=cl)
    res = torch.randn_like(y).contiguous(memory_format=cl)
    bias = torch.randn(32, device="cuda")
    buf = torch.zeros(2, 96, 20, 24, device="cuda", dtype=torch.half).contiguous(memory_format=cl)
    want = (torch.nn.functional.silu(y.float() + bias.view(1, -1, 1, 1)) + res.float())
    src = y.clone(memory_format=cl)
    out = ops.conv_epilogue(src, bias, True, res, True, buf, 32)
    assert torch.allclose(out.float(), want, atol=4e-3, rtol=2e-3)
    assert torch.equal(buf[:, 32:64], out) and float(buf[:, :32].abs().max()) == 0 and float(buf[:, 64:].abs().max()) == 0
    src2 = y.clone(memory_format=cl)
    assert ops.conv_epilogue(src2, None, False, None, False, buf, 64) is None
    assert torch.equal(buf[:, 64:], y) and torch.equal(src2, y)
    a = torch.randn(2, 16, 5, 7, device="cuda").half().contiguous(memory_format=cl)
    b = torch.randn(2, 24, 10, 14, device="cuda").half().contiguous(memory_format=cl)
    got = ops.upsample2x_concat(a, b)
    assert torch.equal(got, torch.cat((torch.nn.functional.interpolate(a, scale_factor=2.0, mode="nearest"), b), 1))


def test_mirror_overlay_and_mask_helpers_are_library_calls(golden):
    """The per-function helpers of the kt_service mirror (overlay of the class images, mask apply, class images from the
    instance masks) run as libeitb200 kernels and reproduce the real reference's outputs."""
    from eitsynthai_b200 import ops
    from eitsynthai_b200.kt_service.ai_tools import utils as U
    masks, cls = segmentation_case(1, 512, 60)
    d = O.create_segmentations_masks(masks, cls, 512)
    code = U._codes_from_class_images(d, torch.device("cuda:0")).cpu().numpy()
    assert np.array_equal(O.code_to_bgr(code), golden["seg1_overlay"])
    img, m = torch.from_numpy(golden["p0_norm"]).cuda(), torch.from_numpy(golden["p0_body"]).cuda()
    assert np.array_equal(U._masked(img, m).cpu().numpy(), golden["p0_normbody"])
    got = ops.class_images(torch.from_numpy(masks.astype(np.float32)).cuda(), torch.from_numpy(cls.astype(np.int32)).cuda()).cpu().numpy()
    for c, name in enumerate(U.CLASS_NAMES):
        assert np.array_equal(got[c], d[name]), name
    with pytest.raises(Exception):
        U._masked(torch.from_numpy(golden["p0_norm"]), torch.from_numpy(golden["p0_body"]))   # host tensors: no CPU fallback


def test_series_batch_runner_graphs_match_eager(pipe):
    """Public throughput engine: CUDA-graph replay and the host path give the same label maps as the
    plain per-chunk calls, and the same coronal decision as ImagingPipeline.rib_select."""
    from eitsynthai_b200.pipeline import SeriesBatchRunner, SeriesMeta
    vol, inst = synth.phantom_series(48, seed=21)
    px_host = torch.from_numpy(vol[None]).pin_memory()
    out_host = torch.zeros((1, 48, 512, 512), dtype=torch.uint8).pin_memory()
    r = SeriesBatchRunner(pipe, [SeriesMeta(inst)], 48, 512, chunk=32)
    r.load(px_host)
    r.capture(warm=1)
    sel_dev = r.step_device().cpu()
    want, _, _ = pipe.segment(torch.from_numpy(vol).cuda())
    got = torch.cat([o[0] for o in r.outs])
    assert torch.equal(got, want)
    sel_host = r.step_host(px_host, out_host)
    torch.cuda.synchronize()
    assert torch.equal(sel_host, sel_dev)
    assert np.array_equal(out_host[0].numpy(), want.cpu().numpy())
    front = pipe.coronal(torch.from_numpy(vol).cuda(), SeriesMeta(inst))
    sel_ref, _, _ = pipe.rib_select(front[None])
    assert torch.equal(sel_ref.cpu(), sel_dev)


def test_series_batch_runner_overlapped_passes(pipe):
    """The overlapped replay (K2 / K5-K7 graphs on side streams, passes joined only by buffer events) and the two-deep host
    path deliver, pass after pass, exactly what the single-stream runner delivers -- also when the pixels change between
    passes (a stale or early read of a hand-over buffer would show up as another pass's labels)."""
    from eitsynthai_b200.pipeline import SeriesBatchRunner, SeriesMeta
    base = synth.phantom_series(48, seed=31)[0]
    # shifted copies: the body outline (hence the label image) must differ between the passes
    vols = [base, np.roll(base, 37, axis=2).copy(), np.roll(base, -53, axis=1).copy()]
    inst = synth.phantom_series(48, seed=31)[1]
    plain = SeriesBatchRunner(pipe, [SeriesMeta(inst)], 48, 512, chunk=16, overlap=False)
    fast = SeriesBatchRunner(pipe, [SeriesMeta(inst)], 48, 512, chunk=16, overlap=True)
    assert fast.overlap and not plain.overlap
    hosts = [torch.from_numpy(v[None]).pin_memory() for v in vols]

    def input_sensitive(r):
        """The random-init network's label image barely depends on its input, which would hide a stale hand-over buffer:
        flip the sign of every prototype of a slice by a checksum of its window image (inside the captured graphs)."""
        plain_stage = r.cnn_stage

        def stage(x, out=None):
            head, protos = plain_stage(x, out=out)
            flip = (x.reshape(x.shape[0], -1)[:, ::997].sum(1, dtype=torch.int64) & 1) == 1   # parity of a strided checksum
            protos.mul_(torch.where(flip, 1.0, -1.0).to(protos.dtype)[:, None, None, None])
            return head, protos
        r.cnn_stage = stage
    for r in (plain, fast):
        input_sensitive(r)
        r.load(hosts[0])
        r.capture(warm=1)
    want, want_sel = [], []
    for h in hosts:
        out = torch.zeros((1, 48, 512, 512), dtype=torch.uint8).pin_memory()
        want_sel.append(plain.step_host(h, out))
        want.append(out)
    assert not torch.equal(want[0], want[1]) and not torch.equal(want[0], want[2])
    # device-resident passes without a join in between, then one join
    fast.load(hosts[0])
    for _ in range(3):
        sel = fast.step_device(join=False)
    fast.join()
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([o[0] for o in fast.outs]).cpu(), want[0][0])
    assert torch.equal(sel.cpu(), want_sel[0])
    # host passes, two in flight, different pixels every pass, two rounds over the three series
    outs = [torch.zeros((1, 48, 512, 512), dtype=torch.uint8).pin_memory() for _ in range(6)]
    handles = []
    for k in range(6):
        handles.append(fast.submit_host(hosts[k % 3], outs[k]))
        if k >= 1:
            got_sel = fast.wait_host(handles[k - 1])
            assert torch.equal(got_sel, want_sel[(k - 1) % 3])
            assert torch.equal(outs[k - 1], want[(k - 1) % 3]), f"pass {k - 1}"
    assert torch.equal(fast.wait_host(handles[5]), want_sel[2]) and torch.equal(outs[5], want[2])
    fast.join()
    torch.cuda.synchronize()
    # the label chains cut into parallel graph branches (label_fan), two chunk sizes in one pass
    fan = SeriesBatchRunner(pipe, [SeriesMeta(inst)], 48, 512, chunk=32, overlap=True, label_fan=3)
    input_sensitive(fan)
    fan.load(hosts[1])
    fan.capture(warm=1)
    out = torch.zeros((1, 48, 512, 512), dtype=torch.uint8).pin_memory()
    for k in (1, 2):
        assert torch.equal(fan.step_host(hosts[k], out), want_sel[k]) and torch.equal(out, want[k])


def test_zip_of_dicom_files_end_to_end(pipe):
    """Real wire format: a zip of (uncompressed) DICOM files through the reference-named entry points."""
    from eitsynthai_b200.kt_service.ai_tools import ai_tools as A
    from eitsynthai_b200.kt_service.ai_tools import dicom_io as D
    vol, inst = synth.phantom_series(40, seed=4)
    z = D.zip_series(vol, inst, custom=1)
    frame = A.DICOMToMask().get_coordinate_slice_from_dicom_frame(z)
    assert frame["status"] == "success"
    code, _, _ = pipe.segment(torch.from_numpy(vol[-1:].copy()).cuda())       # the last dataset of the archive
    assert np.array_equal(frame["label_codes"], code[0].cpu().numpy())
    obj = A.DICOMSequencesToMaskCustom()
    front, px, i_slices, custom = obj._search_front_slise(D.zip_series(vol, inst, custom=1))
    assert custom == 1 and np.array_equal(px, vol)
    assert [int(s.InstanceNumber) for s in i_slices] == sorted(int(i) for i in inst)
    assert np.array_equal(front, O.front_slice_norm(vol[np.argsort(inst, kind="stable")]))


def test_full_size_series_invariants(pipe):
    """BASELINE configs[2] at full size (320 slices): properties that need no CPU run of the whole series --
    the coronal image equals the (cheap) oracle, label maps do not depend on how the series is chunked or
    batched, body masks are per-slice functions, and a z-split of the series gives the same coronal rows."""
    from eitsynthai_b200 import ops
    from eitsynthai_b200.pipeline import SeriesBatchRunner, SeriesMeta
    vol, inst = synth.phantom_series(320, seed=0, shuffle_seed=17)
    meta = SeriesMeta(inst)
    dev_px = torch.from_numpy(vol).cuda()
    srt = vol[np.argsort(inst, kind="stable")]
    assert np.array_equal(pipe.coronal(dev_px, meta).cpu().numpy(), O.front_slice_norm(srt))
    r = SeriesBatchRunner(pipe, [meta], 320, 512, chunk=160)
    r.load(torch.from_numpy(vol[None]))
    r.capture(warm=1)
    sel = r.step_device().cpu()
    big = torch.cat([o[0] for o in r.outs])
    small = torch.cat([pipe.segment(dev_px[c:c + 64])[0] for c in range(0, 320, 64)])
    assert torch.equal(big, small)                                   # chunk / graph invariance of 84 M label pixels
    for k in (0, 131, 319):                                          # single-slice calls agree with the batch
        assert torch.equal(pipe.segment(dev_px[k:k + 1])[0][0], big[k])
        assert np.array_equal(ops.body_mask(dev_px[k:k + 1])[0].cpu().numpy(), O.body_mask(vol[k], -1024, 1))
    order = np.argsort(inst, kind="stable").astype(np.int32)
    lo, hi = order[:160], order[160:]                                # the two shards a 2-GPU run would hold
    rows = []
    for part in (lo, hi):
        sub = dev_px[torch.from_numpy(np.sort(part)).cuda().long()]
        sub_inst = inst[np.sort(part)]
        rr, mm = pipe.coronal(sub, SeriesMeta(sub_inst), return_rows=True)
        rows.append(rr)
    assert np.array_equal(torch.cat(rows).cpu().numpy(), O.front_rows(srt))
    assert sel.shape == (1, 4)


def test_post_cnn_chain_matches_oracle_256(pipe):
    """The 256-pixel route (ai_tools.py:138-146 picks the 256 model): every stage at the other supported size."""
    from eitsynthai_b200 import ops
    px = np.stack([synth.phantom_slice(s, size=256) for s in (21, 22)])
    pxd = torch.from_numpy(px).cuda()
    body = ops.body_mask(pxd, 1, -1024, True)
    _, x = ops.hu_window(pxd, body_mask=body, want_u8=False, nchw_dtype=pipe.dtype, channels_last=True)
    with torch.no_grad():
        head, protos = pipe._net(pipe.axial_model_256, x)
    assert head.shape == (2, 40, 1344) and protos.shape == (2, 32, 64, 64)
    code, body2, n = pipe.segment(pxd)
    for b in range(2):
        wbody = O.body_mask(px[b], -1024, 1)
        assert np.array_equal(body2[b].cpu().numpy(), wbody)
        r = Y.postprocess(head[b].float().cpu(), protos[b].float().cpu(), 4, (256, 256), (256, 256))
        union = O.class_union_masks(r["masks"].numpy(), r["cls"].numpy().astype(int), 256)
        want = O.create_color_codes(union, wbody)
        got = code[b].cpu().numpy()
        assert (got != want).mean() <= 1e-4, (b, int((got != want).sum()))


def test_nii_zip_upload_end_to_end(pipe):
    import io
    import zipfile
    from eitsynthai_b200.kt_service.ai_tools import ai_tools as A
    from eitsynthai_b200.kt_service.ai_tools import nifti_io as N
    hu = np.stack([synth.phantom_hu(s).astype(np.int16) for s in range(3)], axis=-1)      # [i, j, k]
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as zf:
        zf.writestr("vol.nii.gz", N.write_nifti(hu, pixdim=(0.7, 0.7, 1.0)))
    buf.seek(0)
    ans = A.NIIToMask().get_coordinate_slice_from_nii(buf)
    assert ans["status"] == "success" and ans["polygons"][0] == str(np.float32(0.7).item())
    want = A.NIIToMask().get_coordinate_slice_from_nii({"hu": np.ascontiguousarray(np.rot90(hu[:, :, 1], k=-1)),
                                                        "pixel_spacing": [0.7, 0.7]})
    assert np.array_equal(ans["label_codes"], want["label_codes"])


def test_fastapi_routes(pipe):
    """The five POST routes of main_kt_service.py (:33,50,69,88,127) over the GPU pipelines."""
    import io
    import zipfile
    from fastapi.testclient import TestClient
    from PIL import Image
    from eitsynthai_b200.kt_service import main_kt_service as M
    from eitsynthai_b200.kt_service.ai_tools import dicom_io as D
    from eitsynthai_b200.kt_service.ai_tools import nifti_io as N
    c = TestClient(M.app)
    vol, inst = synth.phantom_series(16, seed=8)
    z = D.zip_series(vol, inst).getvalue()
    r = c.post("/uploadDicomFrame", files={"file": ("s.zip", z, "application/zip")})
    assert r.status_code == 200 and r.json()["status"] == "success" and r.json()["label_shape"] == [512, 512]
    assert c.post("/uploadDicomFrame", files={"file": ("s.zip", b"junk", "application/zip")}).status_code == 400
    r = c.post("/uploadDicomSequence", files={"file": ("s.zip", z, "application/zip")})
    assert r.status_code == 200                   # random-init rib model: < 7 right-side ribs -> the reference's [] sentinel
    assert r.json() == [] or r.json()["status"] == "success"
    img = O.apply_mask(O.classic_norm(vol[0]), O.body_mask(vol[0], -1024, 1))
    buf, png = io.BytesIO(), io.BytesIO()
    Image.fromarray(img).save(png, format="PNG")
    with zipfile.ZipFile(buf, "w") as zf:
        zf.writestr("slice.png", png.getvalue())
    r = c.post("/uploadImageAxialSlice", files={"file": ("i.zip", buf.getvalue(), "application/zip")})
    assert r.status_code == 200 and r.json()["mesh_classes"] == []
    hu = np.stack([synth.phantom_hu(s).astype(np.int16) for s in range(3)], axis=-1)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as zf:
        zf.writestr("vol.nii.gz", N.write_nifti(hu))
    r = c.post("/uploadNII", files={"file": ("n.zip", buf.getvalue(), "application/zip")})
    assert r.status_code == 200 and len(r.json()["mesh_classes"]) > 100


def test_results_objects_carry_the_reference_fields_and_rebuild_the_fused_label_image():
    """``predict_results`` returns what the reference's ``_axial_slice_predict`` returns (ai_tools.py:129-158): an object
    with ``masks.data`` / ``boxes.cls`` / ``orig_shape``; the per-function path create_segmentations_masks ->
    create_color_output on it gives the label image of the fused K5 -> K6 -> K7 path, and ``Detections.from_ultralytics``
    exposes the fields search_number_axial_slice reads."""
    from eitsynthai_b200.kt_service.ai_tools import ai_tools as A
    from eitsynthai_b200.kt_service.ai_tools import utils as U
    from eitsynthai_b200.kt_service.ai_tools.results import Detections
    from oracle import imaging as OI
    svc = A.ImageToMask()
    px = synth.phantom_slice(2)
    u8 = OI.apply_mask(OI.classic_norm(px), OI.body_mask(px, -1024, 1))
    res, t = svc.predict_results(u8)
    assert res.orig_shape == (512, 512) and len(res) == len(res.boxes) > 0
    assert tuple(res.masks.data.shape) == (len(res), 512, 512) and tuple(res.boxes.data.shape) == (len(res), 6)
    det = Detections.from_ultralytics(res)
    assert det.xyxy.shape == (len(res), 4) and det.mask.dtype == bool and det.class_id.dtype.kind == "i"
    color = U.create_color_output(U.create_segmentations_masks(res))            # per-function path, no body mask
    code, _, n = svc.pipeline.segment_u8(torch.from_numpy(u8[None]).to(svc.device))
    assert np.array_equal(color, OI.code_to_bgr(code[0].cpu().numpy()))
