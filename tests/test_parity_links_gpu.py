"""Links of the hot path that round 1 left without a hardware parity test ("next round" item 1):
the whole rib chain, eitb_scale_boxes, the K10 kernels in isolation, and the N-rank sharded run against the
single-rank run on NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from eitsynthai_b200 import synth
from oracle import imaging as O
from oracle import yolo_post as Y

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------ teacher-forced rib head
def rib_teacher_head(net_hw, front_hw, n_ribs=12, seed=0, dup=3):
    """A rib-model head [1, 4+1+32, A] for a letterboxed coronal image: ``n_ribs`` boxes right and left of the
    midline at known rows of the ORIGINAL image (plus jittered duplicates for NMS to remove), everything else
    below the confidence threshold."""
    rng = np.random.default_rng(seed)
    nh, nw = net_hw
    H0, W0 = front_hw
    A = sum((nh // s) * (nw // s) for s in (8, 16, 32))
    head = np.zeros((37, A), np.float32)
    head[0] = rng.uniform(0, nw, A); head[1] = rng.uniform(0, nh, A)
    head[2:4] = rng.uniform(4, 40, (2, A))
    head[4] = rng.uniform(0.0, 0.25, A)
    head[5:] = rng.normal(0, 0.3, (32, A))
    gain = min(nh / H0, nw / W0)
    pad_x, pad_y = round((nw - W0 * gain) / 2 - 0.1), round((nh - H0 * gain) / 2 - 0.1)
    slots = rng.permutation(A)
    k = 0
    ys = np.linspace(0.12 * H0, 0.88 * H0, n_ribs) + rng.uniform(-3, 3, n_ribs)
    for r, y0 in enumerate(ys):
        for side in (-1, 1):
            cx0 = W0 / 2 + side * (110 + 5 * r) + rng.uniform(-2, 2)
            for d in range(dup):
                a = slots[k]; k += 1
                j = rng.normal(0, 0.8, 4) if d else np.zeros(4)
                w0, h0 = 42 + j[2], 14 + j[3]
                head[0, a] = (cx0 + j[0]) * gain + pad_x
                head[1, a] = (y0 + j[1]) * gain + pad_y
                head[2, a], head[3, a] = w0 * gain, h0 * gain
                head[4, a] = rng.uniform(0.5, 0.95) - 0.1 * d
    return head[None]


class _Teacher:
    def __init__(self, head):
        self.head = head

    def __call__(self, x, gray=False):
        return self.head.to(x.device), None


@pytest.fixture(scope="module")
def pipe():
    from eitsynthai_b200.pipeline import ImagingPipeline
    return ImagingPipeline("cuda:0")


@pytest.mark.parametrize("n_slices,custom", [(320, 0), (320, 7), (200, -3)])
def test_rib_chain_matches_oracle(pipe, n_slices, custom):
    """letterbox -> head -> K5 -> eitb_scale_boxes -> K4 on the B200 against oracle.cpu_path.rib_select_cpu fed with
    the same head (reference ai_tools.py:107-127, utils.py:166-269): boxes and the selected index bit for bit."""
    from oracle import cpu_path
    vol, inst = synth.phantom_series(n_slices, seed=2, shuffle_seed=None, size=512)
    front = O.front_slice_norm(vol)
    nh, nw, top, bottom, left, right = Y.letterbox_geometry(n_slices, 512, 640)
    net_hw = (nh + top + bottom, nw + left + right)
    head = torch.from_numpy(rib_teacher_head(net_hw, (n_slices, 512), seed=n_slices + custom))
    want_sel, _ = cpu_path.rib_select_cpu(vol, _Teacher(head), custom)
    dets_cpu, _ = Y.nms(head[0], 1)
    want_boxes = Y.scale_boxes(net_hw, dets_cpu[:, :4], (n_slices, 512)).numpy()
    assert len(want_sel) == 3, "the teacher head must give >= 7 right-side ribs"

    saved = pipe.ribs_model
    pipe.ribs_model = _Teacher(head.cuda())
    try:
        fr = torch.from_numpy(front).cuda()
        cus = torch.tensor([custom], dtype=torch.int32, device="cuda")
        sel, boxes, k = pipe.rib_select(fr[None], cus)
    finally:
        pipe.ribs_model = saved
    n = int(k[0])
    assert n == want_boxes.shape[0]
    assert np.array_equal(boxes[0, :n].cpu().numpy(), want_boxes)             # same floats, same order
    assert bool((boxes[0, n:] == 0).all())
    got = sel[0].cpu().tolist()
    assert got[3] == 1 and got[:3] == want_sel


def test_rib_chain_too_few_ribs_gives_sentinel(pipe):
    head = torch.from_numpy(rib_teacher_head((416, 640), (320, 512), n_ribs=5, seed=3)).cuda()
    saved = pipe.ribs_model
    pipe.ribs_model = _Teacher(head)
    try:
        front = torch.zeros((1, 320, 512), dtype=torch.uint8, device="cuda")
        sel, _, _ = pipe.rib_select(front)
    finally:
        pipe.ribs_model = saved
    assert sel[0, 3].item() == 0                                  # < 7 right-side boxes: the reference returns []


@pytest.mark.parametrize("net,orig", [((416, 640), (320, 512)), ((640, 448), (600, 400)), ((512, 512), (512, 512)),
                                       ((640, 640), (333, 777)), ((256, 640), (97, 512))])
def test_scale_boxes_matches_oracle(net, orig):
    """eitb_scale_boxes against ultralytics scale_boxes restated in oracle.yolo_post: non-square shapes, boxes
    outside the image (clamping), rows >= n zeroed."""
    from eitsynthai_b200 import host, ops
    rng = np.random.default_rng(net[0] + orig[1])
    B, max_det, D = 3, 40, 38
    dets = rng.uniform(-60, max(net) + 60, (B, max_det, D)).astype(np.float32)
    n = np.asarray([40, 17, 0], np.int32)
    gain, pad_x, pad_y = host.scale_boxes_params(net, orig)
    got = ops.scale_boxes(torch.from_numpy(dets).cuda(), torch.from_numpy(n).cuda(), gain, pad_x, pad_y, orig[1], orig[0]).cpu().numpy()
    for b in range(B):
        want = Y.scale_boxes(net, torch.from_numpy(dets[b, :n[b], :4]), orig).numpy()
        assert np.array_equal(got[b, :n[b]], want)
        assert not got[b, n[b]:].any()


# ------------------------------------------------------------------------------------ K10 in isolation
@pytest.mark.parametrize("nc,hw", [(4, (64, 64)), (1, (52, 80)), (4, (32, 32))])
def test_head_decode_kernel_matches_fp32(nc, hw):
    """DFL softmax expectation, dist2bbox, stride scaling, sigmoid and concat of eitb_yolo_head_decode against plain
    PyTorch fp32 on the same fp16 branch outputs (ultralytics Detect._inference)."""
    from eitsynthai_b200 import ops
    torch.manual_seed(nc * 100 + hw[0])
    dev = torch.device("cuda:0")
    B, nm = 3, 32
    shapes = [(hw[0] >> i, hw[1] >> i) for i in range(3)]
    cl = torch.channels_last
    box = [(torch.randn((B, 64, h, w), device=dev) * 2).half().contiguous(memory_format=cl) for h, w in shapes]
    cls = [(torch.randn((B, nc, h, w), device=dev) * 2).half().contiguous(memory_format=cl) for h, w in shapes]
    mc = [torch.randn((B, nm, h, w), device=dev).half().contiguous(memory_format=cl) for h, w in shapes]
    biases = tuple([torch.randn((c,), device=dev) for _ in shapes] for c in (64, nc, nm))
    head = ops.yolo_head_decode(box, cls, mc, (8, 16, 32), nc, nm, biases).float()
    # fp32 reference
    refs = []
    for i, (h, w) in enumerate(shapes):
        b = box[i].float() + biases[0][i].view(1, -1, 1, 1)
        dist = (b.view(B, 4, 16, h * w).softmax(2) * torch.arange(16, device=dev).view(1, 1, 16, 1)).sum(2)
        sy, sx = torch.meshgrid(torch.arange(h, device=dev) + 0.5, torch.arange(w, device=dev) + 0.5, indexing="ij")
        anc = torch.stack((sx.flatten(), sy.flatten()))[None]
        x1y1, x2y2 = anc - dist[:, :2], anc + dist[:, 2:]
        xywh = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * (8 << i)
        c = (cls[i].float() + biases[1][i].view(1, -1, 1, 1)).flatten(2).sigmoid()
        m = (mc[i].float() + biases[2][i].view(1, -1, 1, 1)).flatten(2)
        refs.append(torch.cat((xywh, c, m), 1))
    ref = torch.cat(refs, 2)
    assert head.shape == ref.shape
    # the head is stored in fp16: half an ulp of the value
    tol = ref.abs() * 1.0e-3 + 2e-3
    assert bool(((head - ref).abs() <= tol).all()), float(((head - ref).abs() - tol).max())


@pytest.mark.parametrize("hw,C", [((16, 16), 256), ((13, 20), 256), ((8, 8), 64)])
def test_sppf_kernel_matches_pytorch(hw, C):
    """x | maxpool5 | maxpool5^2 | maxpool5^3 of eitb_sppf_pool_concat: max is exact in fp16 -> bit-exact."""
    from eitsynthai_b200 import ops
    torch.manual_seed(hw[0])
    x = torch.randn((3, C, *hw), device="cuda").half().contiguous(memory_format=torch.channels_last)
    got = ops.sppf_pool_concat(x)
    y = [x.float()]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    assert torch.equal(got.float(), torch.cat(y, 1))


# ------------------------------------------------------------------------------------ sharded == single rank, NCCL
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, n_slices, S, out_dir, flip):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from eitsynthai_b200 import sharded
        from eitsynthai_b200.pipeline import ImagingPipeline, SeriesBatchRunner, SeriesMeta
        pipe = ImagingPipeline(dev, seed=0)
        z0, z1 = sharded.shard_range(n_slices, world, rank)
        vols, metas = [], []
        for s in range(S):
            v, i = synth.phantom_series(n_slices, seed=s, shuffle_seed=3 + s, z_range=(z0, z1))
            vols.append(v)
            metas.append(SeriesMeta(i, patient_position="FFS" if (flip and s == 1) else "HFS"))
        runner = SeriesBatchRunner(pipe, metas, n_slices, 512, chunk=16, use_graphs=False)
        runner.load(torch.from_numpy(np.stack(vols)))
        sel = runner.step_eager()
        codes = torch.cat([runner.slice_stage(runner.flat[a:b], a, b)[0] for a, b in runner.bounds])
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), sel=sel.cpu().numpy(), codes=codes.cpu().numpy(), z=np.asarray([z0, z1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("flip", [False, True])
def test_two_rank_nccl_equals_single_rank(tmp_path, flip):
    """The slice-sharded run on 2 GPUs over NCCL gives the same selected indices and the same label maps as one
    rank on the whole series; with an FFS series in the batch (z reversed after the gather)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from eitsynthai_b200.pipeline import ImagingPipeline, SeriesBatchRunner, SeriesMeta
    n_slices, S = 48, 2
    mp.spawn(_nccl_worker, args=(2, _free_port(), n_slices, S, str(tmp_path), flip), nprocs=2, join=True)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    pipe = ImagingPipeline("cuda:0", seed=0)
    vols, metas = [], []
    for s in range(S):
        v, i = synth.phantom_series(n_slices, seed=s, shuffle_seed=None)
        vols.append(v)
        metas.append(SeriesMeta(i, patient_position="FFS" if (flip and s == 1) else "HFS"))
    runner = SeriesBatchRunner(pipe, metas, n_slices, 512, chunk=16, use_graphs=False)
    runner.load(torch.from_numpy(np.stack(vols)))
    sel = runner.step_eager().cpu().numpy()
    assert np.array_equal(got[0]["sel"], sel) and np.array_equal(got[1]["sel"], sel)
    # label maps: rank r holds slices [z0, z1) of every series, in its own (shuffled) file order -> compare as sets per z
    for r in range(2):
        z0, z1 = got[r]["z"]
        nl = z1 - z0
        for s in range(S):
            _, inst = synth.phantom_series(n_slices, seed=s, shuffle_seed=3 + s, z_range=(int(z0), int(z1)))
            codes_r = got[r]["codes"][s * nl:(s + 1) * nl]
            one, _ = pipe.segment(torch.from_numpy(vols[s][z0:z1]).cuda())[0:2]
            one = one.cpu().numpy()
            for k, inum in enumerate(inst):
                assert np.array_equal(codes_r[k], one[int(inum) - 1 - z0])
