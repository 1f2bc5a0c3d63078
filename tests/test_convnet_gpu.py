"""K11/K12 (own tcgen05 convolutions) against plain PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


CASES = [
    # cin, cout, k, s, H, W, B, act, res
    (32, 64, 3, 2, 64, 64, 3, True, False),
    (16, 32, 3, 1, 40, 24, 2, True, True),        # 32-byte K rows, map smaller than / not a multiple of the tile
    (96, 128, 1, 1, 32, 32, 2, True, False),      # 64-byte K rows (3 chunks of 32)
    (128, 128, 3, 1, 13, 20, 3, True, True),      # rib-network geometry: tiles overhang the map
    (256, 512, 3, 2, 16, 16, 5, True, False),     # two N tiles
    (768, 512, 1, 1, 8, 8, 5, True, False),       # 8x8 maps: two images per tile, odd image count
    (128, 4, 1, 1, 16, 16, 2, False, False),      # nc outputs: padded to 16 MMA columns, 8 stored
    (128, 1, 1, 1, 13, 20, 1, False, False),
    (64, 64, 3, 1, 16, 16, 1, False, True),
]


@pytest.mark.parametrize("cin,cout,k,s,H,W,B,act,has_res", CASES)
def test_conv2d_matches_fp32(cin, cout, k, s, H, W, B, act, has_res):
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + k + s)
    x = torch.randn((B, H, W, cin), generator=g).half().to(dev)
    w = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).half().to(dev)
    bias = torch.randn((cout,), generator=g).to(dev)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    res = torch.randn((B, Ho, Wo, cout), generator=g).half().to(dev) if has_res else None
    y = conv(Act(x), PackedConv.from_weight(w, bias, s, 1, act), res=Act(res) if has_res else None)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, s, k // 2)
    if act:
        ref = F.silu(ref)
    if has_res:
        ref = ref + res.permute(0, 3, 1, 2).float()
    got = y.buf[..., :cout].permute(0, 3, 1, 2).float()
    assert got.shape == ref.shape
    tol = 2e-3 * float(ref.abs().max()) + 2e-3                     # fp16 output rounding
    assert float((got - ref).abs().max()) <= tol


@pytest.mark.parametrize("cin,cout,H,W,B,act", [(256, 256, 32, 32, 3, True), (256, 128, 64, 64, 2, True), (64, 32, 26, 40, 2, True),
                                               (512, 256, 8, 8, 5, False), (128, 64, 4, 6, 3, True)])
def test_conv1x1_with_upsampled_addend(cin, cout, H, W, B, act):
    """res_mode 2: y = act(conv1x1(x) + bias + up2(pre)) == Conv1x1(Concat(Upsample(a), x)) with pre = Wa a."""
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(cin + cout + H)
    ca = 2 * cin                                                   # channels of the low-resolution input
    a = torch.randn((B, H // 2, W // 2, ca), generator=g).half().to(dev)
    x = torch.randn((B, H, W, cin), generator=g).half().to(dev)
    w = (torch.randn((cout, ca + cin, 1, 1), generator=g) / (ca + cin) ** 0.5).half().to(dev)
    bias = torch.randn((cout,), generator=g).to(dev)
    yb = torch.full((B, H, W, cout + 16), 3.0, device=dev).half()
    t = conv(Act(a), PackedConv.from_weight(w[:, :ca].contiguous(), None, 1, 1, False))
    conv(Act(x), PackedConv.from_weight(w[:, ca:].contiguous(), bias, 1, 1, act), out=Act(yb, 8, cout), pre=t)
    cat = torch.cat([F.interpolate(a.permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest"), x.permute(0, 3, 1, 2).float()], 1)
    ref = F.conv2d(cat, w.float(), bias)
    if act:
        ref = F.silu(ref)
    got = yb[..., 8:8 + cout].permute(0, 3, 1, 2).float()
    # the low-resolution partial sum is rounded to fp16 once more than in the unsplit convolution
    assert float((got - ref).abs().max()) <= 3e-3 * float(ref.abs().max()) + 3e-3
    assert bool((yb[..., :8] == 3).all()) and bool((yb[..., 8 + cout:] == 3).all())


def test_conv2d_reads_and_writes_channel_slices():
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    torch.manual_seed(3)
    xb = torch.randn((2, 32, 32, 192), device=dev).half()
    w = (torch.randn((64, 64, 3, 3), device=dev) / 24).half()
    bias = torch.randn((64,), device=dev)
    yb = torch.full((2, 32, 32, 160), 7.0, device=dev).half()
    conv(Act(xb, 64, 64), PackedConv.from_weight(w, bias, 1, 1, True), out=Act(yb, 32, 64))
    ref = F.silu(F.conv2d(xb[..., 64:128].permute(0, 3, 1, 2).float(), w.float(), bias, 1, 1)).permute(0, 2, 3, 1)
    assert float((yb[..., 32:96].float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 2e-3
    assert bool((yb[..., :32] == 7).all()) and bool((yb[..., 96:] == 7).all())       # neighbours untouched


def test_conv_transpose_phases():
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    torch.manual_seed(4)
    x = torch.randn((2, 16, 16, 128), device=dev).half()
    up = torch.nn.ConvTranspose2d(128, 128, 2, 2, 0, bias=True).to(dev).half()
    out = torch.empty((2, 32, 32, 128), device=dev, dtype=torch.float16)
    for dy in range(2):
        for dx in range(2):
            L = PackedConv.from_weight(up.weight[:, :, dy, dx].t().contiguous()[:, :, None, None], up.bias, 1, 1, False)
            conv(Act(x), L, out=Act(out), up=(2, dy, dx))
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2).float(), up.weight.float(), up.bias.float(), 2).permute(0, 2, 3, 1)
    assert float((out.float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 2e-3


def test_depthwise_and_stem():
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    torch.manual_seed(5)
    x = torch.randn((2, 20, 13, 128), device=dev).half()
    w = (torch.randn((128, 1, 3, 3), device=dev) / 3).half()
    b = torch.randn((128,), device=dev)
    y = conv(Act(x), PackedConv.from_weight(w, b, 1, 128, True))
    ref = F.silu(F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, 1, 1, 1, 128)).permute(0, 2, 3, 1)
    assert float((y.buf.float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 2e-3
    xs = torch.rand((2, 37, 64, 3), device=dev).half()
    ws = (torch.randn((32, 3, 3, 3), device=dev) / 5).half()
    bs = torch.randn((32,), device=dev)
    ys = conv(Act(xs), PackedConv.from_weight(ws, bs, 2, 1, True))
    refs = F.silu(F.conv2d(xs.permute(0, 3, 1, 2).float(), ws.float(), bs, 2, 1)).permute(0, 2, 3, 1)
    assert ys.buf.shape == refs.shape
    assert float((ys.buf.float() - refs).abs().max()) <= 2e-3 * float(refs.abs().max()) + 2e-3
    # replicated gray input: the 9-tap form with channel-summed weights
    xg = xs[..., :1].expand(-1, -1, -1, 3).contiguous()
    yg = conv(Act(xg), PackedConv.from_weight(ws, bs, 2, 1, True), gray=True)
    refg = F.silu(F.conv2d(xg.permute(0, 3, 1, 2).float(), ws.float(), bs, 2, 1)).permute(0, 2, 3, 1)
    assert float((yg.buf.float() - refg).abs().max()) <= 2e-3 * float(refg.abs().max()) + 2e-3
    # the u8 image itself: u8 -> fp16 / 255 -> three equal channels fused into the stem's tile loader; odd sizes, tile edges
    from eitsynthai_b200.convnet import stem_u8
    for (h, w_) in ((37, 64), (512, 512), (208, 130)):
        u8 = torch.randint(0, 256, (3, h, w_), device=dev, dtype=torch.uint8)
        yu = stem_u8(u8, PackedConv.from_weight(ws, bs, 2, 1, True))
        xin = (u8.float() / 255).half().float()[:, None].expand(-1, 3, -1, -1)
        refu = F.silu(F.conv2d(xin, ws.float(), bs, 2, 1)).permute(0, 2, 3, 1)
        assert yu.buf.shape == refu.shape
        assert float((yu.buf.float() - refu).abs().max()) <= 2e-3 * float(refu.abs().max()) + 2e-3
        # the CUDA-core form of the same kernel (the default runs the contraction as warp-level MMAs), and no activation
        from eitsynthai_b200 import cabi
        cabi.load().eitb_stem_debug(1)
        try:
            ysc = stem_u8(u8, PackedConv.from_weight(ws, bs, 2, 1, True))
        finally:
            cabi.load().eitb_stem_debug(0)
        assert float((ysc.buf.float() - refu).abs().max()) <= 2e-3 * float(refu.abs().max()) + 2e-3
        assert float((ysc.buf.float() - yu.buf.float()).abs().max()) <= 2e-3 * float(refu.abs().max()) + 2e-3
        yl = stem_u8(u8, PackedConv.from_weight(ws, bs, 2, 1, False))
        refl = F.conv2d(xin, ws.float(), bs, 2, 1).permute(0, 2, 3, 1)
        assert float((yl.buf.float() - refl).abs().max()) <= 1e-3 * float(refl.abs().max()) + 1e-3


@pytest.mark.parametrize("C,H,W,B", [(128, 64, 64, 3), (256, 16, 16, 5), (64, 33, 17, 2), (512, 13, 20, 2), (24, 9, 31, 2)])
def test_depthwise_tma_tiles_slices_and_fallback(C, H, W, B):
    """K12 depthwise: the TMA-staged kernel (C % 64 == 0) on tile edges, odd maps and channel slices of wider buffers;
    other channel counts take the direct kernel."""
    from eitsynthai_b200.convnet import Act, PackedConv, conv
    dev = _dev()
    torch.manual_seed(C + H)
    xb = torch.randn((B, H, W, C + 16), device=dev).half()
    w = (torch.randn((C, 1, 3, 3), device=dev) / 3).half()
    b = torch.randn((C,), device=dev)
    yb = torch.full((B, H, W, C + 24), 5.0, device=dev).half()
    conv(Act(xb, 8, C), PackedConv.from_weight(w, b, 1, C, True), out=Act(yb, 16, C))
    ref = F.silu(F.conv2d(xb[..., 8:8 + C].permute(0, 3, 1, 2).float(), w.float(), b, 1, 1, 1, C)).permute(0, 2, 3, 1)
    assert float((yb[..., 16:16 + C].float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 2e-3
    assert bool((yb[..., :16] == 5).all()) and bool((yb[..., 16 + C:] == 5).all())


@pytest.mark.parametrize("nc,size", [(4, (512, 512)), (4, (256, 256)), (1, (416, 640))])
def test_network_matches_fp32_pytorch(nc, size):
    """The whole YOLO11s-seg forward on own kernels against the same weights run by PyTorch in fp32."""
    from eitsynthai_b200.convnet import ConvNet
    from eitsynthai_b200.yolo_seg import build_model
    dev = _dev()
    m16 = build_model(nc, dev, torch.float16, seed=11)
    m32 = build_model(nc, dev, torch.float32, seed=11)
    net = ConvNet(m16)
    torch.manual_seed(1)
    x = (torch.randint(0, 256, (2, 1, *size), device=dev).float() / 255).expand(-1, 3, -1, -1)
    x16 = x.half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        head32, proto32 = m32(x16.float())
        head16, proto16 = m16(x16)                                 # cuDNN fp16 + K9 (round-1 path)
        head, proto = net(x16, gray=True)
        head_c, proto_c = net(x16)                                 # generic 27-tap stem: same result up to fp32 summation order
        u8 = (x[:, 0] * 255).round().to(torch.uint8).contiguous()
        head_u, proto_u = net(u8)                                  # preprocess fused into the stem (the production path)
    assert head.shape == head32.shape and proto.shape == proto32.shape
    assert float((proto_c.float() - proto.float()).abs().max()) <= 5e-3 * float(proto32.abs().max())
    assert float((proto_u.float() - proto.float()).abs().max()) <= 5e-3 * float(proto32.abs().max())
    assert float((head_u.float() - head.float()).abs().max()) <= 5e-3 * float(head32.abs().max())

    def rel(a, b):
        return float((a.float() - b).abs().max() / b.abs().max().clamp_min(1e-6))
    # boxes, scores, coefficients and prototypes within the error of an fp16 network against fp32,
    # and no worse than twice what cuDNN fp16 gives on the same weights
    for name, got, base, ref in (("box", head[:, :4], head16[:, :4], head32[:, :4]),
                                 ("cls", head[:, 4:4 + nc], head16[:, 4:4 + nc], head32[:, 4:4 + nc]),
                                 ("coef", head[:, 4 + nc:], head16[:, 4 + nc:], head32[:, 4 + nc:]),
                                 ("proto", proto, proto16, proto32)):
        e_own, e_cudnn = rel(got, ref), rel(base, ref)
        assert e_own <= max(2.0 * e_cudnn, 5e-3), (name, e_own, e_cudnn)
