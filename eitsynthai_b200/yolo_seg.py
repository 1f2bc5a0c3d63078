"""YOLO11s-seg in plain PyTorch (random-init): the CNN between K1 and K5.

The reference reaches the network only through ``ultralytics.YOLO(path, task='segment')``
(kt_service/ai_tools/ai_tools.py:69-71, 121-122, 153); ultralytics is not installed here and the
weights are unavailable offline, so the architecture is restated from the published
``yolo11-seg.yaml`` at scale ``s`` and initialised randomly.  It runs
through cuDNN (channels-last, half precision); it is PyTorch-owned and not one of the
hand-written kernels.  Output: ``head [B, 4+nc+32, A]`` (xywh in input pixels, sigmoid class
scores, mask coefficients) and ``protos [B, 32, S/4, S/4]`` -- the operands of K5 and K6.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv(nn.Module):
    """Conv2d(bias=False) + BatchNorm2d + SiLU.  ``fuse()`` folds the BatchNorm into the convolution
    (what ultralytics does before predicting); on a CUDA half tensor the folded bias and the SiLU
    then run as one in-place libeitb200 pass (K9) right after cuDNN's convolution."""

    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True) if act else nn.Identity()
        self.has_act = act
        self.fused_bias = None

    @torch.no_grad()
    def fuse(self):
        bn = self.bn
        scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        self.conv.weight.data = (self.conv.weight.float() * scale.view(-1, 1, 1, 1)).to(self.conv.weight.dtype)
        self.fused_bias = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
        self.bn = nn.Identity()

    def forward(self, x, residual=None, out=None, out_off=0, keep=True):
        """``residual`` is added after the activation (Bottleneck shortcut); ``out``/``out_off`` also place
        the result in a channel slice of a wider tensor (concat buffer); ``keep=False`` skips the
        stand-alone result when only the slice is needed (returns None)."""
        if self.fused_bias is None:
            y = self.act(self.bn(self.conv(x)))
        else:
            y = self.conv(x)
            if (y.is_cuda and y.dtype != torch.float32 and y.shape[1] % 8 == 0
                    and y.is_contiguous(memory_format=torch.channels_last)
                    and (out is None or (out.shape[1] % 8 == 0 and out_off % 8 == 0))):
                from . import ops
                if residual is None and out is None:
                    return ops.bias_act_(y, self.fused_bias, self.has_act)
                return ops.conv_epilogue(y, self.fused_bias, self.has_act, residual, keep, out, out_off)
            y = self.act(y + self.fused_bias.to(y.dtype).view(1, -1, 1, 1))
        if residual is not None:
            y = y + residual
        if out is not None:
            out[:, out_off:out_off + y.shape[1]] = y
        return y if keep else None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if self.fused_bias is not None:
            self.fused_bias = fn(self.fused_bias).float()
        return self


def _split_conv(cv: "Conv", c: int):
    """Split a fused Conv with 2c output channels into two Convs with c channels each."""
    assert cv.fused_bias is not None, "fuse() first"
    out = []
    for sl in (slice(0, c), slice(c, 2 * c)):
        k = cv.conv.kernel_size[0]
        n = Conv(cv.conv.in_channels, c, k, cv.conv.stride[0], cv.conv.groups, cv.has_act)
        n.conv.weight.data = cv.conv.weight.data[sl].clone()
        n.bn = nn.Identity()
        n.fused_bias = cv.fused_bias[sl].clone()
        out.append(n)
    return out


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0])
        self.cv2 = Conv(c_, c2, k[1])
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None, out_off=0, keep=True):
        return self.cv2(self.cv1(x), residual=x if self.add else None, out=out, out_off=out_off, keep=keep)


class C3k(nn.Module):
    def __init__(self, c1, c2, n=2, shortcut=True, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1)
        self.cv2 = Conv(c1, c_, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, (k, k), 1.0) for _ in range(n)))

    def forward(self, x, out=None, out_off=0, keep=True):
        c_ = self.cv1.conv.out_channels
        buf = torch.empty((x.shape[0], 2 * c_, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                          memory_format=torch.channels_last)
        h = self.cv1(x)
        for i, m in enumerate(self.m):
            last = i == len(self.m) - 1
            h = m(h, out=buf if last else None, out_off=0, keep=not last)
        self.cv2(x, out=buf, out_off=c_, keep=False)
        return self.cv3(buf, out=out, out_off=out_off, keep=keep)


class C3k2(nn.Module):
    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(C3k(self.c, self.c, 2, shortcut) if c3k else Bottleneck(self.c, self.c, shortcut, (3, 3), 0.5)
                               for _ in range(n))

    def split_cv1(self):
        """Two convolutions instead of conv + channel split: the halves come out contiguous, so the
        bottleneck that reads the second half (and adds it back) needs no strided copy."""
        self.cv1a, self.cv1b = _split_conv(self.cv1, self.c)
        del self.cv1

    def forward(self, x):
        if not hasattr(self, "cv1a"):
            y = list(self.cv1(x).chunk(2, 1))
            y.extend(m(y[-1]) for m in self.m)
            return self.cv2(torch.cat(y, 1))
        # every producer writes its channel slice of the concat buffer from its own epilogue
        c, n = self.c, len(self.m)
        buf = torch.empty((x.shape[0], (2 + n) * c, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                          memory_format=torch.channels_last)
        self.cv1a(x, out=buf, out_off=0, keep=False)
        h = self.cv1b(x, out=buf, out_off=c)
        for i, m in enumerate(self.m):
            last = i == n - 1
            h = m(h, out=buf, out_off=(2 + i) * c, keep=not last)
        return self.cv2(buf)


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1)
        self.cv2 = Conv(c_ * 4, c2, 1)
        self.m = nn.MaxPool2d(k, 1, k // 2)

    def forward(self, x):
        y0 = self.cv1(x)
        if getattr(self, "fused_tails", False) and y0.is_cuda and y0.dtype == torch.float16 and y0.is_contiguous(memory_format=torch.channels_last) \
                and y0.shape[1] % 8 == 0 and y0.shape[2] * y0.shape[3] <= 2048 and self.m.kernel_size == 5:
            from . import ops
            return self.cv2(ops.sppf_pool_concat(y0))
        y = [y0]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class Attention(nn.Module):
    def __init__(self, dim, num_heads, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        self.qkv = Conv(dim, dim + self.key_dim * num_heads * 2, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x).reshape(B, self.num_heads, self.key_dim * 2 + self.head_dim, N)
        q, k, v = qkv.split([self.key_dim, self.key_dim, self.head_dim], dim=2)
        attn = (q.transpose(-2, -1) @ k) * self.scale
        attn = attn.softmax(dim=-1)
        o = (v @ attn.transpose(-2, -1)).reshape(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        return self.proj(o)


class PSABlock(nn.Module):
    def __init__(self, c, attn_ratio=0.5, num_heads=4):
        super().__init__()
        self.attn = Attention(c, num_heads, attn_ratio)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))

    def forward(self, x):
        x = x + self.attn(x)
        return x + self.ffn(x)


class C2PSA(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, 0.5, self.c // 64) for _ in range(n)))

    def split_cv1(self):
        self.cv1a, self.cv1b = _split_conv(self.cv1, self.c)
        del self.cv1

    def forward(self, x):
        a, b = (self.cv1a(x), self.cv1b(x)) if hasattr(self, "cv1a") else self.cv1(x).split((self.c, self.c), dim=1)
        return self.cv2(torch.cat((a, self.m(b)), 1))


class Proto(nn.Module):
    def __init__(self, c1, c_=256, c2=32):
        super().__init__()
        self.cv1 = Conv(c1, c_, 3)
        self.upsample = nn.ConvTranspose2d(c_, c_, 2, 2, 0, bias=True)
        self.cv2 = Conv(c_, c_, 3)
        self.cv3 = Conv(c_, c2, 1)

    def strip_bias(self):
        """Serve the transposed convolution's bias through K9 instead of PyTorch's broadcast add."""
        self.up_bias = self.upsample.bias.detach().float().clone()
        self.upsample.bias = None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if getattr(self, "up_bias", None) is not None:
            self.up_bias = fn(self.up_bias).float()
        return self

    def forward(self, x):
        y = self.upsample(self.cv1(x))
        ub = getattr(self, "up_bias", None)
        if ub is not None:
            if y.is_cuda and y.dtype != torch.float32 and y.is_contiguous(memory_format=torch.channels_last):
                from . import ops
                ops.bias_act_(y, ub, False)
            else:
                y = y + ub.to(y.dtype).view(1, -1, 1, 1)
        return self.cv3(self.cv2(y))


class Segment(nn.Module):
    """YOLO11 Segment head (Detect with DFL + mask coefficients + prototypes), inference form."""
    reg_max = 16

    def __init__(self, nc, ch, nm=32, npr=128):
        super().__init__()
        self.nc, self.nm, self.nl = nc, nm, len(ch)
        self.stride = (8, 16, 32)
        c2 = max(16, ch[0] // 4, self.reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        c4 = max(ch[0] // 4, nm)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(Conv(x, x, 3, g=x), Conv(x, c3, 1)),
                                               nn.Sequential(Conv(c3, c3, 3, g=c3), Conv(c3, c3, 1)),
                                               nn.Conv2d(c3, nc, 1)) for x in ch)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, nm, 1)) for x in ch)
        self.proto = Proto(ch[0], npr, nm)
        self.register_buffer("bins", torch.arange(self.reg_max, dtype=torch.float32).view(1, 1, self.reg_max, 1), persistent=False)
        self._anchors = {}

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)

    def strip_bias(self):
        """The last convolution of every branch runs bias-free; the biases are added inside the K10 decode
        (PyTorch would run nine broadcast adds per forward for them)."""
        self.head_bias = tuple([br[i][-1].bias.detach().float().clone() for i in range(self.nl)]
                               for br in (self.cv2, self.cv3, self.cv4))
        for br in (self.cv2, self.cv3, self.cv4):
            for i in range(self.nl):
                br[i][-1].bias = None
        self.proto.strip_bias()

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if getattr(self, "head_bias", None) is not None:
            self.head_bias = tuple([fn(t).float() for t in g] for g in self.head_bias)
        return self

    def shift_cls_bias(self, shift: float):
        if getattr(self, "head_bias", None) is not None:
            for t in self.head_bias[1]:
                t += shift
        else:
            for b in self.cv3:
                b[-1].bias.data += shift

    def _grid(self, shapes, device, dtype):
        key = (tuple(shapes), str(device), dtype)
        if key not in self._anchors:
            pts, strides = [], []
            for (h, w), s in zip(shapes, self.stride):
                sy, sx = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float32) + 0.5,
                                        torch.arange(w, device=device, dtype=torch.float32) + 0.5, indexing="ij")
                pts.append(torch.stack((sx, sy), -1).view(-1, 2))
                strides.append(torch.full((h * w, 1), float(s), device=device))
            self._anchors[key] = (torch.cat(pts).t().contiguous().to(dtype), torch.cat(strides).t().contiguous().to(dtype))
        return self._anchors[key]

    def forward(self, feats):
        protos = self.proto(feats[0])
        B = protos.shape[0]
        box, cls, mc, shapes = [], [], [], []
        for i, x in enumerate(feats):
            shapes.append(x.shape[2:])
            box.append(self.cv2[i](x))
            cls.append(self.cv3[i](x))
            mc.append(self.cv4[i](x))
        cl = torch.channels_last
        if getattr(self, "fused_tails", False) and box[0].is_cuda and box[0].dtype == torch.float16 and all(
                t.is_contiguous(memory_format=cl) or t.shape[1] == 1 for t in box + cls + mc):
            from . import ops
            return ops.yolo_head_decode(box, cls, mc, self.stride, self.nc, self.nm, getattr(self, "head_bias", None)), protos
        hb = getattr(self, "head_bias", None)
        if hb is not None:                                        # biases were stripped from the convolutions
            box, cls, mc = ([t + g[i].to(t.dtype).view(1, -1, 1, 1) for i, t in enumerate(ts)]
                            for ts, g in ((box, hb[0]), (cls, hb[1]), (mc, hb[2])))
        box, cls, mc = (torch.cat([t.flatten(2) for t in ts], 2) for ts in (box, cls, mc))
        A = box.shape[2]
        anchors, strides = self._grid(shapes, box.device, box.dtype)
        dist = (box.view(B, 4, self.reg_max, A).softmax(2) * self.bins.to(box.dtype)).sum(2)       # DFL
        lt, rb = dist.chunk(2, 1)
        x1y1, x2y2 = anchors.unsqueeze(0) - lt, anchors.unsqueeze(0) + rb
        xywh = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
        head = torch.cat((xywh, cls.sigmoid(), mc), 1)
        return head, protos


def _up_cat(a, b):
    """Upsample(x2, nearest) + Concat of the neck (yaml layers 11-12, 14-15)."""
    if a.is_cuda and a.dtype != torch.float32 and a.is_contiguous(memory_format=torch.channels_last) \
            and b.is_contiguous(memory_format=torch.channels_last):
        from . import ops
        return ops.upsample2x_concat(a, b)
    return torch.cat((F.interpolate(a, scale_factor=2.0, mode="nearest"), b), 1)


class YOLO11sSeg(nn.Module):
    """yolo11-seg.yaml, scale s (depth 0.50, width 0.50, max_channels 1024)."""

    def __init__(self, nc=4):
        super().__init__()
        self.nc = nc
        self.l0 = Conv(3, 32, 3, 2)
        self.l1 = Conv(32, 64, 3, 2)
        self.l2 = C3k2(64, 128, 1, False, 0.25)
        self.l3 = Conv(128, 128, 3, 2)
        self.l4 = C3k2(128, 256, 1, False, 0.25)
        self.l5 = Conv(256, 256, 3, 2)
        self.l6 = C3k2(256, 256, 1, True)
        self.l7 = Conv(256, 512, 3, 2)
        self.l8 = C3k2(512, 512, 1, True)
        self.l9 = SPPF(512, 512, 5)
        self.l10 = C2PSA(512, 512, 1)
        self.l13 = C3k2(512 + 256, 256, 1, False)
        self.l16 = C3k2(256 + 256, 128, 1, False)
        self.l17 = Conv(128, 128, 3, 2)
        self.l19 = C3k2(128 + 256, 256, 1, False)
        self.l20 = Conv(256, 256, 3, 2)
        self.l22 = C3k2(256 + 512, 512, 1, True)
        self.head = Segment(nc, (128, 256, 512), 32, 128)
        self.head.bias_init()

    def forward(self, x):
        x = self.l2(self.l1(self.l0(x)))
        p3 = self.l4(self.l3(x))
        p4 = self.l6(self.l5(p3))
        p5 = self.l10(self.l9(self.l8(self.l7(p4))))
        u4 = self.l13(_up_cat(p5, p4))
        n3 = self.l16(_up_cat(u4, p3))
        n4 = self.l19(torch.cat((self.l17(n3), u4), 1))
        n5 = self.l22(torch.cat((self.l20(n4), p5), 1))
        return self.head((n3, n4, n5))

    @torch.no_grad()
    def shift_class_bias(self, sample: torch.Tensor, conf: float = 0.3, frac: float = 0.01) -> float:
        """Random-init weights score ~0 everywhere, so nothing would reach NMS.  Shift the
        class-branch biases by one constant so that ``frac`` of the anchors of ``sample`` exceed ``conf``.
        Returns the shift (reported with every benchmark)."""
        head, _ = self(sample)
        s = head[:, 4:4 + self.nc].amax(1).float().flatten().clamp(1e-6, 1 - 1e-6)
        logit = torch.log(s) - torch.log1p(-s)
        q = torch.quantile(logit[torch.randperm(logit.numel(), device=logit.device)[:1_000_000]], 1.0 - frac)
        shift = float(math.log(conf / (1 - conf)) - q)
        self.head.shift_cls_bias(shift)
        return shift


def finalize_model(m: YOLO11sSeg, device, dtype=torch.float16, fuse: bool = True) -> YOLO11sSeg:
    """Fold BatchNorm, split/strip what the fused kernels take over, move to the device; inference only."""
    m = m.eval()
    if fuse:
        for mod in m.modules():
            if isinstance(mod, Conv):
                mod.fuse()
        for mod in list(m.modules()):
            if isinstance(mod, (C3k2, C2PSA)):
                mod.split_cv1()
            if isinstance(mod, (SPPF, Segment)):
                mod.fused_tails = True                          # K10: one-pass SPPF pooling and head decode
            if isinstance(mod, Segment):
                mod.strip_bias()
    m = m.to(device=device, dtype=dtype).to(memory_format=torch.channels_last)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def build_model(nc: int, device, dtype=torch.float16, seed: int = 0, fuse: bool = True) -> YOLO11sSeg:
    """Seeded random-init network (the reference's weights are not distributed; see ``weights.load_model``)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = YOLO11sSeg(nc).eval()
    torch.random.set_rng_state(g)
    return finalize_model(m, device, dtype, fuse)
