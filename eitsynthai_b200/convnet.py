"""YOLO11s-seg forward pass on libeitb200's own convolution kernels (K11/K12).

``YOLO11sSeg`` (yolo_seg.py) restates the architecture the reference loads through
``ultralytics.YOLO`` (kt_service/ai_tools/ai_tools.py:69-71) and owns the weights; this module packs
those weights once and runs the same graph without cuDNN: every Conv is one ``eitb_conv2d_nhwc``
launch (tcgen05 implicit GEMM, TMA operands, bias + SiLU + residual + concat-slice write fused), the
stem and the depthwise convolutions are K12 launches, SPPF / head decode are K10, and only the two
batched matmuls + softmax of the single attention block stay in PyTorch.  Activations are NHWC fp16
buffers; a producer writes straight into the channel slice of the concat buffer its consumer reads.
"""
from __future__ import annotations

import torch

from . import cabi, ops
from .yolo_seg import (C2PSA, C3k, C3k2, Bottleneck, Conv, YOLO11sSeg)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class Act:
    """Channels [off, off + c) of an NHWC fp16 buffer [B, H, W, Ctot]."""
    __slots__ = ("buf", "off", "c")

    def __init__(self, buf: torch.Tensor, off: int = 0, c: int | None = None):
        self.buf, self.off = buf, off
        self.c = buf.shape[3] - off if c is None else c

    @property
    def bhw(self):
        return self.buf.shape[0], self.buf.shape[1], self.buf.shape[2]

    def slice(self, off: int, c: int) -> "Act":
        return Act(self.buf, self.off + off, c)

    def nchw(self) -> torch.Tensor:
        """The slice as a logical [B, C, H, W] tensor (a channels-last view when it spans the buffer)."""
        return self.buf[..., self.off:self.off + self.c].permute(0, 3, 1, 2)


#: when a dict, ``conv`` adds the algorithmic work of every launch to it (bench.py's roofline accounting):
#: "flops" (2 * MACs), "bytes" (input + output + residual + weights, each counted once), "launches"
STATS = None


def _account(L, B, Ho, Wo, x: Act, res) -> None:
    g = L.cin if L.kind == "dw" else 1
    STATS["flops"] = STATS.get("flops", 0) + 2 * B * Ho * Wo * L.cout * (L.cin // g) * L.k * L.k
    H, W = x.buf.shape[1], x.buf.shape[2]
    STATS["bytes"] = STATS.get("bytes", 0) + 2 * B * (H * W * L.cin + Ho * Wo * L.cout * (2 if res is not None else 1)) + L.w.numel() * L.w.element_size()
    STATS["launches"] = STATS.get("launches", 0) + 1
    k = STATS.setdefault("by_kind", {}).setdefault(L.kind, [0, 0])
    k[0] += 2 * B * Ho * Wo * L.cout * (L.cin // g) * L.k * L.k
    k[1] += 1


def _new(B, H, W, C, dev) -> Act:
    return Act(torch.empty((B, H, W, C), dtype=torch.float16, device=dev))


class PackedConv:
    """Weights of one convolution in the layout its kernel reads."""

    def __init__(self, kind, w, bias, cin, cout, k, s, act):
        self.kind, self.w, self.bias, self.cin, self.cout, self.k, self.s, self.act = kind, w, bias, cin, cout, k, s, act

    @staticmethod
    def _gemm_weight(w: torch.Tensor) -> torch.Tensor:
        cout, cin, k, _ = w.shape
        cp = (cout + 15) // 16 * 16
        out = torch.zeros((k * k, cp, cin), dtype=torch.float16, device=w.device)
        out[:, :cout] = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).to(torch.float16)
        return out.contiguous()

    @classmethod
    def from_weight(cls, w: torch.Tensor, bias, stride: int, groups: int, act: bool) -> "PackedConv":
        cout, cin_g, k, _ = w.shape
        b = None if bias is None else bias.detach().float().contiguous()
        if groups == 1 and cin_g == 3:
            assert k == 3 and stride == 2, "stem shape"
            return cls("stem", w.detach().float().permute(2, 3, 1, 0).reshape(27, cout).contiguous(), b, 3, cout, 3, 2, act)
        if groups == 1:
            return cls("gemm", cls._gemm_weight(w.detach()), b, cin_g, cout, k, stride, act)
        assert groups == cout and cin_g == 1 and k == 3 and stride == 1, "depthwise 3x3 only"
        return cls("dw", w.detach().reshape(cout, 9).t().contiguous().to(torch.float16), b, cout, cout, 3, 1, act)

    @classmethod
    def from_module(cls, m) -> "PackedConv":
        if isinstance(m, Conv):
            assert m.fused_bias is not None, "fuse() the model first"
            c = m.conv
            return cls.from_weight(c.weight, m.fused_bias, c.stride[0], c.groups, m.has_act)
        return cls.from_weight(m.weight, m.bias, m.stride[0], m.groups, False)          # plain nn.Conv2d

    @classmethod
    def merged(cls, mods) -> "PackedConv":
        """Several 1x1 Convs reading the same input as one convolution (output channels side by side)."""
        w = torch.cat([m.conv.weight for m in mods], 0)
        b = torch.cat([m.fused_bias for m in mods], 0)
        return cls.from_weight(w, b, 1, 1, mods[0].has_act)


def conv(x: Act, L: PackedConv, out: Act | None = None, res: Act | None = None, up=(1, 0, 0), gray: bool = False,
         pre: Act | None = None) -> Act:
    """y = act(conv(x) + bias) [+ res], written to ``out`` (a slice) or to a fresh buffer.  ``gray``: the stem's
    three input channels are equal (a replicated gray image).  ``pre``: a half-resolution map added, 2x nearest
    upsampled, BEFORE the activation (1x1 convolutions only): y = act(conv(x) + bias + up2(pre))."""
    if pre is not None:
        assert res is None and L.kind == "gemm" and L.k == 1 and L.s == 1
        res = pre
    B, H, W = x.bhw
    dev = x.buf.device
    assert x.c == L.cin, (x.c, L.cin)
    pad = L.k // 2
    Ho, Wo = (H + 2 * pad - L.k) // L.s + 1, (W + 2 * pad - L.k) // L.s + 1
    if out is None:
        out = _new(B, Ho * up[0], Wo * up[0], max(8, (L.cout + 7) // 8 * 8), dev)
        out.c = L.cout
    if STATS is not None:
        _account(L, B, Ho, Wo, x, res)
    with torch.cuda.device(dev):
        if L.kind == "gemm":
            cabi.call("eitb_conv2d_nhwc", x.buf.data_ptr(), B, H, W, x.buf.shape[3], x.off, L.cin, L.w.data_ptr(),
                      0 if L.bias is None else L.bias.data_ptr(), L.cout, L.k, L.s, int(L.act),
                      0 if res is None else res.buf.data_ptr(), 0 if res is None else res.buf.shape[3],
                      0 if res is None else res.off, 2 if pre is not None else 1, out.buf.data_ptr(), out.buf.shape[3], out.off,
                      up[0], up[1], up[2], _stream(dev))
        elif L.kind == "dw":
            assert res is None and up[0] == 1
            cabi.call("eitb_dwconv3x3_nhwc", x.buf.data_ptr(), B, H, W, x.buf.shape[3], x.off, L.cin, L.w.data_ptr(),
                      0 if L.bias is None else L.bias.data_ptr(), int(L.act), out.buf.data_ptr(), out.buf.shape[3], out.off,
                      _stream(dev))
        else:
            assert res is None and up[0] == 1 and x.off == 0 and x.buf.shape[3] == 3
            cabi.call("eitb_stem_conv3x3s2_nhwc", x.buf.data_ptr(), B, H, W, L.w.data_ptr(),
                      0 if L.bias is None else L.bias.data_ptr(), L.cout, int(L.act), int(gray), out.buf.data_ptr(),
                      out.buf.shape[3], out.off, _stream(dev))
    return out


def stem_u8(img: torch.Tensor, L: PackedConv) -> Act:
    """The stem on the u8 window image [B, H, W] (K1's u8 output): preprocess (u8 -> fp16 / 255, three equal channels)
    fused into the kernel's tile loader."""
    assert L.kind == "stem" and img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    B, H, W = img.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = _new(B, Ho, Wo, L.cout, img.device)
    if STATS is not None:
        STATS["flops"] = STATS.get("flops", 0) + 2 * B * Ho * Wo * L.cout * 27
        STATS["bytes"] = STATS.get("bytes", 0) + B * H * W + 2 * B * Ho * Wo * L.cout + L.w.numel() * 4
        STATS["launches"] = STATS.get("launches", 0) + 1
        k = STATS.setdefault("by_kind", {}).setdefault("stem", [0, 0])
        k[0] += 2 * B * Ho * Wo * L.cout * 27
        k[1] += 1
    with torch.cuda.device(img.device):
        cabi.call("eitb_stem_conv3x3s2_nhwc", img.data_ptr(), B, H, W, L.w.data_ptr(), 0 if L.bias is None else L.bias.data_ptr(),
                  L.cout, int(L.act), 2, out.buf.data_ptr(), out.buf.shape[3], 0, _stream(img.device))
    return out


class ConvNet:
    """Executes a fused ``YOLO11sSeg`` with libeitb200 kernels.  ``net(x)`` takes the channels-last
    [B, 3, H, W] fp16 input K1 produces and returns (head [B, 4+nc+nm, A] fp16, protos [B, 32, H/4, W/4]
    channels-last) like the module it wraps."""

    def __init__(self, model: YOLO11sSeg):
        self.nc = model.nc
        self.model = model
        P = PackedConv.from_module
        self.p = {}
        for name in ("l0", "l1", "l3", "l5", "l7", "l17", "l20"):
            self.p[name] = P(getattr(model, name))
        for name in ("l2", "l4", "l6", "l8", "l19", "l22"):
            self.p[name] = self._pack_c3k2(getattr(model, name))
        # the two blocks behind Upsample + Concat (yaml 11-13, 14-16): cv1 split at the upsampled channels
        self.p["l13"] = self._pack_c3k2(model.l13, split=model.l10.cv2.conv.out_channels)
        self.p["l16"] = self._pack_c3k2(model.l16, split=model.l13.cv2.conv.out_channels)
        self.p["l9"] = (P(model.l9.cv1), P(model.l9.cv2))
        self.p["l10"] = self._pack_c2psa(model.l10)
        h = model.head
        self.head_bias = getattr(h, "head_bias", None)
        assert self.head_bias is not None, "build_model(fuse=True) strips the branch biases into K10"
        self.p["box"] = [[P(m) for m in br] for br in h.cv2]
        self.p["cls"] = [[P(br[0][0]), P(br[0][1]), P(br[1][0]), P(br[1][1]), P(br[2])] for br in h.cv3]
        self.p["mc"] = [[P(m) for m in br] for br in h.cv4]
        pr = h.proto
        up_w = pr.upsample.weight.detach()                         # [Cin, Cout, 2, 2]
        up_b = getattr(pr, "up_bias", None)
        if up_b is None and pr.upsample.bias is not None:
            up_b = pr.upsample.bias.detach().float()
        # ConvTranspose2d(k 2, s 2) = four 1x1 convolutions, one per output phase (dy, dx); the two dx phases of a row are
        # neighbouring pixels of the NHWC output, i.e. one 2C-channel "pixel" at stride 2: one launch per dy with the
        # two taps' output channels side by side reads the input twice instead of four times
        self.p["proto"] = (P(pr.cv1),
                           [PackedConv.from_weight(torch.cat([up_w[:, :, dy, 0].t(), up_w[:, :, dy, 1].t()], 0).contiguous()[:, :, None, None],
                                                   None if up_b is None else torch.cat([up_b, up_b]), 1, 1, False)
                            for dy in range(2)],
                           P(pr.cv2), P(pr.cv3))
        self.stride, self.nm = h.stride, h.nm

    # ------------------------------------------------------------------ packing of the composite blocks
    @staticmethod
    def _halves(m):
        return [m.cv1a, m.cv1b] if hasattr(m, "cv1a") else None

    def _pack_bottleneck(self, b: Bottleneck):
        return ("b", PackedConv.from_module(b.cv1), PackedConv.from_module(b.cv2), b.add)

    def _pack_c3k(self, m: C3k):
        return ("c3k", PackedConv.merged([m.cv1, m.cv2]), PackedConv.from_module(m.cv3), [self._pack_bottleneck(b) for b in m.m],
                m.cv1.conv.out_channels)

    def _pack_c3k2(self, m: C3k2, split: int | None = None):
        """``split``: the block reads Concat(Upsample(a), b) with ``split`` channels in ``a``: cv1 becomes the pair
        (Wa without bias / activation, Wb with both) for  act(up2(Wa a) + Wb b + bias)."""
        hv = self._halves(m)
        mods = hv if hv else [m.cv1]
        if split is None:
            cv1 = PackedConv.merged(mods)
        else:
            w = torch.cat([x.conv.weight for x in mods], 0)
            b = torch.cat([x.fused_bias for x in mods], 0)
            cv1 = (PackedConv.from_weight(w[:, :split].contiguous(), None, 1, 1, False),
                   PackedConv.from_weight(w[:, split:].contiguous(), b, 1, 1, mods[0].has_act))
        inner = [self._pack_c3k(b) if isinstance(b, C3k) else self._pack_bottleneck(b) for b in m.m]
        return (cv1, PackedConv.from_module(m.cv2), inner, m.c)

    def _pack_c2psa(self, m: C2PSA):
        hv = self._halves(m)
        cv1 = PackedConv.merged(hv) if hv else PackedConv.from_module(m.cv1)
        P = PackedConv.from_module
        blocks = []
        for blk in m.m:
            a = blk.attn
            # qkv output channels regrouped from [head][q | k | v] to [q of all heads | k of all heads | v of all heads]:
            # V (and the attention output) then is a contiguous channel slice [head][d] of the buffer -- the layout the
            # position-encoding depthwise conv and proj expect -- and needs no gather copy
            per = 2 * a.key_dim + a.head_dim
            idx = [h * per + off + i for off, n in ((0, a.key_dim), (a.key_dim, a.key_dim), (2 * a.key_dim, a.head_dim))
                   for h in range(a.num_heads) for i in range(n)]
            idx = torch.tensor(idx, device=a.qkv.conv.weight.device)
            qkv = PackedConv.from_weight(a.qkv.conv.weight.detach()[idx], a.qkv.fused_bias[idx], 1, 1, a.qkv.has_act)
            blocks.append((qkv, P(a.pe), P(a.proj), P(blk.ffn[0]), P(blk.ffn[1]), a.num_heads, a.key_dim, a.head_dim, a.scale))
        return (cv1, P(m.cv2), blocks, m.c)

    # ------------------------------------------------------------------ block executors
    def _bottleneck(self, x: Act, pk, out: Act | None) -> Act:
        _, cv1, cv2, add = pk
        return conv(conv(x, cv1), cv2, out=out, res=x if add else None)

    def _c3k(self, x: Act, pk, out: Act | None) -> Act:
        _, cv12, cv3, bns, c_ = pk
        B, H, W = x.bhw
        buf = _new(B, H, W, 2 * c_, x.buf.device)
        conv(x, cv12, out=buf)                                     # [cv1(x) | cv2(x)]
        h = buf.slice(0, c_)
        for i, b in enumerate(bns):
            last = i == len(bns) - 1
            h = self._bottleneck(h, b, buf.slice(0, c_) if last else None)   # the last one overwrites cv1(x): no longer needed
        return conv(buf, cv3, out=out)

    def _c3k2(self, x: Act, pk, out: Act | None = None, low: Act | None = None) -> Act:
        """``low``: the block's input is Concat(Upsample(low), x) (cv1 packed with ``split``): Wa runs at the low
        resolution and its result enters the full-resolution Wb convolution before the activation."""
        cv1, cv2, inner, c = pk
        B, H, W = x.bhw
        n = len(inner)
        buf = _new(B, H, W, (2 + n) * c, x.buf.device)
        if low is None:
            conv(x, cv1, out=buf.slice(0, 2 * c))
        else:
            conv(x, cv1[1], out=buf.slice(0, 2 * c), pre=conv(low, cv1[0]))
        h = buf.slice(c, c)
        for i, b in enumerate(inner):
            dst = buf.slice((2 + i) * c, c)
            h = self._c3k(h, b, dst) if b[0] == "c3k" else self._bottleneck(h, b, dst)
        return conv(buf, cv2, out=out)

    def _c2psa(self, x: Act, pk, out: Act | None = None) -> Act:
        cv1, cv2, blocks, c = pk
        B, H, W = x.bhw
        dev = x.buf.device
        buf = _new(B, H, W, 2 * c, dev)
        conv(x, cv1, out=buf)
        b = buf.slice(c, c)
        for qkv_l, pe_l, proj_l, f0, f1, heads, kd, hd, scale in blocks:
            qkv = conv(b, qkv_l)                                    # [B, H, W, q | k | v], each [head][d]
            N = H * W
            flat = qkv.buf.view(B, N, -1)
            q = flat[..., :heads * kd].view(B, N, heads, kd).transpose(1, 2)                  # [B, heads, N, kd] (strided views)
            k = flat[..., heads * kd:2 * heads * kd].view(B, N, heads, kd).transpose(1, 2)
            v = flat[..., 2 * heads * kd:].view(B, N, heads, hd).transpose(1, 2)              # [B, heads, N, hd]
            att = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale)      # the one PyTorch-owned op of the network
            pe = conv(qkv.slice(2 * heads * kd, heads * hd), pe_l)  # depthwise position encoding straight from the V slice
            pe.buf.view(B, N, heads, hd).add_(att.transpose(1, 2))  # o = attention + pe(v), in the NHWC buffer
            x1 = conv(pe, proj_l, res=b)
            conv(conv(x1, f0), f1, out=b, res=x1)
        return conv(buf, cv2, out=out)

    # ------------------------------------------------------------------ whole network
    @torch.no_grad()
    def __call__(self, x: torch.Tensor, gray: bool = False, out=None):
        """``gray``: the three channels of ``x`` are equal (K1 / letterbox output) -- lets the stem read one.
        ``out`` = (head [B, 4+nc+nm, A] fp16, protos [B, H/4, W/4, nm] fp16 NHWC): buffers the last two kernels write
        into instead of fresh allocations (hand-over buffers of CUDA graphs that replay on different streams)."""
        p = self.p
        dev = x.device
        if x.dtype == torch.uint8:                                  # the u8 window image [B, H, W]: preprocess fused into the stem
            a = stem_u8(x, p["l0"])
        else:
            assert x.is_cuda and x.dtype == torch.float16 and x.shape[1] == 3 and x.is_contiguous(memory_format=torch.channels_last)
            a = conv(Act(x.permute(0, 2, 3, 1)), p["l0"], gray=gray)    # NHWC view of the channels-last input
        a = conv(a, p["l1"])
        a = self._c3k2(a, p["l2"])
        p3 = self._c3k2(conv(a, p["l3"]), p["l4"])
        p4 = self._c3k2(conv(p3, p["l5"]), p["l6"])
        a = self._c3k2(conv(p4, p["l7"]), p["l8"])
        y0 = conv(a, p["l9"][0])
        a = conv(Act(ops.sppf_pool_concat(y0.nchw()).permute(0, 2, 3, 1)), p["l9"][1])
        B = x.shape[0]
        cat22 = _new(B, a.buf.shape[1], a.buf.shape[2], 256 + 512, dev)              # [l20(n4) | p5]: p5 is born in place
        p5 = self._c2psa(a, p["l10"], out=cat22.slice(256, 512))
        # neck: Upsample + Concat never exist -- a 1x1 convolution commutes with nearest upsampling, so cv1 of l13 / l16
        # runs its upsampled half at the low resolution and adds it inside the epilogue of the other half (K11 res_mode 2);
        # the two down-path concats are written in place by their producers
        cat19 = _new(B, p4.buf.shape[1], p4.buf.shape[2], 128 + 256, dev)            # [l17(n3) | u4]
        u4 = self._c3k2(p4, p["l13"], out=cat19.slice(128, 256), low=p5)
        n3 = self._c3k2(p3, p["l16"], low=u4)
        conv(n3, p["l17"], out=cat19.slice(0, 128))
        n4 = self._c3k2(cat19, p["l19"])
        conv(n4, p["l20"], out=cat22.slice(0, 256))
        n5 = self._c3k2(cat22, p["l22"])
        return self._head((n3, n4, n5), out)

    def _head(self, feats, out=None):
        p = self.p
        cv1, ups, cv2, cv3 = p["proto"]
        t = conv(feats[0], cv1)
        B, H, W = t.bhw
        up = _new(B, 2 * H, 2 * W, ups[0].cout // 2, t.buf.device)
        for dy in range(2):
            conv(t, ups[dy], out=up, up=(2, dy, 0))
        protos = conv(conv(up, cv2), cv3, out=Act(out[1]) if out is not None else None)
        box, cls, mc = [], [], []
        for i, f in enumerate(feats):
            b = p["box"][i]
            box.append(conv(conv(conv(f, b[0]), b[1]), b[2]).nchw())
            c = p["cls"][i]
            cl = conv(conv(conv(conv(conv(f, c[0]), c[1]), c[2]), c[3]), c[4])
            cls.append(cl.buf.permute(0, 3, 1, 2))                  # padded to 8 channels per pixel
            m = p["mc"][i]
            mc.append(conv(conv(conv(f, m[0]), m[1]), m[2]).nchw())
        head = ops.yolo_head_decode(box, cls, mc, self.stride, self.nc, self.nm, self.head_bias, cls_cstride=cls[0].shape[1],
                                    out=out[0] if out is not None else None)
        return head, protos.nchw()
