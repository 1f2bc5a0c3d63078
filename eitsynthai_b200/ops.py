"""Device-tensor wrappers around the C-ABI (``cabi``): PyTorch supplies memory and streams,
libeitb200 does the work.  Every function enqueues on the current CUDA stream of the
tensor's device and returns device tensors; nothing here computes on the CPU.
"""
from __future__ import annotations

import torch

from . import cabi

_DT = {torch.float32: cabi.F32, torch.float16: cabi.F16, torch.bfloat16: cabi.BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libeitb200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------ K1
def hu_window(px: torch.Tensor, lo: int = -160, hi: int = 240, rot180: bool = True,
              body_mask: torch.Tensor | None = None, want_u8: bool = True,
              nchw_dtype: torch.dtype | None = torch.float16, channels_last: bool = False,
              u8_out: torch.Tensor | None = None, nchw_out: torch.Tensor | None = None):
    """[B,H,W] int16 -> (u8 [B,H,W] or None, NCHW [B,3,H,W] or None).  classic_norm + mask + /255.
    ``channels_last`` returns the same logical tensor in torch.channels_last memory format; ``u8_out`` / ``nchw_out``:
    buffers to write into instead of fresh allocations."""
    _chk(px, torch.int16, "px")
    B, H, W = px.shape
    if body_mask is not None:
        _chk(body_mask, torch.uint8, "body_mask")
        assert body_mask.shape == px.shape
    u8 = None
    if want_u8:
        u8 = torch.empty((B, H, W), dtype=torch.uint8, device=px.device) if u8_out is None else u8_out
        _chk(u8, torch.uint8, "u8_out")
        assert u8.shape == px.shape
    nchw = None
    if nchw_dtype is not None:
        if nchw_out is None:
            nchw = torch.empty((B, 3, H, W), dtype=nchw_dtype, device=px.device,
                               memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        else:
            nchw = nchw_out
            assert nchw.dtype == nchw_dtype and tuple(nchw.shape) == (B, 3, H, W) and nchw.is_cuda
            assert nchw.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    with torch.cuda.device(px.device):
        cabi.call("eitb_hu_window_nchw", px.data_ptr(), B, H, W, lo, hi, int(rot180), _ptr(body_mask),
                  _ptr(u8), _ptr(nchw), _DT.get(nchw_dtype, cabi.F32), int(channels_last), _stream(px))
    return u8, nchw


def u8_to_nchw(gray: torch.Tensor, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    _chk(gray, torch.uint8, "gray")
    B, H, W = gray.shape
    out = torch.empty((B, 3, H, W), dtype=dtype, device=gray.device)
    with torch.cuda.device(gray.device):
        cabi.call("eitb_u8_to_nchw", gray.data_ptr(), B, H, W, out.data_ptr(), _DT[dtype], _stream(gray))
    return out


# ------------------------------------------------------------------------------------ K2
def body_mask(px: torch.Tensor, slope: int = 1, intercept: int = -1024, flipud: bool = True,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """[B,H,W] int16 -> [B,H,W] u8 {0,255}: get_axial_slice_body_mask(_nii).  ``out``: write into this buffer."""
    _chk(px, torch.int16, "px")
    B, H, W = px.shape
    if out is None:
        out = torch.empty((B, H, W), dtype=torch.uint8, device=px.device)
    else:
        _chk(out, torch.uint8, "out")
        assert out.shape == px.shape
    lib = cabi.load()
    nbytes = lib.eitb_body_mask_workspace_bytes(B, H, W)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=px.device)
    with torch.cuda.device(px.device):
        cabi.call("eitb_body_mask", px.data_ptr(), B, H, W, slope, intercept, int(flipud), out.data_ptr(),
                  ws.data_ptr(), nbytes, _stream(px))
    return out


def cc_label(mask: torch.Tensor, connectivity: int = 4, link_outside: bool = False) -> torch.Tensor:
    """[B,H,W] u8 -> int32 labels (smallest pixel index of the component, -2 unset, -1 frame-connected)."""
    _chk(mask, torch.uint8, "mask")
    B, H, W = mask.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        cabi.call("eitb_cc_label", mask.data_ptr(), B, H, W, connectivity, int(link_outside), out.data_ptr(), _stream(mask))
    return out


# ------------------------------------------------------------------------------------ K3
def front_rows(px: torch.Tensor, order: torch.Tensor | None, n: int, row: int, flip_x: bool, flip_z: bool,
               minmax: torch.Tensor | None = None):
    """Coronal mid-rows [n,W] int16 + running (min,max) int32[2]."""
    _chk(px, torch.int16, "px")
    _, H, W = px.shape
    if order is not None:
        _chk(order, torch.int32, "order")
    rows = torch.empty((n, W), dtype=torch.int16, device=px.device)
    if minmax is None:
        minmax = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=px.device)
    with torch.cuda.device(px.device):
        cabi.call("eitb_front_rows", px.data_ptr(), _ptr(order), n, H, W, row, int(flip_x), int(flip_z),
                  rows.data_ptr(), minmax.data_ptr(), _stream(px))
    return rows, minmax


def front_rows_batch(px: torch.Tensor, order: torch.Tensor | None, geom: torch.Tensor, minmax: torch.Tensor | None = None):
    """px [S,n,H,W] int16 (may be a view with its own series stride), order [S,n] int32, geom [S,3] int32
    (row, flip_x, flip_z per series) -> rows [S,n,W] int16, minmax [S,2] int32: one launch for the batch."""
    if not px.is_cuda or px.dtype != torch.int16 or px.dim() != 4 or not px[0].is_contiguous():
        raise ValueError("px must be a CUDA int16 [S,n,H,W] tensor with contiguous series")
    S, n, H, W = px.shape
    _chk(geom, torch.int32, "geom")
    if order is not None:
        _chk(order, torch.int32, "order")
    rows = torch.empty((S, n, W), dtype=torch.int16, device=px.device)
    if minmax is None:
        minmax = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=px.device).repeat(S, 1)
    with torch.cuda.device(px.device):
        cabi.call("eitb_front_rows_batch", px.data_ptr(), px.stride(0) if S > 1 else n * H * W, _ptr(order), geom.data_ptr(),
                  S, n, H, W, rows.data_ptr(), minmax.data_ptr(), _stream(px))
    return rows, minmax


def rows_h2d(px_host: torch.Tensor, row: int, out: torch.Tensor) -> torch.Tensor:
    """px_host [n,H,W] int16 pinned host tensor -> out [n,W] (device): row ``row`` of every slice, one strided DMA."""
    if px_host.is_cuda or not px_host.is_pinned() or px_host.dtype != torch.int16 or not px_host.is_contiguous():
        raise ValueError("px_host must be a contiguous pinned int16 host tensor")
    n, H, W = px_host.shape
    _chk(out, torch.int16, "out")
    assert out.numel() == n * W
    with torch.cuda.device(out.device):
        cabi.call("eitb_rows_h2d", px_host.data_ptr(), n, H, W, row, out.data_ptr(), _stream(out))
    return out


def minmax_u8(rows: torch.Tensor, minmax: torch.Tensor) -> torch.Tensor:
    _chk(rows, torch.int16, "rows")
    _chk(minmax, torch.int32, "minmax")
    out = torch.empty(rows.shape, dtype=torch.uint8, device=rows.device)
    with torch.cuda.device(rows.device):
        cabi.call("eitb_minmax_u8", rows.data_ptr(), rows.numel(), minmax.data_ptr(), out.data_ptr(), _stream(rows))
    return out


def letterbox_nchw(gray: torch.Tensor, nh: int, nw: int, top: int, left: int, outH: int, outW: int,
                   dtype: torch.dtype = torch.float16) -> torch.Tensor:
    _chk(gray, torch.uint8, "gray")
    B, H, W = gray.shape
    out = torch.empty((B, 3, outH, outW), dtype=dtype, device=gray.device)
    with torch.cuda.device(gray.device):
        cabi.call("eitb_letterbox_nchw", gray.data_ptr(), B, H, W, nh, nw, top, left, outH, outW,
                  out.data_ptr(), _DT[dtype], _stream(gray))
    return out


# ------------------------------------------------------------------------------------ K4
def rib_select(xyxy: torch.Tensor, k: torch.Tensor, image_width: float = 512.0,
               custom: torch.Tensor | None = None) -> torch.Tensor:
    """[S,max_k,4] f32 boxes, [S] counts -> [S,4] int32 (y6, y7, mid+custom, ok)."""
    _chk(xyxy, torch.float32, "xyxy")
    _chk(k, torch.int32, "k")
    S, max_k, _ = xyxy.shape
    if custom is not None:
        _chk(custom, torch.int32, "custom")
    out = torch.empty((S, 4), dtype=torch.int32, device=xyxy.device)
    with torch.cuda.device(xyxy.device):
        cabi.call("eitb_rib_select", xyxy.data_ptr(), k.data_ptr(), S, max_k, float(image_width), _ptr(custom),
                  out.data_ptr(), _stream(xyxy))
    return out


# ------------------------------------------------------------------------------------ K5
def nms(head: torch.Tensor, nc: int, conf: float = 0.3, iou: float = 0.7, max_det: int = 300,
        max_wh: float = 7680.0, want_idx: bool = True):
    """[B,4+nc+nm,A] -> dets [B,max_det,6+nm] f32, keep_idx [B,max_det] i32, n [B] i32."""
    if head.dtype not in _DT:
        raise TypeError(f"head dtype {head.dtype} unsupported")
    _chk(head, None, "head")
    B, Cc, A = head.shape
    nm = Cc - 4 - nc
    dets = torch.zeros((B, max_det, 6 + nm), dtype=torch.float32, device=head.device)
    idx = torch.full((B, max_det), -1, dtype=torch.int32, device=head.device) if want_idx else None
    n = torch.zeros((B,), dtype=torch.int32, device=head.device)
    with torch.cuda.device(head.device):
        cabi.call("eitb_nms", head.data_ptr(), _DT[head.dtype], B, nc, nm, A, conf, iou, max_det, max_wh,
                  dets.data_ptr(), _ptr(idx), n.data_ptr(), 0, 0, _stream(head))
    return dets, idx, n


def scale_boxes(dets: torch.Tensor, n: torch.Tensor, gain: float, pad_x: float, pad_y: float, orig_w: float,
                orig_h: float) -> torch.Tensor:
    """dets [B,max_det,D] (network px) -> xyxy [B,max_det,4] in original-image px (ultralytics scale_boxes)."""
    _chk(dets, torch.float32, "dets")
    _chk(n, torch.int32, "n")
    B, max_det, D = dets.shape
    out = torch.empty((B, max_det, 4), dtype=torch.float32, device=dets.device)
    with torch.cuda.device(dets.device):
        cabi.call("eitb_scale_boxes", dets.data_ptr(), n.data_ptr(), B, max_det, D, float(gain), float(pad_x),
                  float(pad_y), float(orig_w), float(orig_h), out.data_ptr(), _stream(dets))
    return out


# ------------------------------------------------------------------------------------ K6
def mask_decode(dets: torch.Tensor, n_det: torch.Tensor, protos: torch.Tensor, variant: int = 0,
                want_area: bool = False, want_bits: bool = False, code_out: torch.Tensor | None = None):
    """dets/n from ``nms`` + protos [B,nm,mh,mw] -> overlay code image [B,4mh,4mw] u8
    (+ per-instance areas [B,max_det] i32, + per-instance bit masks [B,max_det,H,W/8] u8).
    ``code_out``: write the code image into this buffer."""
    _chk(dets, torch.float32, "dets")
    _chk(n_det, torch.int32, "n_det")
    if not protos.is_cuda:
        raise ValueError("protos must be a CUDA tensor")
    B, nm, mh, mw = protos.shape
    if protos.is_contiguous():
        nhwc = 0
    elif protos.is_contiguous(memory_format=torch.channels_last):
        nhwc = 1
    else:
        raise ValueError("protos must be contiguous (NCHW) or channels-last")
    max_det = dets.shape[1]
    assert dets.shape[2] == 6 + nm and dets.shape[0] == B
    H, W = 4 * mh, 4 * mw
    if code_out is None:
        code = torch.empty((B, H, W), dtype=torch.uint8, device=dets.device)
    else:
        _chk(code_out, torch.uint8, "code_out")
        assert tuple(code_out.shape) == (B, H, W)
        code = code_out
    area = torch.empty((B, max_det), dtype=torch.int32, device=dets.device) if want_area else None
    bits = torch.empty((B, max_det, H, W // 8), dtype=torch.uint8, device=dets.device) if want_bits else None
    with torch.cuda.device(dets.device):
        cabi.call("eitb_mask_decode", dets.data_ptr(), n_det.data_ptr(), max_det, protos.data_ptr(),
                  _DT[protos.dtype], nhwc, B, nm, mh, mw, H, W, variant, code.data_ptr(), _ptr(area), _ptr(bits),
                  0, 0, _stream(dets))
    return code, area, bits


# ------------------------------------------------------------------------------------ K7
def label_cleanup(code: torch.Tensor, body: torch.Tensor | None) -> torch.Tensor:
    """In-place clear_color_output (when ``body`` is given) + highlight_small_masks on code images."""
    _chk(code, torch.uint8, "code")
    B, H, W = code.shape
    if body is not None:
        _chk(body, torch.uint8, "body")
    lib = cabi.load()
    nbytes = lib.eitb_label_cleanup_workspace_bytes(B, H, W)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=code.device)
    with torch.cuda.device(code.device):
        cabi.call("eitb_label_cleanup", code.data_ptr(), _ptr(body), B, H, W, ws.data_ptr(), nbytes, _stream(code))
    return code


def apply_mask_u8(img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """img [..., H, W] or [H, W, C] u8, mask [H, W] u8 -> img where mask != 0 else 0 (cv2.bitwise_and(x, x, mask=m))."""
    _chk(img, torch.uint8, "img")
    _chk(mask, torch.uint8, "mask")
    n_px = mask.numel()
    assert img.numel() % n_px == 0 and img.numel() // n_px <= 4
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        cabi.call("eitb_apply_mask_u8", img.data_ptr(), mask.data_ptr(), n_px, img.numel() // n_px, out.data_ptr(), _stream(img))
    return out


def class_images(masks: torch.Tensor, cls: torch.Tensor) -> torch.Tensor:
    """masks [n, S, S] fp32, cls [n] int32 -> [4, S, S, 3] u8: create_segmentations_masks' four BGR class images."""
    _chk(masks, torch.float32, "masks")
    _chk(cls, torch.int32, "cls")
    n, H, W = masks.shape
    out = torch.empty((4, H, W, 3), dtype=torch.uint8, device=masks.device)
    with torch.cuda.device(masks.device):
        cabi.call("eitb_class_images", masks.data_ptr(), cls.data_ptr(), n, H * W, out.data_ptr(), _stream(masks))
    return out


def bgr_or_code(bgr: torch.Tensor, value: int, code: torch.Tensor) -> torch.Tensor:
    """code |= value wherever the [H, W, 3] u8 image is non-zero (overlay_segmentation_masks on colour codes)."""
    _chk(bgr, torch.uint8, "bgr")
    _chk(code, torch.uint8, "code")
    assert bgr.numel() == code.numel() * 3
    with torch.cuda.device(code.device):
        cabi.call("eitb_bgr_or_code", bgr.data_ptr(), code.numel(), int(value), code.data_ptr(), _stream(code))
    return code


def codes_to_bgr(code: torch.Tensor) -> torch.Tensor:
    _chk(code, torch.uint8, "code")
    out = torch.empty(code.shape + (3,), dtype=torch.uint8, device=code.device)
    with torch.cuda.device(code.device):
        cabi.call("eitb_codes_to_bgr", code.data_ptr(), out.data_ptr(), code.numel(), _stream(code))
    return out


# ------------------------------------------------------------------------------------ K13
class LabelPolygons:
    """Device-resident result of ``label_polygons``: per image the reference's polygon list."""

    def __init__(self, n, cls, off, xy, status, max_polys, max_points):
        self.n, self.cls, self.off, self.xy, self.status = n, cls, off, xy, status
        self.max_polys, self.max_points = max_polys, max_points

    def to_host(self):
        """[(status, [(class id, (n, 2) int32 array of x, y), ...])] per image (one device->host copy each)."""
        n, cls, off, xy, st = (t.cpu().numpy() for t in (self.n, self.cls, self.off, self.xy, self.status))
        out = []
        for b in range(len(n)):
            polys = [(int(cls[b, i]), xy[b, off[b, i]:off[b, i + 1]].copy()) for i in range(int(n[b]))]
            out.append((int(st[b]), polys))
        return out


def label_polygons(code: torch.Tensor, body: torch.Tensor | None, max_polys: int = 1024, max_points: int = 0) -> LabelPolygons:
    """create_list_crd_from_color_output + get_only_body_mask_contours (utils.py:1191-1279, 1157-1188) on code images."""
    _chk(code, torch.uint8, "code")
    B, H, W = code.shape
    if body is not None:
        _chk(body, torch.uint8, "body")
    max_points = max_points or H * W // 4
    dev = code.device
    n = torch.zeros(B, dtype=torch.int32, device=dev)
    cls = torch.zeros((B, max_polys), dtype=torch.int32, device=dev)
    off = torch.zeros((B, max_polys + 1), dtype=torch.int32, device=dev)
    xy = torch.empty((B, max_points, 2), dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    lib = cabi.load()
    nbytes = lib.eitb_label_polygons_workspace_bytes(B, H, W, max_polys)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        cabi.call("eitb_label_polygons", code.data_ptr(), _ptr(body), B, H, W, max_polys, max_points, n.data_ptr(), cls.data_ptr(),
                  off.data_ptr(), xy.data_ptr(), status.data_ptr(), ws.data_ptr(), nbytes, _stream(code))
    return LabelPolygons(n, cls, off, xy, status, max_polys, max_points)


def polygons_for_mesh(lp: LabelPolygons):
    """The polygon list as K8 reads it (outer contour removed, short polygons dropped, rings closed, ascending area):
    ``(poly_xy [B, max_points + max_polys, 2] f64, poly_off [B, max_polys + 1], poly_cls [B, max_polys], P [B])``."""
    B = lp.n.shape[0]
    dev = lp.n.device
    out_xy = torch.empty((B, lp.max_points + lp.max_polys, 2), dtype=torch.float64, device=dev)
    out_off = torch.zeros((B, lp.max_polys + 1), dtype=torch.int32, device=dev)
    out_cls = torch.zeros((B, lp.max_polys), dtype=torch.int32, device=dev)
    out_n = torch.zeros(B, dtype=torch.int32, device=dev)
    lib = cabi.load()
    nbytes = lib.eitb_polygons_for_mesh_workspace_bytes(B, lp.max_polys)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        cabi.call("eitb_polygons_for_mesh", lp.n.data_ptr(), lp.cls.data_ptr(), lp.off.data_ptr(), lp.xy.data_ptr(), B, lp.max_polys,
                  lp.max_points, out_xy.data_ptr(), out_off.data_ptr(), out_cls.data_ptr(), out_n.data_ptr(), ws.data_ptr(), nbytes,
                  _stream(lp.n))
    return out_xy, out_off, out_cls, out_n


# ------------------------------------------------------------------------------------ K9
def bias_act_(x: torch.Tensor, bias: torch.Tensor | None, silu: bool = True) -> torch.Tensor:
    """In-place bias + SiLU on a channels-last [B,C,H,W] half tensor (conv epilogue of the CNN)."""
    if not x.is_cuda or x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("x must be a channels-last CUDA tensor")
    B, Cc, H, W = x.shape
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    with torch.cuda.device(x.device):
        cabi.call("eitb_bias_act_nhwc", x.data_ptr(), _DT[x.dtype], B * H * W, Cc, _ptr(bias), int(silu), _stream(x))
    return x


def conv_epilogue(src: torch.Tensor, bias: torch.Tensor | None, silu: bool, residual: torch.Tensor | None = None,
                  inplace: bool = True, out2: torch.Tensor | None = None, out2_off: int = 0):
    """y = act(src + bias) [+ residual] (shortcut added after the activation) for a channels-last half tensor; y replaces ``src`` (``inplace``)
    and/or lands in channels [out2_off, out2_off + C) of the channels-last tensor ``out2``."""
    cl = torch.channels_last
    if not src.is_cuda or src.dim() != 4 or not src.is_contiguous(memory_format=cl):
        raise ValueError("src must be a channels-last CUDA tensor")
    B, Cc, H, W = src.shape
    if residual is not None and (residual.shape != src.shape or not residual.is_contiguous(memory_format=cl)):
        raise ValueError("residual must match src")
    if out2 is not None and (not out2.is_contiguous(memory_format=cl) or out2.shape[0] != B or out2.shape[2:] != src.shape[2:]):
        raise ValueError("out2 must be a channels-last tensor with the same batch and spatial size")
    with torch.cuda.device(src.device):
        cabi.call("eitb_conv_epilogue_nhwc", src.data_ptr(), _DT[src.dtype], B * H * W, Cc, _ptr(bias), int(silu),
                  _ptr(residual), src.data_ptr() if inplace else 0, _ptr(out2), out2.shape[1] if out2 is not None else 0,
                  out2_off, _stream(src))
    return src if inplace else None


def upsample2x_concat(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cat((nearest_upsample_x2(a), b), dim=1) for channels-last half tensors, in one pass."""
    cl = torch.channels_last
    if not (a.is_cuda and a.is_contiguous(memory_format=cl) and b.is_contiguous(memory_format=cl)):
        raise ValueError("a and b must be channels-last CUDA tensors")
    B, Ca, h, w = a.shape
    Cb = b.shape[1]
    assert b.shape[0] == B and b.shape[2] == 2 * h and b.shape[3] == 2 * w and a.dtype == b.dtype
    out = torch.empty((B, Ca + Cb, 2 * h, 2 * w), dtype=a.dtype, device=a.device, memory_format=cl)
    with torch.cuda.device(a.device):
        cabi.call("eitb_upsample2x_concat_nhwc", a.data_ptr(), b.data_ptr(), out.data_ptr(), _DT[a.dtype], B, h, w, Ca, Cb, _stream(a))
    return out


# ------------------------------------------------------------------------------------ K10
def yolo_head_decode(box, cls, mc, strides, nc: int, nm: int, biases=None, cls_cstride: int = 0,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """Lists of the 3 per-level branch outputs (channels-last fp16 [B,64|nc|nm,h,w]) -> head [B,4+nc+nm,A].
    ``biases`` = (box, cls, mc) lists of per-level float32 bias vectors folded into the decode."""
    import ctypes as C
    cl = torch.channels_last
    B = box[0].shape[0]
    for t in list(box) + list(cls) + list(mc):
        if not (t.is_cuda and t.dtype == torch.float16 and (cls_cstride or t.is_contiguous(memory_format=cl) or t.shape[1] == 1)):
            raise ValueError("branch outputs must be channels-last fp16 CUDA tensors")
    hs = (C.c_int * 3)(*[t.shape[2] for t in box])
    ws = (C.c_int * 3)(*[t.shape[3] for t in box])
    st = (C.c_int * 3)(*strides)
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    A = sum(t.shape[2] * t.shape[3] for t in box)
    if out is None:
        head = torch.empty((B, 4 + nc + nm, A), dtype=torch.float16, device=box[0].device)
    else:
        _chk(out, torch.float16, "out")
        assert tuple(out.shape) == (B, 4 + nc + nm, A)
        head = out
    bb = cb = mb = None
    if biases is not None:
        for group in biases:
            for t in group:
                _chk(t, torch.float32, "bias")
        bb, cb, mb = (arr(g) for g in biases)
    with torch.cuda.device(head.device):
        cabi.call("eitb_yolo_head_decode", arr(box), arr(cls), arr(mc), bb, cb, mb, hs, ws, st, B, nc, nm, cls_cstride, head.data_ptr(),
                  _stream(head))
    return head


def sppf_pool_concat(x: torch.Tensor) -> torch.Tensor:
    """x [B,C,h,w] channels-last fp16 -> [B,4C,h,w]: x | maxpool5 | maxpool9 | maxpool13."""
    if not (x.is_cuda and x.dtype == torch.float16 and x.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("x must be a channels-last fp16 CUDA tensor")
    B, Cc, h, w = x.shape
    out = torch.empty((B, 4 * Cc, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        cabi.call("eitb_sppf_pool_concat", x.data_ptr(), B, h, w, Cc, out.data_ptr(), _stream(x))
    return out


# ------------------------------------------------------------------------------------ K8
def tri_label(nodes_xy: torch.Tensor, tri: torch.Tensor, poly_xy: torch.Tensor, poly_off: torch.Tensor,
              poly_cls: torch.Tensor, outer_cls: int = 4) -> torch.Tensor:
    _chk(nodes_xy, torch.float64, "nodes_xy")
    _chk(tri, torch.int64, "tri")
    _chk(poly_xy, torch.float64, "poly_xy")
    _chk(poly_off, torch.int32, "poly_off")
    _chk(poly_cls, torch.int32, "poly_cls")
    T, P, V = tri.shape[0], poly_cls.shape[0], poly_xy.shape[0]
    out = torch.empty((T,), dtype=torch.int32, device=tri.device)
    nbytes = cabi.load().eitb_tri_label_workspace_bytes(P, V)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=tri.device)
    with torch.cuda.device(tri.device):
        cabi.call("eitb_tri_label", nodes_xy.data_ptr(), nodes_xy.shape[0], tri.data_ptr(), T, poly_xy.data_ptr(),
                  poly_off.data_ptr(), poly_cls.data_ptr(), P, V, outer_cls, out.data_ptr(), ws.data_ptr(), nbytes,
                  _stream(tri))
    return out


def tri_label_raster(nodes_xy: torch.Tensor, tri: torch.Tensor, code: torch.Tensor, outer_cls: int = 4) -> torch.Tensor:
    _chk(nodes_xy, torch.float64, "nodes_xy")
    _chk(tri, torch.int64, "tri")
    _chk(code, torch.uint8, "code")
    H, W = code.shape
    out = torch.empty((tri.shape[0],), dtype=torch.int32, device=tri.device)
    with torch.cuda.device(tri.device):
        cabi.call("eitb_tri_label_raster", nodes_xy.data_ptr(), nodes_xy.shape[0], tri.data_ptr(), tri.shape[0],
                  code.data_ptr(), H, W, outer_cls, out.data_ptr(), _stream(tri))
    return out
