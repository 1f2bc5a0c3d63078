"""CPU oracle for the ultralytics-owned pre/post-processing (TEST INFRASTRUCTURE ONLY).

``ultralytics`` is an un-pinned, un-vendored dependency of the reference
(kt_service/requirements.txt:25) that is not installed here, so rows a11-a13 of
SURVEY.md §8 are *restated from its published algorithm* (8.3.x semantics, SURVEY.md
Appendix A.2-A.4) on torch-CPU, with ``torchvision.ops.nms`` -- the very kernel
ultralytics calls -- as the NMS ground truth.  Call sites in the reference:
``ai_tools.py:121-122`` (rib model, conf=0.3, default imgsz 640) and ``ai_tools.py:153``
(axial model, conf=0.3, imgsz 256/512).

PARITY UNPINNED at this boundary: the reference holds no test or golden vector for NMS
or mask decode; the restatement is anchored on the call-site
arguments and on how the results are consumed (utils.py:476-478, 515; ai_tools.py:123).
"""
from __future__ import annotations

import math

import cv2
import numpy as np
import torch
import torchvision

# ---------------------------------------------------------------------------- a11


def letterbox_geometry(h: int, w: int, imgsz: int, stride: int = 32):
    """LetterBox(auto=True) geometry: resized (nh, nw), padding (top, bottom, left, right)."""
    r = min(imgsz / h, imgsz / w)
    nw, nh = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - nw, imgsz - nh
    dw, dh = dw % stride, dh % stride                          # auto=True: minimum rectangle
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return nh, nw, top, bottom, left, right


def letterbox_u8(img: np.ndarray, imgsz: int) -> np.ndarray:
    """(H,W) or (H,W,3) u8 -> letterboxed u8 image (pad value 114), cv2.INTER_LINEAR resize."""
    h, w = img.shape[:2]
    nh, nw, top, bottom, left, right = letterbox_geometry(h, w, imgsz)
    if (nh, nw) != (h, w):
        img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    val = 114 if img.ndim == 2 else (114, 114, 114)
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=val)


def preprocess(gray_u8: np.ndarray, imgsz: int, dtype=torch.float32) -> torch.Tensor:
    """Gray u8 image replicated to 3 channels (cvtColor on a 2-D array, ai_tools.py:120,135)
    -> letterbox -> NCHW -> dtype -> /255 (BasePredictor.preprocess)."""
    lb = letterbox_u8(gray_u8, imgsz)
    t = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lb[None, None], (1, 3) + lb.shape)))
    t = t.to(dtype)
    t /= 255
    return t


# ---------------------------------------------------------------------------- a12


def nms(pred: torch.Tensor, nc: int, conf_thres: float = 0.3, iou_thres: float = 0.7,
        max_det: int = 300, max_wh: float = 7680.0, max_nms: int = 30000):
    """``non_max_suppression`` for one image.

    pred: (4+nc+nm, A) float32 -- xywh, class scores, mask coefficients.
    Returns ``(dets (n, 6+nm) [x1,y1,x2,y2,conf,cls,coef...], anchor_idx (n,))`` in
    descending-score order (ties by ascending anchor index: stable sort).
    """
    pred = pred.float()
    xc = pred[4:4 + nc].amax(0) > conf_thres
    x = pred.t()
    idx = torch.nonzero(xc).flatten()
    x = x[xc]
    if x.shape[0] == 0:
        return torch.zeros((0, 6 + pred.shape[0] - 4 - nc)), idx
    xy, wh = x[:, 0:2], x[:, 2:4]
    box = torch.cat((xy - wh / 2, xy + wh / 2), 1)             # xywh2xyxy
    cls = x[:, 4:4 + nc]
    mask = x[:, 4 + nc:]
    conf, j = cls.max(1, keepdim=True)
    x = torch.cat((box, conf, j.float(), mask), 1)
    sel = conf.view(-1) > conf_thres
    x, idx = x[sel], idx[sel]
    if x.shape[0] > max_nms:
        o = x[:, 4].argsort(descending=True)[:max_nms]
        x, idx = x[o], idx[o]
    c = x[:, 5:6] * max_wh
    keep = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
    return x[keep], idx[keep]


# ---------------------------------------------------------------------------- a13


def crop_mask(masks: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """Zero everything outside [x1,x2) x [y1,y2) (float comparison on arange grids)."""
    _, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, dtype=x1.dtype)[None, None, :]
    c = torch.arange(h, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def crop_mask_int(masks: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """Late-2025 ``crop_mask`` branch for ``n < 50`` masks on the CPU: the rounded
    integer box used as Python slice bounds, in place -- a negative bound counts from the end, as slicing does."""
    masks = masks.clone()
    for i, (x1, y1, x2, y2) in enumerate(boxes.round().int().tolist()):
        masks[i, :y1] = 0
        masks[i, y2:] = 0
        masks[i, :, :x1] = 0
        masks[i, :, x2:] = 0
    return masks


def process_mask(protos: torch.Tensor, coef: torch.Tensor, boxes: torch.Tensor, shape,
                 variant: str = "logit", crop: str = "float") -> torch.Tensor:
    """``ops.process_mask(..., upsample=True)``.

    protos (nm, mh, mw), coef (n, nm), boxes (n, 4) xyxy in network-input pixels,
    shape = (ih, iw).  variant "logit": 8.3.x (interpolate then > 0); "sigmoid": 8.0-8.2
    (sigmoid before crop, > 0.5 after interpolation).  crop "float": comparison on arange grids (every
    version on the GPU); "cpu": what the reference's CPU-only image runs with a late-2025 ultralytics --
    integer slicing when n < 50, the float comparison otherwise.  Returns (n, ih, iw) uint8 {0,1}.
    """
    c, mh, mw = protos.shape
    ih, iw = shape
    n = coef.shape[0]
    if n == 0:
        return torch.zeros((0, ih, iw), dtype=torch.uint8)
    masks = (coef.float() @ protos.float().view(c, -1)).view(-1, mh, mw)
    if variant == "sigmoid":
        masks = masks.sigmoid()
    ratios = torch.tensor([[mw / iw, mh / ih, mw / iw, mh / ih]])
    scaled = boxes.float() * ratios
    masks = crop_mask_int(masks, scaled) if (crop == "cpu" and n < 50) else crop_mask(masks, scaled)
    masks = torch.nn.functional.interpolate(masks[None], (ih, iw), mode="bilinear",
                                            align_corners=False)[0]
    return (masks > (0.5 if variant == "sigmoid" else 0.0)).to(torch.uint8)


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape) -> torch.Tensor:
    """Network-input px -> original-image px (gain, pad with the -0.1 rounding, clip)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    b = boxes.clone().float()
    b[:, [0, 2]] -= pad_x
    b[:, [1, 3]] -= pad_y
    b[:, :4] /= gain
    b[:, [0, 2]] = b[:, [0, 2]].clamp(0, img0_shape[1])
    b[:, [1, 3]] = b[:, [1, 3]].clamp(0, img0_shape[0])
    return b


def postprocess(pred: torch.Tensor, protos: torch.Tensor, nc: int, net_shape, orig_shape,
                conf: float = 0.3, iou: float = 0.7, variant: str = "logit",
                drop_empty: bool = True, crop: str = "float"):
    """One image through NMS -> mask decode -> box rescale -> empty-mask filter.

    Returns dict(boxes (n,4) original px, conf, cls, masks (n, ih, iw) u8, anchor_idx).
    """
    dets, idx = nms(pred, nc, conf, iou)
    masks = process_mask(protos, dets[:, 6:], dets[:, :4], net_shape, variant, crop)
    boxes = scale_boxes(net_shape, dets[:, :4], orig_shape) if dets.shape[0] else dets[:, :4]
    if drop_empty and masks.shape[0]:
        keep = masks.sum((-2, -1)) > 0
        dets, boxes, masks, idx = dets[keep], boxes[keep], masks[keep], idx[keep]
    return {"boxes": boxes, "conf": dets[:, 4], "cls": dets[:, 5], "masks": masks,
            "anchor_idx": idx, "dets": dets}
