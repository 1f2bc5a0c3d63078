"""CPU statement of the per-series hot path, assembled from the oracle pieces (TEST
INFRASTRUCTURE / bench.py CPU baseline only -- never imported by the product package).

Follows DICOMSequencesToMask.get_coordinate_slice_from_dicom (ai_tools.py:188-231) stage by
stage: coronal image -> rib model -> slice pick; per slice classic_norm, body mask, bitwise_and,
ultralytics preprocess, CNN (same PyTorch module on CPU, fp32), NMS, process_mask,
create_segmentations_masks/overlay, clear_color_output, highlight_small_masks.  ``kind``: "port"
(vectorised numpy/OpenCV/torch restatement; the reference's own np.vectorize body mask is ~300x
slower than the port's).
"""
from __future__ import annotations

import numpy as np
import torch

from . import imaging as O
from . import yolo_post as Y


@torch.no_grad()
def segment_slice_cpu(px: np.ndarray, model, intercept: int = -1024, slope: int = 1, variant: str = "logit"):
    """One stored int16 slice -> final label code image (u8)."""
    size = px.shape[0]
    norm = O.classic_norm(px)
    body = O.body_mask(px, intercept, slope)
    x = Y.preprocess(O.apply_mask(norm, body), size, torch.float32)
    head, protos = model(x)
    r = Y.postprocess(head[0].float(), protos[0].float(), 4, (size, size), (size, size), variant=variant)
    union = O.class_union_masks(r["masks"].numpy(), r["cls"].numpy().astype(int), size)
    return O.create_color_codes(union, body), int(r["masks"].shape[0])


@torch.no_grad()
def rib_select_cpu(slices_sorted: np.ndarray, rib_model, custom: int = 0):
    front = O.front_slice_norm(slices_sorted)
    x = Y.preprocess(front, 640, torch.float32)
    head, _ = rib_model(x)
    dets, _ = Y.nms(head[0].float(), 1)
    boxes = Y.scale_boxes(x.shape[2:], dets[:, :4], front.shape) if dets.shape[0] else dets[:, :4]
    return O.search_number_axial_slice(boxes.numpy(), custom), front
