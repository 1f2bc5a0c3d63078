"""Freeze golden vectors from the ACTUAL reference code (run in the authoring container).

    python oracle/gen_golden.py

imports ``/root/reference/kt_service/ai_tools/utils.py`` unmodified (stub-import, see
``oracle/ref_import.py``), runs its functions on the seeded synthetic inputs of
``eitsynthai_b200.synth`` and stores *outputs only* (inputs are regenerated from their
seeds) in ``tests/golden/reference_vectors.npz`` + ``reference_polygons.json``.
The GPU box has no reference tree; tests there read these files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eitsynthai_b200 import synth                      # noqa: E402
from oracle import yolo_post                           # noqa: E402
from oracle.ref_import import DuckDataset, load_reference_utils   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# the 19 rib boxes printed in the reference docstring, utils.py:171-189
DOCSTRING_BOXES = np.array([
    [100.45, 109.37, 116.43, 129.18], [412.88, 162.68, 426.44, 182.76],
    [90.846, 146.93, 105.72, 168.55], [67.141, 236.86, 82.394, 262.65],
    [79.154, 189.92, 94.161, 213.11], [392.32, 93.775, 409.2, 111.35],
    [114.18, 76.355, 130.5, 92.696], [317.95, 19.249, 335.81, 31.386],
    [131.96, 45.55, 147.82, 59.8], [426.9, 243.08, 439.85, 269.9],
    [180.57, 8.3435, 198.91, 21.686], [404.69, 125.41, 419.11, 144.29],
    [60.132, 291.74, 70.879, 312.55], [373.74, 62.977, 389.99, 78.234],
    [152.17, 26.801, 169.74, 38.47], [416.98, 201.76, 430.79, 226.25],
    [346.93, 39.076, 365.61, 51.212], [435.91, 303.05, 446.96, 323.76],
    [59.205, 352.68, 68.983, 362.67]], dtype=np.float32)


class _Det:
    def __init__(self, xyxy):
        self.xyxy = xyxy


class _Results:
    """Duck-typed ultralytics Results for create_segmentations_masks (utils.py:476-478)."""

    class _B:
        pass

    def __init__(self, masks_u8, cls, size):
        self.masks = _Results._B()
        self.masks.data = torch.from_numpy(masks_u8)
        self.boxes = _Results._B()
        self.boxes.cls = torch.from_numpy(np.asarray(cls, np.float32))
        self.orig_shape = (size, size)


def segmentation_case(seed: int, size: int = 512, noise: int = 0):
    """Instance masks for the label-image tests: decoded teacher heads (+ optional specks)."""
    head, protos = synth.teacher_heads(seed=seed, size=size)
    r = yolo_post.postprocess(torch.from_numpy(head), torch.from_numpy(protos), 4,
                              (size, size), (size, size))
    masks = r["masks"].numpy().copy()
    cls = r["cls"].numpy().astype(np.int64)
    if noise:
        rng = np.random.default_rng(seed + 77)
        for _ in range(noise):
            k = int(rng.integers(0, len(masks)))
            y, x = rng.integers(2, size - 4, 2)
            hh, ww = rng.integers(1, 4, 2)
            masks[k, y:y + hh, x:x + ww] ^= 1
    return masks, cls


def main():
    u = load_reference_utils()
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    polys = {}

    # ---- a6 classic_norm / a8 body mask / a9 apply, two storage conventions
    for tag, seed, intercept in (("p0", 0, -1024), ("p3hu", 3, 0)):
        px = synth.phantom_slice(seed, intercept)
        ds = DuckDataset(px, intercept=intercept, slope=1)
        norm = u.classic_norm(px)
        body = u.get_axial_slice_body_mask(ds)
        out[f"{tag}_norm"] = norm
        out[f"{tag}_body"] = body
        import cv2
        out[f"{tag}_normbody"] = cv2.bitwise_and(norm, norm, mask=body)
    out["p3hu_body_nii"] = u.get_axial_slice_body_mask_nii(synth.phantom_hu(3).astype(np.int16))
    # every representable window input -> LUT (classic_norm on a 1x65536 'image')
    allv = np.arange(-32768, 32768, dtype=np.int16).reshape(256, 256)
    out["norm_all_int16"] = u.classic_norm(allv)

    # ---- a2/a3 coronal image for several orientations
    vol, inst = synth.phantom_series(40, seed=5, size=512)
    order = np.argsort(inst, kind="stable")
    srt = vol[order]
    img3d = np.stack(list(srt), axis=-1)
    import cv2
    for tag, pp, iop, po in (("hfs", "HFS", [1, 0, 0, 0, 1, 0], None),
                             ("ffs", "FFS", [1, 0, 0, 0, 1, 0], None),
                             ("ffs_neg", "FFS", [-1, 0, 0, 0, -1, 0], ["L", "P"]),
                             ("hfp", "HFP", [1, 0, 0, 0, -1, 0], ["L", "A"])):
        sag = u.axial_to_sagittal(img3d, pp, iop, po)
        mid = sag.shape[-1] // 2
        front = np.ascontiguousarray(sag[:, :, mid])
        out[f"front_{tag}_raw"] = front
        out[f"front_{tag}_u8"] = cv2.normalize(front, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)

    # ---- a5 known-answer: docstring boxes
    out["rib_kat_custom0"] = np.array(u.search_number_axial_slice(_Det(DOCSTRING_BOXES), 0))
    out["rib_kat_custom2"] = np.array(u.search_number_axial_slice(_Det(DOCSTRING_BOXES), 2))
    out["rib_kat_few"] = np.array(u.search_number_axial_slice(_Det(DOCSTRING_BOXES[:8]), 0), dtype=np.int64)

    # ---- a15-a20 label image + polygons
    for tag, seed, size, noise, use_body in (("seg0", 0, 512, 0, True), ("seg1", 1, 512, 60, True),
                                             ("seg2", 2, 256, 25, False), ("seg3", 3, 512, 200, True)):
        masks, cls = segmentation_case(seed, size, noise)
        res = _Results(masks, cls, size)
        d = u.create_segmentations_masks(res)
        for name in ("bone", "muscles", "lung", "adipose"):
            out[f"{tag}_cls_{name}"] = d[name][..., 0] | d[name][..., 1] | d[name][..., 2]
        overlay = u.overlay_segmentation_masks(d)
        out[f"{tag}_overlay"] = overlay
        body = None
        if use_body:
            px = synth.phantom_slice(seed, -1024 if seed % 2 == 0 else 0, size=size)
            body = u.get_axial_slice_body_mask(DuckDataset(px, intercept=-1024 if seed % 2 == 0 else 0))
            out[f"{tag}_clear"] = u.clear_color_output(body, overlay)
        color = u.create_color_output(d, body)
        out[f"{tag}_color"] = color
        polys[tag] = u.create_list_crd_from_color_output(color, [0.753906, 0.753906], body)

    np.savez_compressed(os.path.join(GOLD, "reference_vectors.npz"), **out)
    with open(os.path.join(GOLD, "reference_polygons.json"), "w") as f:
        json.dump(polys, f)
    print("wrote", len(out), "arrays;", {k: len(v) for k, v in polys.items()}, "polygon lists")


if __name__ == "__main__":
    main()
