"""Mirror of the hot-path functions of the reference's kt_service/ai_tools/utils.py.

Same names, argument meaning and error behaviour (every reference function wraps its body in
``try / except: logger.error(...); return <empty sentinel>`` -- e.g. utils.py:310-313,
520-523, 583-585), numpy in / numpy out; the arithmetic runs in libeitb200 on the GPU.  There is
no CPU implementation behind these: without the CUDA library they log and return the sentinel,
exactly like any other failure in the reference.

Label images travel as *code* images on the device (one byte per pixel: 0 black, 1 muscle,
3 adipose, 6 lung, 7 bone = B<<2|G<<1|R of the reference's BGR colours, utils.py:468-473);
``create_color_output`` returns the BGR image the reference returns.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ... import host, ops

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)

CLASS_NAMES = ("bone", "muscles", "lung", "adipose")        # class ids 0..3, utils.py:498-505
_DEVICE = None


def set_device(device) -> None:
    global _DEVICE
    _DEVICE = torch.device(device)


def _dev():
    global _DEVICE
    if _DEVICE is None:
        from .. import config
        _DEVICE = torch.device(config.device())
    return _DEVICE


def _to_dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(_dev())


# ---------------------------------------------------------------------------- DICOM tag helpers
def get_rescale_intercept(ds):
    """utils.py:621-637"""
    try:
        return int(ds[(0x0028, 0x1052)].value)
    except Exception:
        logger.error("get_rescale_intercept failed")
        return []


def get_rescale_slope(ds):
    """utils.py:640-656"""
    try:
        return int(ds[(0x0028, 0x1053)].value)
    except Exception:
        logger.error("get_rescale_slope failed")
        return []


def get_pixel_spacing(ds):
    try:
        return [float(v) for v in ds[(0x0028, 0x0030)].value]
    except Exception:
        logger.error("get_pixel_spacing failed")
        return []


def get_axial_slice_size(img):
    """utils.py:1282-1307: the height if it is 256 or 512, else the (clobbered) default ``[]``."""
    try:
        h = img.shape[0]
        return h if h in (256, 512) else []
    except Exception:
        logger.error("get_axial_slice_size failed")
        return []


# ---------------------------------------------------------------------------- a6 / a8 / a9
def classic_norm(volume, window_level=40, window_width=400):
    """utils.py:272-313 -- HU window -> u8 -> rotate 180 (K1)."""
    try:
        v = np.asarray(volume)
        lo, hi = window_level - window_width // 2, window_level + window_width // 2
        px = _to_dev(v.reshape((-1,) + v.shape[-2:]), np.int16)
        u8, _ = ops.hu_window(px, lo, hi, True, None, True, None)
        return u8.cpu().numpy().reshape(v.shape)
    except Exception as e:
        logger.error(f"classic_norm failed: {e}")
        return []


def get_axial_slice_body_mask(ds):
    """utils.py:526-585 (K2): flipud, HU rescale, (-500, 1000), 5x5 open, largest contour filled."""
    try:
        px = _to_dev(np.asarray(ds.pixel_array)[None], np.int16)
        m = ops.body_mask(px, get_rescale_slope(ds), get_rescale_intercept(ds), True)
        return m[0].cpu().numpy()
    except Exception as e:
        logger.error(f"get_axial_slice_body_mask failed: {e}")
        return []


def get_axial_slice_body_mask_nii(hu_img):
    """utils.py:588-618: the same on an HU image, no flip, no rescale."""
    try:
        m = ops.body_mask(_to_dev(np.asarray(hu_img)[None], np.int16), 1, 0, False)
        return m[0].cpu().numpy()
    except Exception as e:
        logger.error(f"get_axial_slice_body_mask_nii failed: {e}")
        return []


def apply_body_mask(norm_u8, body_mask):
    """cv2.bitwise_and(x, x, mask=m) of ai_tools.py:212 (fused into K1 on the batched path); on the device."""
    try:
        img, m = _to_dev(norm_u8, np.uint8), _to_dev(body_mask, np.uint8)
        return _masked(img, m).cpu().numpy()
    except Exception as e:
        logger.error(f"apply_body_mask failed: {e}")
        return []


def _masked(img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    return ops.apply_mask_u8(img.contiguous(), mask.contiguous())


# ---------------------------------------------------------------------------- a2 / a3
def convert_to_3d(i_slices):
    """utils.py:73-111: sorts the caller's list in place by int(InstanceNumber); returns the
    (H, W, N) volume *view* and the three orientation tags.  The stack itself is never
    materialised on the GPU path (only one row per slice is used)."""
    try:
        i_slices.sort(key=lambda s: int(s.InstanceNumber))
        vol = np.stack([np.asarray(s.pixel_array) for s in i_slices], axis=-1)
        tag = lambda s, t: s[t].value if _has(s, t) else None
        s0 = i_slices[0]
        return vol, tag(s0, (0x0018, 0x5100)), tag(s0, (0x0020, 0x0037)), tag(s0, (0x0020, 0x0020))
    except Exception as e:
        logger.error(f"convert_to_3d failed: {e}")
        return [], None, None, None


def _has(ds, t):
    try:
        ds[t]
        return True
    except Exception:
        return False


def front_slice_from_series(pixels, instance_numbers, patient_position="HFS", image_orientation=(1, 0, 0, 0, 1, 0),
                            patient_orientation=None):
    """convert_to_3d + axial_to_sagittal + mid-plane + cv2.normalize (utils.py:73-163,
    ai_tools.py:95-101) in one pass over the series: (N,H,W) int16 in file order -> (N,W) u8 (K3)."""
    try:
        px = _to_dev(pixels, np.int16)
        n, H, W = px.shape
        order = torch.from_numpy(host.instance_order(instance_numbers)).to(px.device)
        row, fx, fz = host.front_geometry(H, patient_position, image_orientation, patient_orientation)
        rows, mm = ops.front_rows(px, order, n, row, fx, fz)
        return ops.minmax_u8(rows, mm).cpu().numpy()
    except Exception as e:
        logger.error(f"front_slice_from_series failed: {e}")
        return []


# ---------------------------------------------------------------------------- a5
def search_number_axial_slice(detections, custom_number_slise=0, image_width=512):
    """utils.py:166-269 (K4): ``detections.xyxy`` (k,4) float32 -> [y6, y7, mid+custom] or []."""
    try:
        xyxy = np.asarray(detections.xyxy, np.float32).reshape(1, -1, 4)
        k = xyxy.shape[1]
        pad = np.zeros((1, max(k, 1), 4), np.float32)
        pad[:, :k] = xyxy
        out = ops.rib_select(_to_dev(pad, np.float32), torch.tensor([k], dtype=torch.int32, device=_dev()),
                             float(image_width), torch.tensor([custom_number_slise], dtype=torch.int32, device=_dev()))
        y6, y7, mid, ok = out[0].cpu().tolist()
        if not ok:
            raise IndexError("fewer than 7 ribs right of the midline")
        return [y6, y7, mid]
    except Exception as e:
        logger.error(f"search_number_axial_slice failed: {e}")
        return []


# ---------------------------------------------------------------------------- a15 / a16 / a19
def create_segmentations_masks(results, img_size=512):
    """utils.py:437-523: ``results.masks.data`` (n,S,S), ``results.boxes.cls`` -> dict of 4 BGR images."""
    try:
        masks = torch.as_tensor(results.masks.data).to(_dev()).float().contiguous()
        cls = torch.as_tensor(results.boxes.cls).to(_dev()).to(torch.int32).contiguous()
        size = int(results.orig_shape[0])
        if masks.ndim != 3 or masks.shape[0] == 0:
            masks = torch.zeros((0, size, size), dtype=torch.float32, device=_dev())
            cls = torch.zeros((0,), dtype=torch.int32, device=_dev())
        imgs = ops.class_images(masks, cls).cpu().numpy()           # one library call: union per class, painted
        return {name: imgs[c] for c, name in enumerate(CLASS_NAMES)}
    except Exception as e:
        logger.error(f"create_segmentations_masks failed: {e}")
        return {}


def _codes_from_class_images(d, device) -> torch.Tensor:
    """overlay_segmentation_masks (utils.py:395-434) on the device: the saturating colour adds of the four
    class images are the OR of their 3-bit colour codes.  Returns the (S, S) u8 code image."""
    code = None
    for name, val in (("bone", 7), ("muscles", 1), ("lung", 6), ("adipose", 3)):
        if name in d:
            img = torch.from_numpy(np.ascontiguousarray(d[name], dtype=np.uint8)).to(device)
            if code is None:
                code = torch.zeros(img.shape[:2], dtype=torch.uint8, device=device)
            ops.bgr_or_code(img, val, code)
    return code


def create_color_output(segmentation_masks_image, only_body_mask=None):
    """utils.py:989-1010 (K7): overlay -> clear_color_output (if a body mask is given) ->
    highlight_small_masks; returns the BGR label image."""
    try:
        if segmentation_masks_image is None or len(segmentation_masks_image) == 0:
            return None
        code = _codes_from_class_images(segmentation_masks_image, _dev())[None].contiguous()
        body = None if only_body_mask is None else _to_dev(np.asarray(only_body_mask)[None], np.uint8)
        ops.label_cleanup(code, body)
        return ops.codes_to_bgr(code)[0].cpu().numpy()
    except Exception as e:
        logger.error(f"create_color_output failed: {e}")
        return []


def codes_to_color(code) -> np.ndarray:
    """(S,S) u8 code image -> the reference's BGR label image (colours of utils.py:468-473) through K7's
    eitb_codes_to_bgr."""
    c = _to_dev(np.ascontiguousarray(code)[None], np.uint8)
    return ops.codes_to_bgr(c)[0].cpu().numpy()


def device_polygons(code, only_body_mask=None):
    """K13 on one code image (numpy or device tensor): the device-resident polygon list (``ops.LabelPolygons``).
    Capacities grow on overflow, so pathological (pure-noise) label images are still answered on the device."""
    c = code if torch.is_tensor(code) else _to_dev(np.ascontiguousarray(code), np.uint8)
    c = c.reshape((1,) + tuple(c.shape[-2:])).contiguous()
    b = None
    if only_body_mask is not None:
        b = only_body_mask if torch.is_tensor(only_body_mask) else _to_dev(np.ascontiguousarray(only_body_mask), np.uint8)
        b = b.reshape(c.shape).contiguous()
    H, W = c.shape[-2:]
    for max_polys, max_points in ((1024, H * W // 4), (H * W // 2, 3 * H * W)):
        lp = ops.label_polygons(c, b, max_polys=max_polys, max_points=max_points)
        if not int(lp.status[0]) & 7:
            return lp
    raise RuntimeError("label image too fragmented for the polygon scratch buffers")


def codes_to_polygons(code, pixel_spacing, only_body_mask=None, device_result=None):
    """create_list_crd_from_color_output (utils.py:1191-1279) on a code image, through K13 (contour following,
    arcLength, approxPolyDP and the body outline run on the device; the host only formats the strings)."""
    lp = device_result if device_result is not None else device_polygons(code, only_body_mask)
    (status, polys), = lp.to_host()
    out = [f"{cls} " + " ".join(f"{x} {y}" for x, y in pts) for cls, pts in polys]
    if only_body_mask is not None and status & 8:
        out.append([])                                            # utils.py:1165: no outline -> the empty list is appended
    return [str(pixel_spacing[0]), str(pixel_spacing[1])] + out


def create_list_crd_from_color_output(color_output, pixel_spacing, only_body_mask=None):
    """utils.py:1191-1279: BGR label image -> ['sx', 'sy', 'cls x y x y ...', ...]."""
    try:
        img = np.asarray(color_output)
        code = (((img[..., 0] > 0) << 2) | ((img[..., 1] > 0) << 1) | (img[..., 2] > 0)).astype(np.uint8)
        return codes_to_polygons(code, pixel_spacing, only_body_mask)
    except Exception as e:
        logger.error(f"create_list_crd_from_color_output failed: {e}")
        return []
