"""Mirror of the reference's kt_service/ai_tools/ai_tools.py: ``DICOMabc`` and its five
subclasses with the same constructor, method names and log-and-return-``[]`` behaviour
(ai_tools.py:37-450), driving the device-resident ``ImagingPipeline``.

Seams to what is outside the hot path:
* series ingest -- a zip of DICOM files is decoded with pydicom when it is importable
  (utils.py:26-70); a series can also be handed over as a list of duck-typed datasets
  (``.pixel_array``, ``.InstanceNumber``, ``ds[(group, elem)].value``), which is what the tests
  and the benchmark do (pydicom / nibabel are not installed in this image);
* Gmsh meshing -- ``mesh=(nodes, triangles)`` or the Delaunay stand-in (mesh_tools.femm_generator);
* the pyEIT simulation and the PNG collage (``get_synthetic_dataset``, ``create_answer``) are not
  mirrored: the answer is a dict with the label image, the polygon list and ``mesh_data``.
"""
from __future__ import annotations

import abc
import logging
import time
import zipfile

import numpy as np
import torch

from .. import config, kt_service_config
from ...pipeline import ImagingPipeline, SeriesMeta
from . import utils
from .mesh_tools.femm_generator import create_mesh
from ...ops import u8_to_nchw as ops_u8_to_nchw

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)

_PIPELINES = {}


def _shared_pipeline(device, ribs, axial256, axial512) -> ImagingPipeline:
    """The reference builds 5 pipeline objects x 3 models at import (main_kt_service.py:24-28);
    one set of networks per (device, weight files) is enough."""
    key = (str(device), ribs, axial256, axial512)
    if key not in _PIPELINES:
        _PIPELINES[key] = ImagingPipeline(device, weights={"ribs": ribs, "axial256": axial256, "axial512": axial512})
    return _PIPELINES[key]


def _tag(ds, t, default=None):
    try:
        return ds[t].value
    except Exception:
        return default


class DICOMabc(abc.ABC):
    def __init__(self, ribs_model_path=None, axial_model_256_path=None, axial_model_512_path=None):
        # a path passed explicitly must exist, as YOLO(path) would raise at ai_tools.py:69-71; the configured
        # defaults fall back to seeded random-init networks with a warning when the files are not deployed
        import os
        for given in (ribs_model_path, axial_model_256_path, axial_model_512_path):
            if given and not os.path.exists(given):
                raise FileNotFoundError(f"model weights not found: {given}")
        self.ribs_model_path = ribs_model_path or kt_service_config.ribs_segm_model
        self.axial_model_256_path = axial_model_256_path or kt_service_config.axial_slice_segm_model_256
        self.axial_model_512_path = axial_model_512_path or kt_service_config.axial_slice_segm_model_512
        self.device = torch.device(config.device())
        utils.set_device(self.device)
        self.pipeline = _shared_pipeline(self.device, self.ribs_model_path, self.axial_model_256_path, self.axial_model_512_path)
        self.ribs_model = self.pipeline.ribs_model
        self.axial_model_256 = self.pipeline.axial_model_256
        self.axial_model_512 = self.pipeline.axial_model_512

    # ------------------------------------------------------------------ ingest seam
    def _read_series(self, zip_buffer):
        """utils.py:26-70 create_dicom_dict: largest series of the archive + custom_input.txt."""
        if isinstance(zip_buffer, (list, tuple)):
            return list(zip_buffer), 0
        if isinstance(zip_buffer, dict):
            return list(zip_buffer["slices"]), int(zip_buffer.get("custom_number_slise", 0))
        from . import dicom_io
        if isinstance(zip_buffer, (bytes, bytearray)):
            import io
            zip_buffer = io.BytesIO(zip_buffer)
        with zipfile.ZipFile(zip_buffer, "r") as zf:
            try:                                                   # uncompressed little-endian files: built-in reader
                return dicom_io.create_dicom_dict(zf)
            except dicom_io.UnsupportedTransferSyntax as e:
                try:                                               # JPEG / JPEG-LS / J2K need pydicom (+ pylibjpeg)
                    import pydicom
                    from pydicom.filebase import DicomBytesIO
                except ImportError:
                    raise RuntimeError(f"transfer syntax {e} needs pydicom + pylibjpeg, which are not installed") from e
                return dicom_io.create_dicom_dict(zf, reader=lambda b: pydicom.dcmread(DicomBytesIO(b)))

    def _series_arrays(self, i_slices):
        from . import dicom_io
        if i_slices and isinstance(i_slices[0], dicom_io.Dataset):
            px = dicom_io.series_to_pinned(i_slices)[0].numpy()   # one copy, file buffer -> pinned memory
        else:
            px = np.stack([np.asarray(s.pixel_array, np.int16) for s in i_slices])
        s0 = i_slices[0]
        meta = SeriesMeta(np.asarray([int(s.InstanceNumber) for s in i_slices]),
                          _tag(s0, (0x0018, 0x5100), "HFS"), _tag(s0, (0x0020, 0x0037), (1, 0, 0, 0, 1, 0)),
                          _tag(s0, (0x0020, 0x0020)), int(_tag(s0, (0x0028, 0x1053), 1)),
                          int(_tag(s0, (0x0028, 0x1052), -1024)),
                          tuple(float(v) for v in _tag(s0, (0x0028, 0x0030), (0.753906, 0.753906))))
        return px, meta

    def _search_front_slise(self, zip_buffer):
        """ai_tools.py:73-105 -> (front_slice u8 (N,W), pixels (N,H,W) in file order, sorted slices, custom)."""
        front, px, i_slices, custom = [], [], [], []
        try:
            i_slices, custom = self._read_series(zip_buffer)
            px, meta = self._series_arrays(i_slices)
            self._meta = meta
            front = self.pipeline.coronal(torch.from_numpy(px).to(self.device), meta).cpu().numpy()
            i_slices.sort(key=lambda s: int(s.InstanceNumber))     # convert_to_3d sorts in place, utils.py:96
        except Exception as e:
            logger.error(f"_search_front_slise failed: {e}")
        return front, px, i_slices, custom

    def _ribs_predict(self, front_slice):
        """ai_tools.py:107-127 -> ``results.Detections`` (the fields of sv.Detections the reference reads)."""
        from .results import Detections
        det = Detections.empty()
        try:
            front = torch.from_numpy(np.ascontiguousarray(front_slice)).to(self.device)
            sel, boxes, k = self.pipeline.rib_select(front[None])
            n = int(k[0])
            det = Detections(boxes[0, :n].cpu().numpy())
            self._sel = sel[0].cpu().tolist()
        except Exception as e:
            logger.error(f"_ribs_predict failed: {e}")
        return det

    def _search_axial_slice(self, detections, i_slices, custom_number_slise=0):
        """ai_tools.py:160-182"""
        axial, numbers = [], []
        try:
            numbers = utils.search_number_axial_slice(detections, custom_number_slise)
            for i in numbers:
                axial.append(i_slices[i])
        except Exception as e:
            logger.error(f"_search_axial_slice failed: {e}")
        return axial, numbers

    def _axial_slice_predict(self, px_or_u8, ds=None):
        """ai_tools.py:129-158 fused with everything up to the label image: returns
        (labels code image (S,S) u8, body mask or None, n_detections, segmentation_time)."""
        t1 = time.time()
        if ds is not None:
            px = torch.from_numpy(np.array(ds.pixel_array, np.int16)[None]).to(self.device)
            code, body, n = self.pipeline.segment(px, int(_tag(ds, (0x0028, 0x1053), 1)), int(_tag(ds, (0x0028, 0x1052), -1024)))
            self._dev_last = (code[0], body[0])                      # stay on the device for K13 / K8
            body = body[0].cpu().numpy()
        else:
            img = np.asarray(px_or_u8)
            if img.ndim == 3 and img.shape[2] >= 3 and not (np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 0], img[..., 2])):
                # a coloured upload: all three channels, BGR -> RGB like ai_tools.py:134
                code, body, n = self.pipeline.segment_bgr_u8(torch.from_numpy(np.ascontiguousarray(img[..., :3], np.uint8)[None]).to(self.device))
            else:
                if img.ndim == 3:
                    img = img[..., 0]
                code, body, n = self.pipeline.segment_u8(torch.from_numpy(np.ascontiguousarray(img, np.uint8)[None]).to(self.device))
            self._dev_last = (code[0], None)
        out = code[0].cpu().numpy()
        return out, body, int(n[0]), round(time.time() - t1, 3)

    def predict_results(self, axial_slice):
        """The reference's ``_axial_slice_predict`` return value (ai_tools.py:129-158): ``(results, segmentation_time)``
        with ``results.masks.data`` / ``results.boxes`` / ``results.orig_shape`` as ``utils.create_segmentations_masks``
        reads them (``results.Results``).  ``axial_slice``: the normalised, body-masked u8 slice (2-D, or 3 equal
        channels).  The service itself uses the fused ``_axial_slice_predict`` above; this is the per-function path."""
        from .results import Boxes, Masks, Results
        t1 = time.time()
        img = np.asarray(axial_slice)
        if img.ndim == 3:
            img = img[..., 0]
        u8 = torch.from_numpy(np.ascontiguousarray(img, np.uint8)[None]).to(self.device)
        x = u8 if self.pipeline.fused_input else ops_u8_to_nchw(u8, self.pipeline.dtype)
        (dets, masks), = self.pipeline.predict_instances(x)
        res = Results(img.shape[:2], Boxes(dets[:, :4], dets[:, 4], dets[:, 5]), Masks(masks) if len(dets) else None,
                      {0: "bone", 1: "muscles", 2: "lung", 3: "adipose"})
        return res, round(time.time() - t1, 3)

    def _finish(self, code, body, pixel_spacing, n_det, seg_time, mesh=None, extra=None):
        """create_answer (utils.py:1019-1058): the reference's seven keys -- ``image`` (base64 PNG; here the colour
        label image, the reference sends a collage of slice and mesh render), ``text_data`` ('' as
        create_segmentation_results_cnt returns, utils.py:1013-1016), ``segmentation_time``, ``saved_file_name`` and
        ``simulation_time`` ('' / 0.0: the pyEIT simulation is outside this path), ``status``, ``message`` -- plus what
        the hot path produced: label codes, polygon list, per-element classes."""
        dev_code, dev_body = getattr(self, "_dev_last", None) or (None, None)
        self._dev_last = None
        if dev_code is None or tuple(dev_code.shape) != tuple(np.shape(code)):
            dev_code, dev_body = code, body
        elif body is None:
            dev_body = None
        elif dev_body is None:
            dev_body = body
        lp = utils.device_polygons(dev_code, dev_body)                # K13: the polygon list stays on the device for K8
        polygons = utils.codes_to_polygons(code, pixel_spacing, body, device_result=lp)
        img_mesh, mesh_data = (create_mesh(polygons[:2], polygons[2:], mesh=mesh, device_polygons=lp)
                               if (mesh is not None or body is not None) else (None, []))
        ans = {"image": self._png_base64(code), "text_data": "", "segmentation_time": seg_time, "saved_file_name": "",
               "simulation_time": 0.0, "status": "success", "message": "Processing completed successfully",
               "label_codes": code, "polygons": polygons, "mesh_data": mesh_data, "detections": n_det}
        if extra:
            ans.update(extra)
        return ans

    @staticmethod
    def _png_base64(code) -> str:
        import base64
        bgr = utils.codes_to_color(code)                             # reference colours (utils.py:468-473)
        try:                                                         # OpenCV's encoder: same image, a fraction of PIL's time
            import cv2
            ok, enc = cv2.imencode(".png", np.ascontiguousarray(bgr))    # imencode takes BGR and stores RGB, like utils.py:1037-1041
            if ok:
                return base64.b64encode(enc.tobytes()).decode("utf-8")
        except ImportError:
            pass
        from io import BytesIO

        from PIL import Image
        buf = BytesIO()
        Image.fromarray(np.ascontiguousarray(bgr[..., ::-1])).save(buf, format="PNG")   # cv2.COLOR_BGR2RGB, utils.py:1037
        return base64.b64encode(buf.getvalue()).decode("utf-8")


class DICOMSequencesToMask(DICOMabc):
    def get_coordinate_slice_from_dicom(self, zip_buffer, mesh=None):
        """ai_tools.py:188-231"""
        answer = []
        try:
            front, _, i_slices, _ = self._search_front_slise(zip_buffer)
            det = self._ribs_predict(front)
            axial, numbers = self._search_axial_slice(det, i_slices)
            ds = axial[-1]
            code, body, n, t = self._axial_slice_predict(None, ds)
            answer = self._finish(code, body, utils.get_pixel_spacing(ds), n, t, mesh, {"number_slice_eit_list": numbers})
        except Exception as e:
            logger.error(f"DICOMSequencesToMask.get_coordinate_slice_from_dicom failed: {e}")
        return answer


class DICOMSequencesToMaskCustom(DICOMSequencesToMask):
    def get_coordinate_slice_from_dicom_custom(self, zip_buffer, answer=None, mesh=None):
        """ai_tools.py:263-307"""
        answer = []
        try:
            front, _, i_slices, custom = self._search_front_slise(zip_buffer)
            det = self._ribs_predict(front)
            axial, numbers = self._search_axial_slice(det, i_slices, custom)
            ds = axial[-1]
            code, body, n, t = self._axial_slice_predict(None, ds)
            answer = self._finish(code, body, utils.get_pixel_spacing(ds), n, t, mesh, {"number_slice_eit_list": numbers})
        except Exception as e:
            logger.error(f"DICOMSequencesToMaskCustom.get_coordinate_slice_from_dicom_custom failed: {e}")
        return answer


class DICOMToMask(DICOMSequencesToMask):
    def get_coordinate_slice_from_dicom_frame(self, zip_buffer, answer=None, mesh=None):
        """ai_tools.py:315-356: the last dataset of the archive, no rib stage."""
        answer = []
        try:
            i_slices, _ = self._read_series(zip_buffer)
            ds = i_slices[-1]
            code, body, n, t = self._axial_slice_predict(None, ds)
            answer = self._finish(code, body, utils.get_pixel_spacing(ds), n, t, mesh)
        except Exception as e:
            logger.error(f"DICOMToMask.get_coordinate_slice_from_dicom_frame failed: {e}")
        return answer


class ImageToMask(DICOMSequencesToMask):
    def get_coordinate_slice_from_image(self, axial_slice_norm_body, mesh=None):
        """ai_tools.py:365-400: a normalised u8 image; no windowing, no body mask, fixed spacing."""
        answer = []
        try:
            img = np.asarray(axial_slice_norm_body)               # gray, three equal channels, or a coloured BGR image
            code, body, n, t = self._axial_slice_predict(img)
            answer = self._finish(code, None, [0.753906, 0.753906], n, t, mesh)
        except Exception as e:
            logger.error(f"ImageToMask.get_coordinate_slice_from_image failed: {e}")
        return answer


class NIIToMask(DICOMSequencesToMask):
    def get_coordinate_slice_from_nii(self, zip_buffer, answer=None, mesh=None):
        """ai_tools.py:408-450: the middle slice of the volume as an HU image.  ``zip_buffer`` is the uploaded
        zip (first ``.nii.gz`` member, nifti_io.get_nii_mean_slice) or a dict(hu=(H,W) int16, pixel_spacing=[..]).
        classic_norm rotates by 180 and the caller rotates back (ai_tools.py:430-431): net no rotation."""
        answer = []
        try:
            if isinstance(zip_buffer, dict):
                hu = np.asarray(zip_buffer["hu"], np.int16)
                spacing = list(zip_buffer.get("pixel_spacing", [0.662, 0.662]))
            else:                                                  # a zip with a .nii.gz member, like the reference
                from . import nifti_io
                if isinstance(zip_buffer, (bytes, bytearray)):
                    import io
                    zip_buffer = io.BytesIO(zip_buffer)
                with zipfile.ZipFile(zip_buffer, "r") as zf:
                    hu, spacing = nifti_io.get_nii_mean_slice(zf)
            px = torch.from_numpy(hu[None].copy()).to(self.device)
            t1 = time.time()
            from ... import ops
            body = ops.body_mask(px, 1, 0, False)
            x = self.pipeline.window_input(px, body, rot180=False)
            code, body, n = self.pipeline._segment_nchw(x, body)
            self._dev_last = (code[0], body[0])
            answer = self._finish(code[0].cpu().numpy(), body[0].cpu().numpy(), spacing, int(n[0]), round(time.time() - t1, 3), mesh)
        except Exception as e:
            logger.error(f"NIIToMask.get_coordinate_slice_from_nii failed: {e}")
        return answer
