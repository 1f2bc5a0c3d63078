"""Model locations and service settings, under the attribute names the reference service imports
(kt_service/kt_service_config.py:1-13; read at kt_service/ai_tools/ai_tools.py:51,58,63).

The weight files live in ``$EITB_WEIGHTS_DIR`` (default: the volume the reference's
docker-compose mounts).  They are not distributed with either project; when a file is absent the
pipeline builds the seeded random-init YOLO11s-seg of ``eitsynthai_b200.yolo_seg`` instead.
"""
import os as _os

_WEIGHTS_DIR = _os.environ.get("EITB_WEIGHTS_DIR", "/app/weights")
_FILES = {
    "ribs_segm_model": "yolov11s_ribs_16_02_100ep_16batch_640_best.pt",              # coronal rib detector, imgsz 640
    "axial_slice_segm_model_256": "yolov11s_axial_11_09_50ep_16batch_256_best.pt",    # 256-pixel axial slices
    "axial_slice_segm_model_512": "yolov11s_axial_16_04_100ep_16batch_512_best.pt",   # 512-pixel axial slices
}
globals().update({name: _os.path.join(_WEIGHTS_DIR, fname) for name, fname in _FILES.items()})

# scalar settings duplicated from ai_fsi_config.toml in the reference (kt_service_config.py:6-13)
service_version, save_log_path = "1.0", ["ai_logs"]
device, weights_ribs, weights_segmentation = "", "/", "/"
