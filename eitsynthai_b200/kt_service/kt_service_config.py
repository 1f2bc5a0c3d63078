# Same names as the reference's kt_service/kt_service_config.py:1-13.  A path that does not exist
# (the Yandex-disk weights are unavailable offline) selects the seeded random-init network.
ribs_segm_model = "/app/weights/yolov11s_ribs_16_02_100ep_16batch_640_best.pt"
axial_slice_segm_model_256 = "/app/weights/yolov11s_axial_11_09_50ep_16batch_256_best.pt"
axial_slice_segm_model_512 = "/app/weights/yolov11s_axial_16_04_100ep_16batch_512_best.pt"

service_version = '1.0'
save_log_path = ['ai_logs']

device = ""
weights_ribs = '/'
weights_segmentation = '/'
