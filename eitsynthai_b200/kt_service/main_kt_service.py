"""Mirror of the reference's kt_service/main_kt_service.py: the same five POST routes
(main_kt_service.py:33,50,69,88,127) over the GPU pipelines.  Pipelines are built on first use
(the reference builds them at import, :24-28).  Responses carry the hot-path results (label codes
as PNG-free JSON lists are large, so only shapes / polygons / mesh classes are returned); the
collage, the EIT simulation and the .dat file belong to the out-of-scope tail of the service."""
from __future__ import annotations

import io
import zipfile

import numpy as np
from fastapi import FastAPI, File, UploadFile
from fastapi.responses import JSONResponse

app = FastAPI()
_objs = {}


def _get(name):
    if name not in _objs:
        from .ai_tools import ai_tools
        _objs[name] = getattr(ai_tools, name)()
    return _objs[name]


def _pack(answer):
    if not answer:                                                  # the pipelines' failure sentinel, returned as is (ai_tools.py:203,229-231)
        return JSONResponse(content=[])
    return JSONResponse(content={"status": answer["status"], "message": answer["message"],
                                 "segmentation_time": answer["segmentation_time"], "text_data": answer["text_data"],
                                 "label_shape": list(answer["label_codes"].shape), "polygons": answer["polygons"],
                                 "mesh_classes": answer["mesh_data"]["CLASS"] if answer["mesh_data"] else []})


async def _run(file: UploadFile, cls_name: str, method: str):
    try:
        buf = io.BytesIO(await file.read())
        if not zipfile.is_zipfile(buf):                             # main_kt_service.py:42-44
            raise zipfile.BadZipFile
        buf.seek(0)
        return _pack(getattr(_get(cls_name), method)(buf))
    except zipfile.BadZipFile:
        return JSONResponse(status_code=400, content={"status": "error", "message": "bad zip file"})
    except Exception as e:                                          # noqa: BLE001
        return JSONResponse(status_code=500, content={"status": "error", "message": str(e)})


@app.post("/uploadDicomSequence")
async def upload_file(file: UploadFile = File(...)):
    return await _run(file, "DICOMSequencesToMask", "get_coordinate_slice_from_dicom")


@app.post("/uploadDicomSequenceCustom")
async def upload_file_custom(file: UploadFile = File(...)):
    return await _run(file, "DICOMSequencesToMaskCustom", "get_coordinate_slice_from_dicom_custom")


@app.post("/uploadDicomFrame")
async def upload_dicom_frame(file: UploadFile = File(...)):
    return await _run(file, "DICOMToMask", "get_coordinate_slice_from_dicom_frame")


@app.post("/uploadImageAxialSlice")
async def upload_image(file: UploadFile = File(...)):
    try:
        from PIL import Image
        with zipfile.ZipFile(io.BytesIO(await file.read())) as zf:
            img = np.array(Image.open(io.BytesIO(zf.read(zf.namelist()[0]))))
        return _pack(_get("ImageToMask").get_coordinate_slice_from_image(img))
    except zipfile.BadZipFile:
        return JSONResponse(status_code=400, content={"status": "error", "message": "bad zip file"})
    except Exception as e:                                          # noqa: BLE001
        return JSONResponse(status_code=500, content={"status": "error", "message": str(e)})


@app.post("/uploadNII")
async def upload_nii(file: UploadFile = File(...)):
    return await _run(file, "NIIToMask", "get_coordinate_slice_from_nii")
