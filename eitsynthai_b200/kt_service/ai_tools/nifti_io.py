"""NIfTI-1 ingest for the ``/uploadNII`` route: the middle axial slice of a ``.nii.gz`` volume.

Restates ``get_nii_mean_slice`` (kt_service/ai_tools/utils.py:1062-1119), which goes through
nibabel==5.3.3 (not installed here): ``nib.load(path).get_fdata().astype(int16)`` -- stored values
scaled by ``scl_slope``/``scl_inter`` when the slope is non-zero, array indexed [i, j, k] with i
fastest in the file -- then slice ``k = int(Z / 2)`` (:1104-1105), ``cv2.rotate(ROTATE_90_CLOCKWISE)``
(:1106) and ``pixdim[1:3]`` as the pixel spacing when both are positive (:1093-1098, default
[0.662, 0.662]).  Only the one slice that is used is decoded and converted.  PARITY UNPINNED against
nibabel itself; the layout follows the NIfTI-1 header definition (348-byte header, vox_offset, datatype
codes 2/4/8/16/64/256/512/768).
"""
from __future__ import annotations

import gzip
import struct
import zipfile

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32}


def read_nifti_mid_slice(data: bytes):
    """Bytes of a .nii or .nii.gz file -> (slice (rows, cols) int16 after the 90-degree clockwise rotation,
    [dx, dy])."""
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    endian = "<"
    (hdr,) = struct.unpack_from("<i", data, 0)
    if hdr != 348:
        (hdr,) = struct.unpack_from(">i", data, 0)
        if hdr != 348:
            raise ValueError("not a NIfTI-1 file")
        endian = ">"
    dim = struct.unpack_from(endian + "8h", data, 40)
    datatype, bitpix = struct.unpack_from(endian + "hh", data, 70)
    pixdim = struct.unpack_from(endian + "8f", data, 76)
    vox_offset, slope, inter = struct.unpack_from(endian + "3f", data, 108)
    if bytes(data[344:347]) not in (b"n+1", b"ni1"):
        raise ValueError("bad NIfTI magic")
    if datatype not in _DTYPES or dim[0] < 3:
        raise ValueError(f"unsupported NIfTI datatype {datatype} / rank {dim[0]}")
    nx, ny, nz = dim[1], dim[2], dim[3]
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    k = int(nz / 2)
    off = int(vox_offset) + k * nx * ny * dt.itemsize
    plane = np.frombuffer(data, dtype=dt, count=nx * ny, offset=off).reshape(ny, nx).T      # [i, j]: i fastest in the file
    vals = plane.astype(np.float64)
    if slope != 0 and not np.isnan(slope):                                                # nibabel's scaling rule
        vals = vals * float(slope) + float(inter)
    sl = vals.astype(np.int16)                                                             # .astype(int16) truncates toward zero
    sl = np.ascontiguousarray(np.rot90(sl, k=-1))                                          # cv2.ROTATE_90_CLOCKWISE
    spacing = [0.662, 0.662]
    dx, dy = float(pixdim[1]), float(pixdim[2])
    if dx > 0 and dy > 0:
        spacing = [dx, dy]
    return sl, spacing


def get_nii_mean_slice(zip_file: zipfile.ZipFile):
    """utils.py:1062-1119: the first ``.nii.gz`` member of the archive."""
    for name in zip_file.namelist():
        low = name.lower()
        if low.endswith(".nii.gz") and not low.endswith(".tar.gz"):
            return read_nifti_mid_slice(zip_file.read(name))
    raise ValueError("no .nii.gz file in the archive")


def write_nifti(volume_ijk: np.ndarray, pixdim=(0.7, 0.7, 1.0), slope: float = 0.0, inter: float = 0.0, gz: bool = True) -> bytes:
    """Minimal single-file NIfTI-1 (little endian) for the tests; ``volume_ijk`` is indexed [i, j, k]."""
    code = {v: k for k, v in _DTYPES.items()}[volume_ijk.dtype.type]
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, volume_ijk.shape[0], volume_ijk.shape[1], volume_ijk.shape[2], 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, code, volume_ijk.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, pixdim[0], pixdim[1], pixdim[2], 0, 0, 0, 0)
    struct.pack_into("<3f", hdr, 108, 352.0, slope, inter)
    hdr[344:348] = b"n+1\x00"
    raw = bytes(hdr) + np.asfortranarray(volume_ijk).tobytes(order="F")
    return gzip.compress(raw, 1) if gz else raw
