"""Series ingest for the hot path: zip of DICOM files -> pinned int16 slices + the tags the path reads.

The reference decodes every zip member with ``pydicom.dcmread`` (utils.py:26-70, pydicom==3.0.1,
not installed here) and then copies all pixels twice (``np.stack(axis=-1)``, utils.py:107).  This
module is the first row of SURVEY §8(f): a minimal reader for the two *uncompressed* little-endian
transfer syntaxes CT scanners export (Implicit VR 1.2.840.10008.1.2 and Explicit VR
1.2.840.10008.1.2.1) that hands the PixelData bytes straight to a pinned buffer, one copy, ready for
the host->device stream.  Deflated Explicit VR, RLE Lossless and JPEG Lossless (process 14) files are decoded
too (zlib; libeitb200's host codecs, csrc/codec_host.cu); the remaining compressed syntaxes (JPEG-LS, JPEG 2000,
lossy JPEG) raise ``UnsupportedTransferSyntax`` (pydicom is used for those when it is importable).  ``write_dicom`` produces files for the synthetic series of the
tests; PARITY UNPINNED against pydicom itself (not available offline) -- the reader is checked on
files it did not write only through the DICOM standard's layout (PS3.5 §7.1, §7.5, PS3.10 §7.1).
"""
from __future__ import annotations

import io
import struct
import zipfile

import numpy as np

IMPLICIT_LE = "1.2.840.10008.1.2"
EXPLICIT_LE = "1.2.840.10008.1.2.1"
DEFLATED_LE = "1.2.840.10008.1.2.1.99"                     # the data set is one raw-deflate stream
RLE_LOSSLESS = "1.2.840.10008.1.2.5"
JPEG_LOSSLESS = ("1.2.840.10008.1.2.4.57", "1.2.840.10008.1.2.4.70")   # process 14, any selection value / SV1
_LONG_VR = {b"OB", b"OW", b"OF", b"SQ", b"UT", b"UN", b"OD", b"OL", b"UC", b"UR", b"OV", b"SV", b"UV"}
#: the tags the hot path reads (utils.py:46-105, 621-656; ai_tools.py:337) -> VR for implicit files
_WANTED = {
    (0x0002, 0x0010): "UI", (0x0020, 0x000E): "UI", (0x0020, 0x0013): "IS", (0x0028, 0x0010): "US",
    (0x0028, 0x0011): "US", (0x0028, 0x0100): "US", (0x0028, 0x0103): "US", (0x0028, 0x1052): "DS",
    (0x0028, 0x1053): "DS", (0x0028, 0x0030): "DS", (0x0018, 0x5100): "CS", (0x0020, 0x0037): "DS",
    (0x0020, 0x0020): "CS", (0x0028, 0x0002): "US", (0x7FE0, 0x0010): "OW",
}


class UnsupportedTransferSyntax(ValueError):
    pass


class _Elem:
    def __init__(self, value):
        self.value = value


class Dataset:
    """Duck-typed like the pydicom dataset the reference uses: ``.pixel_array``, ``.InstanceNumber``,
    ``.SeriesInstanceUID`` and ``ds[(group, element)].value``."""

    def __init__(self):
        self._tags = {}
        self._pixel_bytes = None
        self._pixel_off = 0
        self._pixel_len = 0

    def __getitem__(self, key):
        return _Elem(self._tags[tuple(key)])

    def __contains__(self, key):
        return tuple(key) in self._tags

    def get(self, key, default=None):
        return self._tags.get(tuple(key), default)

    @property
    def InstanceNumber(self):
        return self._tags[(0x0020, 0x0013)]

    @property
    def SeriesInstanceUID(self):
        return self._tags.get((0x0020, 0x000E), "")

    @property
    def shape(self):
        return int(self._tags[(0x0028, 0x0010)]), int(self._tags[(0x0028, 0x0011)])

    @property
    def pixel_dtype(self):
        if int(self._tags.get((0x0028, 0x0100), 16)) != 16 or int(self._tags.get((0x0028, 0x0002), 1)) != 1:
            raise UnsupportedTransferSyntax("only 16-bit single-sample images are on the hot path")
        return np.int16 if int(self._tags.get((0x0028, 0x0103), 1)) == 1 else np.uint16

    def pixel_view(self) -> np.ndarray:
        """Zero-copy (H, W) view of the stored pixel values inside the file buffer."""
        h, w = self.shape
        return np.frombuffer(self._pixel_bytes, dtype=self.pixel_dtype, count=h * w, offset=self._pixel_off).reshape(h, w)

    @property
    def pixel_array(self) -> np.ndarray:
        return self.pixel_view()


def _convert(vr: str, raw: bytes):
    if vr in ("US",):
        vals = struct.unpack("<%dH" % (len(raw) // 2), raw)
        return vals[0] if len(vals) == 1 else list(vals)
    txt = raw.decode("latin-1").rstrip(" \x00")
    if vr == "IS":
        parts = [int(p) for p in txt.split("\\") if p.strip()]
        return parts[0] if len(parts) == 1 else parts
    if vr == "DS":
        parts = [float(p) for p in txt.split("\\") if p.strip()]
        return parts[0] if len(parts) == 1 else parts
    if vr == "CS":
        parts = [p.strip() for p in txt.split("\\")]
        return parts[0] if len(parts) == 1 else parts
    return txt


def _skip_undefined(buf: bytes, pos: int, explicit: bool) -> int:
    """Skip a sequence / item of undefined length starting at ``pos`` (just after its header)."""
    while pos + 8 <= len(buf):
        g, e = struct.unpack_from("<HH", buf, pos)
        if g == 0xFFFE:
            (ln,) = struct.unpack_from("<I", buf, pos + 4)
            pos += 8
            if e in (0xE0DD, 0xE00D):                       # sequence / item delimiter
                return pos
            if e == 0xE000:                                 # item
                pos = _skip_undefined(buf, pos, explicit) if ln == 0xFFFFFFFF else pos + ln
            continue
        pos, _, _, _ = _next_element(buf, pos, explicit, skip_only=True)
    return pos


def _next_element(buf: bytes, pos: int, explicit: bool, skip_only: bool = False):
    """Parse one data element header; returns (position after the element, tag, vr, (value offset, length))."""
    g, e = struct.unpack_from("<HH", buf, pos)
    tag = (g, e)
    if explicit and g != 0xFFFE:
        vr = buf[pos + 4:pos + 6]
        if vr in _LONG_VR:
            (ln,) = struct.unpack_from("<I", buf, pos + 8)
            voff = pos + 12
        else:
            (ln,) = struct.unpack_from("<H", buf, pos + 6)
            voff = pos + 8
        vrs = vr.decode("latin-1")
    else:
        (ln,) = struct.unpack_from("<I", buf, pos + 4)
        voff = pos + 8
        vrs = _WANTED.get(tag, "UN")
    if ln == 0xFFFFFFFF:
        if tag == (0x7FE0, 0x0010):
            return voff, tag, vrs, (voff, -1)                # encapsulated: fragments follow (read_dicom decodes them)
        return _skip_undefined(buf, voff, explicit), tag, vrs, (voff, 0)
    return voff + ln, tag, vrs, (voff, ln)


def read_dicom(data: bytes) -> Dataset:
    """Parse the tags of ``_WANTED`` and locate PixelData; the pixel bytes are not copied."""
    buf = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
    ds = Dataset()
    pos = 0
    syntax = IMPLICIT_LE
    if len(buf) >= 132 and bytes(buf[128:132]) == b"DICM":
        pos = 132
        while pos + 8 <= len(buf) and struct.unpack_from("<H", buf, pos)[0] == 0x0002:      # file meta: explicit VR LE
            pos, tag, vr, (voff, ln) = _next_element(buf, pos, True)
            if tag == (0x0002, 0x0010):
                syntax = bytes(buf[voff:voff + ln]).decode("latin-1").rstrip(" \x00")
    if syntax not in (IMPLICIT_LE, EXPLICIT_LE, DEFLATED_LE, RLE_LOSSLESS) + JPEG_LOSSLESS:
        raise UnsupportedTransferSyntax(syntax)                    # JPEG-LS, JPEG 2000, lossy JPEG, big endian: pydicom's job
    if syntax == DEFLATED_LE:
        import zlib
        buf = bytes(buf[:pos]) + zlib.decompress(bytes(buf[pos:]), -15)
    explicit = syntax != IMPLICIT_LE
    ds._tags[(0x0002, 0x0010)] = syntax
    while pos + 8 <= len(buf):
        pos, tag, vr, (voff, ln) = _next_element(buf, pos, explicit)
        if tag == (0x7FE0, 0x0010):
            if ln < 0:                                             # encapsulated pixel data
                if syntax != RLE_LOSSLESS and syntax not in JPEG_LOSSLESS:
                    raise UnsupportedTransferSyntax(f"encapsulated PixelData in {syntax}")
                px = _decode_encapsulated(buf, voff, syntax, ds)
                ds._pixel_bytes, ds._pixel_off, ds._pixel_len = px, 0, len(px)
            else:
                ds._pixel_bytes, ds._pixel_off, ds._pixel_len = buf, voff, ln
            break
        if tag in _WANTED:
            ds._tags[tag] = _convert(_WANTED[tag] if not explicit else vr, bytes(buf[voff:voff + ln]))
    if ds._pixel_bytes is None:
        raise ValueError("no PixelData element")
    return ds


def _fragments(buf, pos: int):
    """Items of an encapsulated PixelData element (PS3.5 A.4): the basic offset table, then the fragments."""
    frags, first = [], True
    while pos + 8 <= len(buf):
        g, e, ln = struct.unpack_from("<HHI", buf, pos)
        pos += 8
        if (g, e) == (0xFFFE, 0xE0DD):
            break
        if (g, e) != (0xFFFE, 0xE000) or ln == 0xFFFFFFFF:
            raise ValueError("malformed encapsulated PixelData")
        if first:
            first = False                                         # basic offset table (possibly empty)
        else:
            frags.append((pos, ln))
        pos += ln
    return frags


def _decode_encapsulated(buf, pos: int, syntax: str, ds: "Dataset") -> bytes:
    """One frame of RLE Lossless or JPEG Lossless pixel data -> little-endian stored values, decoded by
    libeitb200's host codecs (csrc/codec_host.cu; the reference leaves this to pydicom + pylibjpeg)."""
    import ctypes as C

    from ... import cabi
    lib = cabi.load()
    h, w = ds.shape
    bps = (int(ds._tags.get((0x0028, 0x0100), 16)) + 7) // 8
    frags = _fragments(buf, pos)
    if not frags:
        raise ValueError("no pixel data fragment")
    raw = bytes(buf)
    if syntax == RLE_LOSSLESS:
        off, ln = frags[0]                                        # one fragment per frame
        out = (C.c_uint8 * (h * w * bps))()
        rc = lib.eitb_rle_decode_frame(raw[off:off + ln], ln, h, w, bps, out)
        if rc != 0:
            raise ValueError(f"RLE frame: {cabi.strerror(rc)}")
        return bytes(out)
    stream = b"".join(raw[o:o + n] for o, n in frags)             # a JPEG stream may be split over fragments
    rows, cols, prec = C.c_int(0), C.c_int(0), C.c_int(0)
    out = (C.c_uint16 * (h * w))()
    rc = lib.eitb_jpeg_lossless_decode(stream, len(stream), C.byref(rows), C.byref(cols), C.byref(prec), out, h * w)
    if rc != 0 or (rows.value, cols.value) != (h, w):
        raise ValueError(f"JPEG lossless frame: {cabi.strerror(rc) if rc else 'size mismatch'}")
    return bytes(out)


# ------------------------------------------------------------------------------------------ writer
def _el(tag, vr: str, value: bytes, explicit: bool) -> bytes:
    if len(value) % 2:
        value += b"\x00" if vr in ("UI", "OB", "OW") else b" "
    if explicit:
        if vr.encode() in _LONG_VR:
            return struct.pack("<HH2sHI", tag[0], tag[1], vr.encode(), 0, len(value)) + value
        return struct.pack("<HH2sH", tag[0], tag[1], vr.encode(), len(value)) + value
    return struct.pack("<HHI", tag[0], tag[1], len(value)) + value


def write_dicom(pixel_array: np.ndarray, instance_number: int = 1, series_uid: str = "1.2.826.0.1.3680043.8.498.1",
                intercept: float = -1024, slope: float = 1, pixel_spacing=(0.753906, 0.753906),
                patient_position: str = "HFS", iop=(1, 0, 0, 0, 1, 0), patient_orientation=None,
                explicit: bool = True, with_sequence: bool = False) -> bytes:
    """A minimal CT image file (Part-10 header, little endian) for the synthetic series."""
    px = np.ascontiguousarray(pixel_array)
    signed = px.dtype == np.int16
    assert px.dtype in (np.int16, np.uint16) and px.ndim == 2
    syntax = EXPLICIT_LE if explicit else IMPLICIT_LE
    meta = _el((0x0002, 0x0001), "OB", b"\x00\x01", True) + _el((0x0002, 0x0010), "UI", syntax.encode(), True)
    meta = _el((0x0002, 0x0000), "UL", struct.pack("<I", len(meta)), True) + meta
    ds = lambda v: "\\".join(repr(float(x)) if isinstance(x, float) else str(x) for x in (v if isinstance(v, (list, tuple)) else [v]))
    body = b""
    body += _el((0x0008, 0x0060), "CS", b"CT", explicit)
    if with_sequence:                                             # an undefined-length sequence the reader must skip
        item = _el((0x0008, 0x0100), "SH", b"113691", explicit)
        seq = struct.pack("<HHI", 0xFFFE, 0xE000, 0xFFFFFFFF) + item + struct.pack("<HHI", 0xFFFE, 0xE00D, 0)
        seq += struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
        hdr = struct.pack("<HH2sHI", 0x0008, 0x1140, b"SQ", 0, 0xFFFFFFFF) if explicit else struct.pack("<HHI", 0x0008, 0x1140, 0xFFFFFFFF)
        body += hdr + seq
    body += _el((0x0018, 0x5100), "CS", patient_position.encode(), explicit)
    body += _el((0x0020, 0x000E), "UI", series_uid.encode(), explicit)
    body += _el((0x0020, 0x0013), "IS", str(int(instance_number)).encode(), explicit)
    if patient_orientation is not None:
        body += _el((0x0020, 0x0020), "CS", "\\".join(patient_orientation).encode(), explicit)
    body += _el((0x0020, 0x0037), "DS", ds(list(iop)).encode(), explicit)
    body += _el((0x0028, 0x0002), "US", struct.pack("<H", 1), explicit)
    body += _el((0x0028, 0x0010), "US", struct.pack("<H", px.shape[0]), explicit)
    body += _el((0x0028, 0x0011), "US", struct.pack("<H", px.shape[1]), explicit)
    body += _el((0x0028, 0x0030), "DS", ds(list(pixel_spacing)).encode(), explicit)
    body += _el((0x0028, 0x0100), "US", struct.pack("<H", 16), explicit)
    body += _el((0x0028, 0x0103), "US", struct.pack("<H", 1 if signed else 0), explicit)
    body += _el((0x0028, 0x1052), "DS", ds(intercept).encode(), explicit)
    body += _el((0x0028, 0x1053), "DS", ds(slope).encode(), explicit)
    body += _el((0x7FE0, 0x0010), "OW", px.tobytes(), explicit)
    return b"\x00" * 128 + b"DICM" + meta + body


# ------------------------------------------------------------------------------------------ series level
def create_dicom_dict(zip_file: zipfile.ZipFile, reader=None):
    """utils.py:26-70: every member not ending in ``.txt`` is tried as a DICOM file; a member that does not
    parse (DICOMDIR, ``__MACOSX/._*`` resource forks, stray json, truncated files) is logged and skipped, as the
    reference's per-file try/except does (utils.py:52-60); group by SeriesInstanceUID; return the largest
    series (file order) and ``int(custom_input.txt)`` (0 when absent; the member must be named exactly so,
    utils.py:46).  Only a compressed transfer syntax aborts (``UnsupportedTransferSyntax``): the caller then
    retries with a decoder that has the codecs (``reader``)."""
    import logging
    log = logging.getLogger(__name__)
    reader = reader or read_dicom
    series, custom = {}, 0
    for name in zip_file.namelist():
        if name.endswith("/"):
            continue
        if name.endswith(".txt"):
            if name == "custom_input.txt":
                try:
                    txt = zip_file.read(name).decode().strip()
                    custom = int(txt) if txt else 0
                except Exception as e:
                    log.error(f"custom_input.txt unreadable: {e}")
            continue
        try:
            ds = reader(zip_file.read(name))
            uid = ds.SeriesInstanceUID
        except UnsupportedTransferSyntax:
            raise
        except Exception as e:
            log.error(f"skipping archive member {name}: {e}")
            continue
        series.setdefault(uid, []).append(ds)
    if not series:
        return [], custom
    return max(series.values(), key=len), custom


def series_to_pinned(i_slices):
    """Datasets of one series (file order) -> (pinned int16 tensor [N,H,W], instance numbers): ONE copy of
    every PixelData block, straight from the file buffer into page-locked memory."""
    import torch
    h, w = i_slices[0].shape
    out = torch.empty((len(i_slices), h, w), dtype=torch.int16)
    try:
        out = out.pin_memory()
    except RuntimeError:                                          # no CUDA runtime (CPU-only tests)
        pass
    dst = out.numpy()
    for k, ds in enumerate(i_slices):
        if ds.shape != (h, w):
            raise ValueError("slices of different size in one series")
        dst[k] = ds.pixel_view().view(np.int16) if ds.pixel_dtype == np.uint16 else ds.pixel_view()
    return out, np.asarray([int(ds.InstanceNumber) for ds in i_slices], np.int64)


def zip_series(slices, instance_numbers, custom: int | None = None, **tags) -> io.BytesIO:
    """Synthetic upload: what frontend_utils.dicom_sequence_to_zip posts (one file per slice, optional custom_input.txt)."""
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_STORED) as zf:
        for k, (px, inst) in enumerate(zip(slices, instance_numbers)):
            zf.writestr(f"slice_{k:04d}.dcm", write_dicom(px, int(inst), **tags))
        if custom is not None:
            zf.writestr("custom_input.txt", str(custom))
    buf.seek(0)
    return buf
