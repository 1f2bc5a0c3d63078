"""Series ingest: the built-in uncompressed-DICOM reader, CPU only."""
import io
import struct
import zipfile

import numpy as np
import pytest

from eitsynthai_b200 import synth
from eitsynthai_b200.kt_service.ai_tools import dicom_io as D


@pytest.mark.parametrize("explicit", [True, False])
@pytest.mark.parametrize("with_sequence", [False, True])
def test_roundtrip_tags_and_pixels(explicit, with_sequence):
    px = synth.phantom_slice(3, size=64)
    raw = D.write_dicom(px, instance_number=17, intercept=-1024, slope=1, pixel_spacing=(0.7, 0.8), patient_position="FFS",
                        iop=(-1, 0, 0, 0, -1, 0), patient_orientation=("L", "P"), explicit=explicit, with_sequence=with_sequence)
    ds = D.read_dicom(raw)
    assert np.array_equal(ds.pixel_array, px) and ds.pixel_array.dtype == np.int16
    assert int(ds.InstanceNumber) == 17 and ds.SeriesInstanceUID.startswith("1.2.826")
    assert int(ds[(0x0028, 0x1052)].value) == -1024 and int(ds[(0x0028, 0x1053)].value) == 1     # utils.py:621-656 usage
    assert ds[(0x0028, 0x0030)].value == [0.7, 0.8]
    assert ds[(0x0018, 0x5100)].value == "FFS" and ds[(0x0020, 0x0020)].value == ["L", "P"]
    assert ds[(0x0020, 0x0037)].value == [-1.0, 0.0, 0.0, 0.0, -1.0, 0.0]
    assert (0x0020, 0x0020) in ds and ds.get((0x9999, 0x0001)) is None


def test_unsigned_pixels_and_bare_dataset_without_preamble():
    px = (synth.phantom_slice(1, size=32).astype(np.int32) + 2000).astype(np.uint16)
    raw = D.write_dicom(px, explicit=False)
    ds = D.read_dicom(raw)
    assert ds.pixel_array.dtype == np.uint16 and np.array_equal(ds.pixel_array, px)
    # implicit VR data set with no Part-10 header at all (old scanners)
    body = raw[raw.index(b"\x08\x00\x60\x00"):]
    assert np.array_equal(D.read_dicom(body).pixel_array, px)


def test_compressed_syntax_is_refused():
    raw = bytearray(D.write_dicom(synth.phantom_slice(0, size=32)))
    i = raw.index(D.EXPLICIT_LE.encode())
    raw[i:i + len(D.EXPLICIT_LE)] = b"1.2.840.10008.1.2.4"          # JPEG family prefix, same length - 2 -> pad
    with pytest.raises(D.UnsupportedTransferSyntax):
        D.read_dicom(bytes(raw))


def test_zip_series_largest_series_custom_input_and_pinned_copy():
    vol, inst = synth.phantom_series(12, seed=2, size=64)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as zf:
        for k in range(12):
            zf.writestr(f"a/{k}.dcm", D.write_dicom(vol[k], int(inst[k]), series_uid="1.2.3.1", explicit=k % 2 == 0))
        for k in range(3):                                           # a scout series that must be ignored
            zf.writestr(f"b/{k}.dcm", D.write_dicom(vol[k][:32, :32].copy(), k + 1, series_uid="1.2.3.2"))
        zf.writestr("custom_input.txt", "-4")
    buf.seek(0)
    with zipfile.ZipFile(buf) as zf:
        slices, custom = D.create_dicom_dict(zf)
    assert custom == -4 and len(slices) == 12
    px, numbers = D.series_to_pinned(slices)
    assert np.array_equal(px.numpy(), vol) and np.array_equal(numbers, inst)
    z = D.zip_series(vol, inst)
    with zipfile.ZipFile(z) as zf:
        s2, c2 = D.create_dicom_dict(zf)
    assert c2 == 0 and len(s2) == 12


def test_nifti_mid_slice_matches_the_reference_recipe():
    """get_nii_mean_slice restated (utils.py:1062-1119): data[:, :, Z//2] rotated 90 degrees clockwise, int16."""
    import cv2
    from eitsynthai_b200.kt_service.ai_tools import nifti_io as N
    rng = np.random.default_rng(0)
    vol = rng.integers(-1000, 2000, (48, 40, 7)).astype(np.int16)          # [i, j, k]
    for gz in (True, False):
        sl, spacing = N.read_nifti_mid_slice(N.write_nifti(vol, pixdim=(0.68, 0.71, 2.5), gz=gz))
        assert np.array_equal(sl, cv2.rotate(vol[:, :, int(7 / 2)], cv2.ROTATE_90_CLOCKWISE)) and sl.dtype == np.int16
        assert spacing == [pytest.approx(0.68), pytest.approx(0.71)]
    f32 = (vol.astype(np.float32) / 3.0)
    sl, _ = N.read_nifti_mid_slice(N.write_nifti(f32, slope=2.0, inter=-5.0))
    want = (f32[:, :, 3].astype(np.float64) * 2.0 - 5.0).astype(np.int16)
    assert np.array_equal(sl, cv2.rotate(want, cv2.ROTATE_90_CLOCKWISE))
    sl, spacing = N.read_nifti_mid_slice(N.write_nifti(vol, pixdim=(0.0, 0.7, 1.0)))
    assert spacing == [0.662, 0.662]                                       # the reference's default
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as zf:
        zf.writestr("custom_input.txt", "")
        zf.writestr("scan/volume.nii.gz", N.write_nifti(vol))
    buf.seek(0)
    with zipfile.ZipFile(buf) as zf:
        s2, _ = N.get_nii_mean_slice(zf)
    assert np.array_equal(s2, cv2.rotate(vol[:, :, 3], cv2.ROTATE_90_CLOCKWISE))


# ---------------------------------------------------------------------- compressed transfer syntaxes
def _packbits(data: bytes) -> bytes:
    """PackBits as PS3.5 G.3.1 describes it: replicate runs of >= 2 equal bytes, literal runs otherwise, each row apart."""
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        j = i
        while j + 1 < n and data[j + 1] == data[i] and j - i < 127:
            j += 1
        if j > i:
            out += bytes([(1 - (j - i + 1)) & 0xFF, data[i]])
            i = j + 1
        else:
            k = i
            while k < n and k - i < 128 and not (k + 1 < n and data[k + 1] == data[k]):
                k += 1
            out += bytes([k - i - 1]) + data[i:k]
            i = k
    return bytes(out)


def _rle_frame(px: np.ndarray) -> bytes:
    """One RLE Lossless frame of a 16-bit image: segment 0 = high bytes, segment 1 = low bytes, rows coded apart."""
    raw = px.astype("<u2")
    segs = []
    for plane in ((raw >> 8).astype(np.uint8), (raw & 0xFF).astype(np.uint8)):
        s = b"".join(_packbits(plane[r].tobytes()) for r in range(plane.shape[0]))
        segs.append(s + (b"\x00" if len(s) % 2 else b""))
    hdr = struct.pack("<16I", 2, 64, 64 + len(segs[0]), *([0] * 13))
    return hdr + segs[0] + segs[1]


def _jpeg_lossless(px: np.ndarray, sel: int = 1, precision: int = 16, restart_rows: int = 0, pt: int = 0) -> bytes:
    """A JPEG lossless (SOF3) encoder for the tests: fixed 17-symbol Huffman table, any predictor, optional restart
    interval of ``restart_rows`` image rows, point transform ``pt``."""
    H, W = px.shape
    bits = [0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0]       # code lengths 3..15 for categories 0..16
    vals = list(range(17))
    codes, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            codes[vals[k]] = (code, ln); code += 1; k += 1
        code <<= 1
    out = bytearray(b"\xff\xd8")
    out += b"\xff\xc4" + struct.pack(">H", 2 + 1 + 16 + 17) + bytes([0x00]) + bytes(bits) + bytes(vals)
    out += b"\xff\xc3" + struct.pack(">HBHHB", 8 + 3, precision, H, W, 1) + bytes([1, 0x11, 0])
    if restart_rows:
        out += b"\xff\xdd" + struct.pack(">HH", 4, restart_rows * W)
    out += b"\xff\xda" + struct.pack(">HB", 6 + 2, 1) + bytes([1, 0x00, sel, 0, pt])
    acc, nb, data = 0, 0, bytearray()

    def put(v, n):
        nonlocal acc, nb
        acc = (acc << n) | (v & ((1 << n) - 1)); nb += n
        while nb >= 8:
            b = (acc >> (nb - 8)) & 0xFF
            data.append(b)
            if b == 0xFF:
                data.append(0)
            nb -= 8
        acc &= (1 << nb) - 1

    def flush():
        nonlocal acc, nb
        if nb:
            put((1 << (8 - nb)) - 1, 8 - nb)

    v = (px.astype(np.int64) & 0xFFFF) >> pt
    rst = 0
    for y in range(H):
        if restart_rows and y and y % restart_rows == 0:
            flush()
            data += bytes([0xFF, 0xD0 + rst % 8]); rst += 1
        first_line = y == 0 or (restart_rows and y % restart_rows == 0)
        for x in range(W):
            if first_line:
                pred = (1 << (precision - pt - 1)) if x == 0 else int(v[y, x - 1])
            elif x == 0:
                pred = int(v[y - 1, 0])
            else:
                ra, rb, rc = int(v[y, x - 1]), int(v[y - 1, x]), int(v[y - 1, x - 1])
                pred = [ra, rb, rc, ra + rb - rc, ra + ((rb - rc) >> 1), rb + ((ra - rc) >> 1), (ra + rb) >> 1][sel - 1]
            d = (int(v[y, x]) - pred) & 0xFFFF
            if d >= 32768 + 1:
                d -= 65536
            if d == 32768:
                put(*codes[16]); continue
            s = 0 if d == 0 else int(abs(d)).bit_length()
            put(*codes[s])
            if s:
                put(d if d > 0 else d + (1 << s) - 1, s)
    flush()
    return bytes(out) + bytes(data) + b"\xff\xd9"


def _encapsulated_file(px, syntax, payload, fragments=1):
    """A Part-10 file in an encapsulated transfer syntax: explicit VR header, PixelData as OB of undefined length."""
    plain = D.write_dicom(px, 3)
    meta_end = plain.index(b"\x08\x00\x60\x00")                    # first data-set element (0008,0060)
    body = plain[meta_end:plain.index(struct.pack("<HH", 0x7FE0, 0x0010), meta_end)]
    meta = D._el((0x0002, 0x0001), "OB", b"\x00\x01", True) + D._el((0x0002, 0x0010), "UI", syntax.encode(), True)
    meta = D._el((0x0002, 0x0000), "UL", struct.pack("<I", len(meta)), True) + meta
    if len(payload) % 2:
        payload += b"\x00"
    step = -(-len(payload) // fragments); step += step % 2
    items = struct.pack("<HHI", 0xFFFE, 0xE000, 0)                 # empty basic offset table
    for o in range(0, len(payload), step):
        part = payload[o:o + step]
        items += struct.pack("<HHI", 0xFFFE, 0xE000, len(part)) + part
    items += struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
    pix = struct.pack("<HH2sHI", 0x7FE0, 0x0010, b"OB", 0, 0xFFFFFFFF) + items
    return b"\x00" * 128 + b"DICM" + meta + body + pix


def test_compressed_transfer_syntaxes_decode_to_the_same_pixels():
    """RLE Lossless, JPEG Lossless (every predictor, restart intervals, point transform, split fragments) and Deflated
    Explicit VR files -- written by the encoders above, following PS3.5 Annex G / ITU-T T.81 Annex H -- decode to the
    pixels of the uncompressed file (libeitb200 host codecs + zlib; no pydicom)."""
    import zlib
    px = synth.phantom_slice(5)[128:224, 100:260].copy()           # 96 x 160, int16, noisy
    px[3, 7], px[4, 7], px[40, 0] = -32768, 32767, -1               # extreme differences (category 16 / 15)
    want = D.read_dicom(D.write_dicom(px, 3)).pixel_array
    assert np.array_equal(want, px)
    ds = D.read_dicom(_encapsulated_file(px, D.RLE_LOSSLESS, _rle_frame(px)))
    assert np.array_equal(ds.pixel_array, px) and int(ds.InstanceNumber) == 3
    for sel in range(1, 8):
        ds = D.read_dicom(_encapsulated_file(px, D.JPEG_LOSSLESS[0], _jpeg_lossless(px, sel)))
        assert np.array_equal(ds.pixel_array, px), sel
    ds = D.read_dicom(_encapsulated_file(px, D.JPEG_LOSSLESS[1], _jpeg_lossless(px, 1, restart_rows=16), fragments=3))
    assert np.array_equal(ds.pixel_array, px)
    p12 = (px.astype(np.int32) & 0x0FFF).astype(np.int16)           # 12-bit data, point transform 2
    got = D.read_dicom(_encapsulated_file(p12, D.JPEG_LOSSLESS[1], _jpeg_lossless(p12, 4, precision=12, pt=2))).pixel_array
    assert np.array_equal(got, (p12 >> 2) << 2)
    plain = D.write_dicom(px, 3)
    cut = plain.index(b"\x08\x00\x60\x00")
    meta = D._el((0x0002, 0x0001), "OB", b"\x00\x01", True) + D._el((0x0002, 0x0010), "UI", D.DEFLATED_LE.encode(), True)
    meta = D._el((0x0002, 0x0000), "UL", struct.pack("<I", len(meta)), True) + meta
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    deflated = b"\x00" * 128 + b"DICM" + meta + co.compress(plain[cut:]) + co.flush()
    assert np.array_equal(D.read_dicom(deflated).pixel_array, px)
    # a syntax without a decoder here still raises the dedicated error (the pydicom seam)
    with pytest.raises(D.UnsupportedTransferSyntax):
        D.read_dicom(_encapsulated_file(px, "1.2.840.10008.1.2.4.90", b"\x00\x00"))
