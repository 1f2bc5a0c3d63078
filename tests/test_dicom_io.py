"""Series ingest: the built-in uncompressed-DICOM reader, CPU only."""
import io
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
