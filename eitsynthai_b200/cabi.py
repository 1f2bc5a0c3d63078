"""ctypes binding of libeitb200.so -- the C-ABI boundary declared in ``include/eitb200.h``.

Nothing here falls back to another implementation: a missing or unloadable library raises
``EitbLibraryError`` and a non-zero status from an entry point raises ``EitbError``.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libeitb200.so")

OK, ERR_BAD_ARG, ERR_WORKSPACE, ERR_LAUNCH, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
F32, F16, BF16 = 0, 1, 2
CODE_BLACK, CODE_MUSCLE, CODE_ADIPOSE, CODE_LUNG, CODE_BONE = 0, 1, 3, 6, 7


class EitbLibraryError(RuntimeError):
    pass


class EitbError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed: {msg} ({code})")
        self.code = code


_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

#: name -> (restype, argtypes); mirrors include/eitb200.h one to one
SIGNATURES = {
    "eitb_strerror": (C.c_char_p, [_i]),
    "eitb_version": (_i, []),
    "eitb_profile_enable": (_i, [_i]),
    "eitb_profile_report": (C.c_longlong, [C.c_char_p, _sz]),
    "eitb_hu_window_nchw": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p]),
    "eitb_u8_to_nchw": (_i, [_p, _i, _i, _i, _p, _i, _p]),
    "eitb_body_mask_workspace_bytes": (_sz, [_i, _i, _i]),
    "eitb_body_mask": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "eitb_cc_label": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "eitb_front_rows": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "eitb_front_rows_batch": (_i, [_p, C.c_longlong, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "eitb_rows_h2d": (_i, [_p, C.c_longlong, _i, _i, _i, _p, _p]),
    "eitb_minmax_u8": (_i, [_p, _i64, _p, _p, _p]),
    "eitb_letterbox_nchw": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "eitb_rib_select": (_i, [_p, _p, _i, _i, _f, _p, _p, _p]),
    "eitb_scale_boxes": (_i, [_p, _p, _i, _i, _i, _f, _f, _f, _f, _f, _p, _p]),
    "eitb_nms_workspace_bytes": (_sz, [_i, _i]),
    "eitb_nms": (_i, [_p, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _sz, _p]),
    "eitb_mask_decode_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eitb_mask_decode": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "eitb_label_cleanup_workspace_bytes": (_sz, [_i, _i, _i]),
    "eitb_label_cleanup": (_i, [_p, _p, _i, _i, _i, _p, _sz, _p]),
    "eitb_codes_to_bgr": (_i, [_p, _p, _i64, _p]),
    "eitb_apply_mask_u8": (_i, [_p, _p, _i64, _i, _p, _p]),
    "eitb_class_images": (_i, [_p, _p, _i, _i64, _p, _p]),
    "eitb_bgr_or_code": (_i, [_p, _i64, _i, _p, _p]),
    "eitb_bias_act_nhwc": (_i, [_p, _i, C.c_longlong, _i, _p, _i, _p]),
    "eitb_conv_epilogue_nhwc": (_i, [_p, _i, C.c_longlong, _i, _p, _i, _p, _p, _p, _i, _i, _p]),
    "eitb_upsample2x_concat_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "eitb_yolo_head_decode": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "eitb_sppf_pool_concat": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "eitb_conv2d_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p]),
    "eitb_conv2d_tuning": (_i, [_i, _i, _i]),
    "eitb_conv2d_debug": (_i, [_i]),
    "eitb_stem_debug": (_i, [_i]),
    "eitb_stem_conv3x3s2_nhwc": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _i, _p, _i, _i, _p]),
    "eitb_dwconv3x3_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p]),
    "eitb_tri_label_workspace_bytes": (_sz, [_i, _i]),
    "eitb_tri_label": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "eitb_tri_label_raster": (_i, [_p, _i64, _p, _i64, _p, _i, _i, _i, _p, _p]),
    "eitb_rle_decode_frame": (_i, [C.c_char_p, _sz, _i, _i, _i, _p]),
    "eitb_jpeg_lossless_decode": (_i, [C.c_char_p, _sz, _p, _p, _p, _p, _sz]),
    "eitb_label_polygons_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "eitb_label_polygons": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "eitb_polygons_for_mesh_workspace_bytes": (_sz, [_i, _i]),
    "eitb_polygons_for_mesh": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libeitb200.so (built in-tree by ``python -m eitsynthai_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EitbLibraryError(
            f"{LIB_PATH} is missing: run `python -m eitsynthai_b200.build` (there is no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:                                     # pragma: no cover
        raise EitbLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise EitbLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().eitb_strerror(code).decode()


def call(name: str, *args) -> None:
    """Call a status-returning entry point; raise ``EitbError`` on a non-zero status."""
    rc = getattr(load(), name)(*args)
    if rc != OK:
        raise EitbError(name, rc, strerror(rc))


def profile_enable(on: bool) -> None:
    load().eitb_profile_enable(int(on))


def profile_report() -> dict:
    """{kernel name: (launches, total ms)} recorded since ``profile_enable(True)``."""
    lib = load()
    need = lib.eitb_profile_report(None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    lib.eitb_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out
