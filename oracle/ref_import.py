"""Stub-import of the *actual* reference module (TEST INFRASTRUCTURE ONLY).

Loads kt_service/ai_tools/utils.py of the reference unmodified through importlib with the four
un-installed third-party modules stubbed (nibabel, pydicom, pydicom.config, pydicom.filebase,
supervision): from the source tree under /root/reference in the authoring container, or -- on the GPU
box, where that tree does not exist -- from the byte code ``oracle/build_ref.py`` compiled into
``oracle/_ref/`` (git-ignored, travels with the gpurun snapshot).  Used by ``oracle/gen_golden.py`` to
freeze golden vectors under ``tests/golden/``, by the live parity tests, and by the CPU-baseline arm of
``bench.py`` for the rows the reference owns.

Nothing under ``eitsynthai_b200/`` may import this file.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("EITB_REFERENCE_ROOT", "/root/reference")
_UTILS = os.path.join(REF_ROOT, "kt_service", "ai_tools", "utils.py")
_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cached = None


def _pyc(name: str):
    p = os.path.join(_REF_DIR, f"{name}.cpython-{sys.version_info.major}{sys.version_info.minor}.pyc")
    return p if os.path.isfile(p) else None


def reference_available() -> bool:
    return os.path.isfile(_UTILS) or _pyc("ref_utils") is not None


def _load(name: str, source: str, pyc_name: str):
    """Module from the reference source file when the tree is mounted, else from oracle/_ref byte code."""
    if os.path.isfile(source):
        spec = importlib.util.spec_from_file_location(name, source)
    else:
        from importlib.machinery import SourcelessFileLoader
        path = _pyc(pyc_name)
        if path is None:
            raise FileNotFoundError(f"{source} (and no oracle/_ref/{pyc_name}.*.pyc: run python -m oracle.build_ref)")
        spec = importlib.util.spec_from_loader(name, SourcelessFileLoader(name, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load_reference_utils():
    """Return the reference ``utils`` module (real code, stubbed imports)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(_UTILS)
    saved = {k: sys.modules.get(k) for k in
             ("nibabel", "pydicom", "pydicom.config", "pydicom.filebase", "supervision")}

    class _Settings:
        reading_validation_mode = None

    cfg = _stub("pydicom.config", settings=_Settings(), IGNORE=0)
    fb = _stub("pydicom.filebase", DicomBytesIO=object)
    pd = _stub("pydicom", config=cfg, filebase=fb, dcmread=None)
    sys.modules.update({
        "nibabel": _stub("nibabel"),
        "pydicom": pd, "pydicom.config": cfg, "pydicom.filebase": fb,
        "supervision": _stub("supervision"),
    })
    try:
        mod = _load("eitb_ref_utils", _UTILS, "ref_utils")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    import logging
    mod.logger.setLevel(logging.CRITICAL)
    _cached = mod
    return mod


class DuckDataset:
    """Duck-typed pydicom dataset: ``.pixel_array`` and ``ds[(g, e)].value``."""

    class _V:
        def __init__(self, v):
            self.value = v

    def __init__(self, pixel_array, intercept=-1024, slope=1, spacing=(0.753906, 0.753906),
                 instance_number=1, patient_position="HFS",
                 iop=(1, 0, 0, 0, 1, 0), patient_orientation=None):
        self.pixel_array = pixel_array
        self.InstanceNumber = instance_number
        self._tags = {
            (0x0028, 0x1052): intercept, (0x0028, 0x1053): slope,
            (0x0028, 0x0030): list(spacing), (0x0018, 0x5100): patient_position,
            (0x0020, 0x0037): list(iop),
        }
        if patient_orientation is not None:
            self._tags[(0x0020, 0x0020)] = patient_orientation

    def __getitem__(self, key):
        return DuckDataset._V(self._tags[tuple(key)])
